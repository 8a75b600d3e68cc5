#!/usr/bin/env python
"""cProfile of the host side of end-to-end steps (per-step sync, as bench.py's e2e): where the launch path spends CPU time."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from oracle import splat_oracle as so
W, H = 1920, 1080
dev = torch.device("cuda", 0)
m = gb.GaussianModel(device=dev); m.create_from_random(1_000_000, 1.0, seed=0)
rd = gb.GaussianRenderer(); st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
cam = gb.Camera.look_at_origin_c0(W, H)
w = [t.to(dev) for t in so.loss_weights(H, W)]
buf = gb.multiview.FlatGradBuffer(m)
def loss_fn(out, vid):
    return (torch.dot(w[0].view(-1), out["image"].view(-1)) + torch.dot(w[1].view(-1), out["alpha"].view(-1))
            + 0.1 * torch.dot(w[2].view(-1), out["depth"].view(-1)))
def step():
    r = gb.multiview.multiview_step(m, rd, [cam], st, loss_fn, buffer=buf, reduce=False)
    return float(r["losses"][0].item())
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
print(f"wall per step (with per-step sync): {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
