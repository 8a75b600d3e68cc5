#!/usr/bin/env python
"""Host cost of one step: the same launch path as bench.py on a scene so small that the GPU time is negligible, so
wall time per step = what the host spends enqueueing (Python, ctypes, torch allocator, autograd).  Diagnostic only."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from importlib import import_module
losses = import_module("mini-3d-gaussian-splatting_b200.losses")
from oracle import splat_oracle as so
W, H = 128, 96
dev = torch.device("cuda", 0)
m = gb.GaussianModel(device=dev); m.create_from_random(2000, 1.0, seed=0)
rd = gb.GaussianRenderer(); st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
cam = gb.Camera.look_at_origin_c0(W, H)
w = [t.to(dev) for t in so.loss_weights(H, W)]
buf = gb.multiview.FlatGradBuffer(m)
def loss_fn(out, vid):
    return losses.weighted_sum_loss([out["image"], out["alpha"], out["depth"]], w, grads=w)
def step():
    return gb.multiview.multiview_step(m, rd, [cam], st, loss_fn, buffer=buf, reduce=False)["losses"][0]
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print(f"host cost per step (tiny scene, GPU time negligible): {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue()[:7000])
