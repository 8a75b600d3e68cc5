#!/usr/bin/env python
"""Steady-state GPU timeline of consecutive bench steps WITHOUT host synchronisation between them (torch.profiler /
CUPTI): idle gaps of the GPU (all streams merged) and where they fall, plus host time per step.  Diagnostic only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from importlib import import_module
mv = import_module("mini-3d-gaussian-splatting_b200.multiview")
losses = import_module("mini-3d-gaussian-splatting_b200.losses")
from oracle import splat_oracle as so

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
W, H = 1920, 1080
dev = torch.device("cuda", 0)
model = gb.GaussianModel(device=dev); model.create_from_random(N, 1.0, seed=0)
rd = gb.GaussianRenderer()
st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
cam = gb.Camera.look_at_origin_c0(W, H)
w_dev = [t.to(dev) for t in so.loss_weights(H, W)]
w_dev[2] = w_dev[2] * 0.1
buf = mv.FlatGradBuffer(model)
def loss_fn(out, vid):
    return losses.weighted_sum_loss([out["image"], out["alpha"], out["depth"]], w_dev, grads=w_dev)
def step():
    return mv.multiview_step(model, rd, [cam], st, loss_fn, buffer=buf, reduce=False)["losses"][0]
for _ in range(5): step()
torch.cuda.synchronize()
# host time per step when the GPU is not the limit: enqueue cost
t0 = time.perf_counter()
for _ in range(20): step()
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"20 steps: host returned after {t_enq*1e3/20:.3f} ms/step, GPU done after {t_all*1e3/20:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(6):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
end = t0
gaps = []
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if s - end > 2.0:
        gaps.append((s - t0, s - end, e.name[:70]))
    end = max(end, t)
span = end - t0
print(f"6 steps: span {span:.1f} us = {span/6:.1f} us/step, idle {sum(g[1] for g in gaps):.1f} us in {len(gaps)} gaps > 2 us")
for off, g, name in gaps:
    print(f"  at {off:9.1f} us: idle {g:7.1f} us before {name}")
# host-side: top CPU ops by self time
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=60))
