#!/usr/bin/env python
"""Kernel timeline of steady-state bench steps (torch.profiler / CUPTI): prints every GPU activity of
the last profiled step with start offset, duration and the idle gap before it.  Diagnostic only."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from importlib import import_module
mv = import_module("mini-3d-gaussian-splatting_b200.multiview")
from oracle import splat_oracle as so

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
W, H = 1920, 1080
dev = torch.device("cuda", 0)
model = gb.GaussianModel(device=dev); model.create_from_random(N, 1.0, seed=0)
rd = gb.GaussianRenderer()
st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
cam = gb.Camera.look_at_origin_c0(W, H)
w_dev = [t.to(dev) for t in so.loss_weights(H, W)]
buf = mv.FlatGradBuffer(model)
def loss_fn(out, w):
    return (w[0] * out["image"]).sum() + (w[1] * out["alpha"]).sum() + 0.1 * (w[2] * out["depth"]).sum()
def step():
    return mv.multiview_step(model, rd, [cam], st, lambda out, vid: loss_fn(out, w_dev), buffer=buf, reduce=False)["losses"][0]
for _ in range(5): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# split into steps by big gaps (synchronize)
steps, cur, last_end = [], [], None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if last_end is not None and s - last_end > 400 and len(cur) > 10:
        steps.append(cur); cur = []
    cur.append(e); last_end = max(last_end or 0, t)
steps.append(cur)
last = steps[-1]
t0 = last[0].time_range.start
prev_end = t0
busy = 0
print(f"{'start_us':>9} {'dur_us':>8} {'gap_us':>7}  name")
for e in last:
    s, t = e.time_range.start, e.time_range.end
    print(f"{s - t0:9.1f} {t - s:8.1f} {s - prev_end:7.1f}  {e.name[:110]}")
    busy += t - s
    prev_end = max(prev_end, t)
print(f"step span {prev_end - t0:.1f} us, busy {busy:.1f} us, idle {prev_end - t0 - busy:.1f} us, kernels {len(last)}")
