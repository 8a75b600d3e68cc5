#!/usr/bin/env python
"""Times gs_raster_fwd / gs_raster_bwd alone on the bench scene (CUDA events, L2 flushed), and prints a
checksum of the outputs so kernel variants (GSPLAT_B200_LIB=...) can be compared.  Diagnostic only."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from importlib import import_module
from oracle import splat_oracle as so
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W, H = 1920, 1080
dev = torch.device("cuda", 0)
model = gb.GaussianModel(device=dev); model.create_from_random(N, 1.0, seed=0)
rd = gb.GaussianRenderer()
rd.fwd_tile_order = os.environ.get("GS_FWD_ORDER", rd.fwd_tile_order)
st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
cam = gb.Camera.look_at_origin_c0(W, H)
rmod = import_module("mini-3d-gaussian-splatting_b200.renderer")
w = [t.to(dev) for t in so.loss_weights(H, W)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
timer = rmod.StageTimer(); rmod.stage_timer.active = timer
chk = None
for r in range(reps + 2):
    for p in (model._xyz, model._scaling, model._rotation, model._opacity, model._features_dc, model._features_rest):
        p.grad = None
    flush.zero_()
    if r == 2: timer.events.clear()
    out = rd.render(cam, model, st)
    out["viewspace_points"].retain_grad()
    if os.environ.get("GS_LOSS") == "image":       # image-only gradient: the lean backward instantiation
        loss = (w[0] * out["image"]).sum()
    else:
        loss = (w[0] * out["image"]).sum() + (w[1] * out["alpha"]).sum() + 0.1 * (w[2] * out["depth"]).sum()
    loss.backward()
    if chk is None:
        chk = (float(out["image"].double().sum()), float(out["alpha"].double().sum()), float(out["depth"].double().sum()),
               float(out["viewspace_points"].grad.double().abs().sum()), float(model._xyz.grad.double().abs().sum()),
               float(model._scaling.grad.double().abs().sum()), float(model._opacity.grad.double().abs().sum()),
               float(model._features_dc.grad.double().abs().sum()))
rmod.stage_timer.active = None
per = timer.summary_ms()
print("lib:", os.environ.get("GSPLAT_B200_LIB", "default"), "fwd order:", rd.fwd_tile_order)
print("  " + "  ".join(f"{k}={sum(v)/len(v)*1000:.1f}us" for k, v in per.items()))
print("  checksums: " + " ".join(f"{c:.9g}" for c in chk))
tc = rd._last_debug["tile_consumed"].float()
print(f"  tile_consumed: min {float(tc.min()):.0f} mean {float(tc.mean()):.1f} max {float(tc.max()):.0f} std {float(tc.std()):.1f}; "
      f"sum of the top 2368 / total = {float(tc.sort(descending=True).values[:2368].sum() / tc.sum()):.3f}")
