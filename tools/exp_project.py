#!/usr/bin/env python
"""Experiment: projection radii of the current library (GSPLAT_B200_LIB selects a variant) against the C port at 1 M
splats -- bit-equal fraction, integer mismatches -- and the projection kernels' CUDA-event times."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gsplat_b200 as gb
from importlib import import_module
from oracle import c_port, splat_oracle as so
rmod = import_module("mini-3d-gaussian-splatting_b200.renderer")
W, H = 1920, 1080
res = {"lib": os.environ.get("GSPLAT_B200_LIB", "default")}
for scene in ("ref", "aniso"):
    s = so.scene_ref_init(1_000_000, 0) if scene == "ref" else so.scene_aniso(1_000_000, 0)
    m = gb.GaussianModel(device="cuda")
    m.create_from_tensors(s["xyz"], s["features_dc"], s["scaling"], s["rotation"], s["opacity"])
    cam = gb.Camera.orbit(1, 8, W, H)
    cam16 = c_port.camera_block(W, H, cam._FoVx, cam._FoVy, cam.world_view_transform().numpy())
    c = c_port.project(cam16, W, H, s["xyz"].numpy(), s["scaling"].numpy(), s["rotation"].numpy(), None, s["opacity"].numpy(), True,
                       s["features_dc"].numpy().reshape(-1, 3))
    rd = gb.GaussianRenderer()
    st = gb.RenderSettings(H, W, torch.zeros(3, device="cuda"))
    out = rd.render(cam, m, st)
    v = c["vis"].astype(bool)
    rg = out["radii"].cpu().numpy()
    cg = out["conics"].detach().cpu().numpy()
    res[scene] = {"vis_mismatch": int((out["visibility_filter"].cpu().numpy() != v).sum()),
                  "radii_bit_equal": float((rg[v] == c["radii"][v]).mean()),
                  "int_radii_mismatch": int((rg[v].astype(np.int64) != c["radii"][v].astype(np.int64)).sum()),
                  "conics_bit_equal": float((cg[v] == c["conics"][v]).mean()),
                  "conics_rel": float(np.abs(cg[v] - c["conics"][v]).max() / np.abs(c["conics"][v]).max())}
    timer = rmod.StageTimer(); rmod.stage_timer.active = timer
    for _ in range(6):
        o = rd.render(cam, m, st)
        (o["image"].sum() + o["depth"].sum()).backward()
    rmod.stage_timer.active = None
    per = {k: float(np.mean(v[1:])) for k, v in timer.summary_ms().items()}
    res[scene]["ms"] = {k: round(per[k], 4) for k in ("project_fwd", "project_bwd", "bin_prepare", "bin_sort")}
print(json.dumps(res))
