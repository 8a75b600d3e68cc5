#!/usr/bin/env python
"""BASELINE.json configs [2], [3] and [4] at full size on one GPU (bench.py covers [1]); prints one JSON object.

  [2] 3 M Gaussians, 1920x1080, inference-only render throughput (no_grad), sweep over 1 M / 2 M / 3 M
  [3] multi-view training batch: M in {8, 16, 32, 64} orbit views of the 1 M scene through multiview_step
      (fused gradient + statistics accumulation), frames/s on ONE GPU (bench.py --gpus N shards the views)
  [4] densification stress: 100 k -> 2 M Gaussians at 1600x1200 by split / clone / prune rounds
"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from oracle import splat_oracle as so        # seeded loss weights only

dev = torch.device("cuda", 0)
res = {}
which = sys.argv[1:] or ["2", "3", "4", "t"]


def timed(fn, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if "2" in which:
    W, H = 1920, 1080
    rd = gb.GaussianRenderer()
    st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
    cam = gb.Camera.look_at_origin_c0(W, H)
    sweep = []
    for n in (1_000_000, 2_000_000, 3_000_000):
        m = gb.GaussianModel(device=dev)
        m.create_from_random(n, 1.0, seed=0)
        with torch.no_grad():
            for _ in range(3):
                out = rd.render(cam, m, st)
            ms = timed(lambda: rd.render(cam, m, st), 10)
        sweep.append({"splats": n, "ms_per_frame": ms, "frames_per_s": 1000.0 / ms, **rd.last_stats,
                      "image_mean": float(out["image"].mean()), "alpha_mean": float(out["alpha"].mean())})
        del m
    res["config2_inference_sweep_1080p"] = sweep

if "3" in which:
    W, H = 1920, 1080
    mv = gb.multiview
    m = gb.GaussianModel(device=dev)
    m.create_from_random(1_000_000, 1.0, seed=0)
    rd = gb.GaussianRenderer()
    rd.fwd_tile_order = os.environ.get("GS_FWD_ORDER", rd.fwd_tile_order)
    st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
    w = [t.to(dev) for t in so.loss_weights(H, W)]
    buf = mv.FlatGradBuffer(m)

    def loss_fn(out, vid):
        return (torch.dot(w[0].view(-1), out["image"].view(-1)) + torch.dot(w[1].view(-1), out["alpha"].view(-1))
                + 0.1 * torch.dot(w[2].view(-1), out["depth"].view(-1)))
    rows = []
    for M in (8, 16, 32, 64):
        cams = [gb.Camera.orbit(k, M, W, H) for k in range(M)]
        mv.multiview_step(m, rd, cams[:2], st, loss_fn, buffer=buf, reduce=False)
        ms = timed(lambda: mv.multiview_step(m, rd, cams, st, loss_fn, buffer=buf, reduce=False), 2)
        rows.append({"views": M, "ms_per_batch": ms, "frames_per_s": 1000.0 * M / ms,
                     "grad_abs_sum": float(buf.flat[:buf.param_elems].abs().sum()), "max_visits": float(buf.vis_count.max())})
    res["config3_multiview_batch_1gpu"] = rows
    del m, buf

if "4" in which:
    W, H = 1600, 1200
    m = gb.GaussianModel(device=dev)
    m.create_from_random(100_000, 1.0, seed=0)
    rd = gb.GaussianRenderer()
    st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
    cams = [gb.Camera.orbit(k, 8, W, H) for k in range(8)]
    ctrl = gb.DensityController(gb.TrainingConfig(densify_grad_threshold=1e-12))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # scene extent 0.2: splats count as "large" (split) while sigma > 0.006, i.e. for five generations of the 0.02 start
    out = gb.training.densification_stress(m, rd, cams, st, ctrl, scene_extent=0.2, target_points=2_000_000, max_rounds=40)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    with torch.no_grad():
        img = rd.render(cams[0], m, st)
    res["config4_densification_1600x1200"] = {
        "start_points": 100_000, "end_points": out["points"], "rounds": len(out["history"]), "seconds": dt,
        "history": [{k: h[k] for k in ("round", "split", "cloned", "pruned", "points", "tile_pairs")} for h in out["history"]],
        "final_frame": {**rd.last_stats, "finite": bool(torch.isfinite(img["image"]).all())}}

if "t" in which:
    # the caller of the hot path: render -> L1 loss -> backward -> fused Adam, 1 M splats, 1080p
    W, H = 1920, 1080
    m = gb.GaussianModel(device=dev)
    m.create_from_random(1_000_000, 1.0, seed=0)
    rd = gb.GaussianRenderer()
    st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
    cams = [gb.Camera.orbit(k, 8, W, H) for k in range(8)]
    with torch.no_grad():
        targets = [(rd.render(c, m, st)["image"] * 0.9 + 0.05).clone() for c in cams]
    opt = gb.GaussianOptimizer(m, gb.TrainingConfig())
    it = [0]

    def one():
        k = it[0] % 8
        gb.train_step(m, rd, cams[k], targets[k], opt, st, it[0])
        it[0] += 1
    for _ in range(5):
        one()
    ms = timed(one, 24)
    res["train_step_1M_1080p"] = {"ms_per_iteration": ms, "iterations_per_s": 1000.0 / ms, "optimizer": "torch.optim.Adam(fused=True), 5 groups",
                                  "loss": "L1 (loss.py:52)", "views": 8}

print(json.dumps(res))
