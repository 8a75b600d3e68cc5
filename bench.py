#!/usr/bin/env python
"""bench.py -- fwd+bwd frames/s of the render hot path at BASELINE.json's config[1]
(1 000 000 random-init Gaussians, 1920x1080, camera C0), one view per GPU per step.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference [...]      # the CPU port of the reference path (oracle/), host cores

A step = one pass of the hot path over one view per rank: render forward (project -> depth sort ->
tile binning -> compositing), the SURVEY 8d fixed-weight loss, backward (compositing and projection
backward), and for N > 1 the gradient/statistics all-reduce.  Prints ONE JSON line (rank 0).

  value      frames/s, whole job, inputs resident in HBM, timed on the device with CUDA events per
             step (L2 flushed before every step, max over ranks)
  e2e        frames/s (wall clock, max(steps, 60) steps) through the public API with every step's host inputs
             (camera + the loss-weight planes as 8-bit fixed point, 10.4 MB, pinned) copied H2D one step ahead on a
             copy stream -- issued after the frame's counter read-back, dequantised there -- and every step's loss
             copied D2H (pinned, non-blocking) and read by the host one step later, all inside the timed region
  roofline   the dominant kernel (largest share of the step): algorithmic bytes / its CUDA-event
             duration vs the measured HBM peak of MEASURED_PEAKS.json, plus `issue`: its fraction of the
             issue-slot roof that actually bounds it (profiles/r2_issue_model.md); `kernels` lists all stages
  cpu_baseline  the oracle's plain-C port of the same path on all host cores: one whole frame, nothing sampled
  GS_BENCH_TRACE=1  prints (stderr) where an end-to-end step spends its time: host enqueue, GPU span
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_SPLATS = 1_000_000
WIDTH, HEIGHT = 1920, 1080
METRIC = "fwd+bwd frames/s at 1080p, 1M Gaussians"
UNIT = "frames/s"
WORKLOAD = "config[1]: 1M random-init Gaussians (create_from_random, CPU seed 0), 1920x1080, camera C0 / orbit views, fwd+bwd"
HBM_FALLBACK_GBS = 6650.0


def bench_loss_weights_u8():
    """The step's loss target: the SURVEY 8d weight planes (image 3, alpha 1, depth 1; seeded) in 8-bit fixed point -- the
    precision ground-truth images come in -- so that the step's host inputs are 10.4 MB rather than 41.5 MB of fp32.  Every
    arm (device-resident, end-to-end, CPU) computes sum(w_img*image) + sum(w_a*alpha) + 0.1*sum(w_d*depth) with w = u8/255."""
    from oracle import splat_oracle as so
    return [(t * 255.0).round().to(torch.uint8) for t in so.loss_weights(HEIGHT, WIDTH)]


def dequantise_weights(u8_planes, out=None):
    """u8 -> fp32 weights (depth plane pre-multiplied by the loss's 0.1); the same ops on every device."""
    res = []
    for i, q in enumerate(u8_planes):
        f = out[i] if out is not None else torch.empty(q.shape, dtype=torch.float32, device=q.device)
        f.copy_(q)
        f.mul_((0.1 if i == 2 else 1.0) / 255.0)
        res.append(f)
    return res


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "10"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
            return
        t0 = time.perf_counter()                 # nvidia-smi needs a moment to attach: wait for its first row
        while time.perf_counter() - t0 < 5.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)
        self.skip = len(open(self.f.name).read().strip().splitlines())   # rows sampled before the timed region

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        lines = open(self.f.name).read().strip().splitlines()
        rows = [r.split(",") for r in lines[getattr(self, "skip", 0):] if r.count(",") >= 8]
        if not rows:
            rows = [r.split(",") for r in lines if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_frame_seconds():
    """One WHOLE config[1] frame on the C port, all host cores: projection + culling + depth sort + tile binning,
    compositing forward and backward over all 8 160 tiles, projection backward -- nothing sampled, nothing
    extrapolated.  Returns (seconds, description)."""
    from oracle import c_port, splat_oracle as so
    cores = c_port.set_num_threads(c_port.host_cores())        # torchrun exports OMP_NUM_THREADS=1: use every core anyway
    st = _cpu_state
    if not st:
        s = so.scene_ref_init(N_SPLATS, 0)
        st["params"] = {k: s[k].numpy() for k in ("xyz", "scaling", "rotation", "opacity", "features_dc")}
        cam = so.camera_c0(WIDTH, HEIGHT)
        st["cam16"] = c_port.camera_block(cam.width, cam.height, cam.fovx, cam.fovy, cam.world_view.numpy())
        st["w"] = [(q.to(torch.float32) / 255.0).numpy() for q in bench_loss_weights_u8()]
    t0 = time.perf_counter()
    c_port.render_fwd_bwd(st["cam16"], WIDTH, HEIGHT, st["params"], np.zeros(3, np.float32), st["w"])
    sec = time.perf_counter() - t0
    return sec, cores, ("C port of the reference path (oracle/splat_oracle.c, OpenMP): one whole config[1] frame -- "
                        "project + cull + depth sort + tile lists + compositing fwd+bwd over all 8160 tiles + projection backward")


_cpu_state = {}


def run_reference_arm(args, emit):
    """--impl reference: the reference's path on the host cores.  The literal Python renderer cannot run this workload
    (~15 days per forward frame), so the arm times its plain-C restatement (`kind: port`).  Every step is one
    whole frame; exactly `steps` timed steps after `warmup` untimed ones; all host cores at every N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    times, cores, sample = [], None, ""
    for i in range(warm + steps):
        sec, cores, sample = cpu_port_frame_seconds()
        if i >= warm:
            times.append(sec)
    sec = float(np.mean(times))
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": 1.0 / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    # the contract is ONE JSON line on stdout: anything a library prints there (NCCL prints its version on
    # communicator creation) is sent to stderr instead, and the result line goes to the saved descriptor
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--splats", type=int, default=N_SPLATS, help="debug only; the reported metric is defined at 1M")
    ap.add_argument("--sh-degree", type=int, default=0, choices=[0, 1, 2, 3],
                    help="0 (default) = the reference's colour, sigmoid of the DC row (renderer.py:88-92); 1..3 = the opt-in "
                         "view-dependent extension, which also reads and differentiates _features_rest [N,15,3]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args, emit)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import gsplat_b200 as gb
    from importlib import import_module
    mv = import_module("mini-3d-gaussian-splatting_b200.multiview")
    rmod = import_module("mini-3d-gaussian-splatting_b200.renderer")
    fused_losses = import_module("mini-3d-gaussian-splatting_b200.losses")
    lib = import_module("mini-3d-gaussian-splatting_b200._lib").load()
    from oracle import splat_oracle as so   # only for the seeded loss weights shared with the tests / CPU arm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.splats

    model = gb.GaussianModel(device=dev)
    model.create_from_random(n, 1.0, seed=0)          # identical scene on every rank (replicated Gaussians)
    rd = gb.GaussianRenderer(sh_degree=args.sh_degree)
    settings = gb.RenderSettings(HEIGHT, WIDTH, torch.zeros(3, device=dev))
    # one view per rank per step: rank r renders orbit view r of `world` (C0 when single-GPU)
    cam = gb.Camera.look_at_origin_c0(WIDTH, HEIGHT) if world == 1 else gb.Camera.orbit(rank, world, WIDTH, HEIGHT)
    w_host = [q.pin_memory() for q in bench_loss_weights_u8()]        # the step's host inputs: 8-bit planes, pinned
    w_dev = dequantise_weights([q.to(dev) for q in w_host])           # fp32 on the device, 0.1 folded into the depth weights
    buf = mv.FlatGradBuffer(model)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def loss_fn(out, w):
        # SURVEY 8d loss  sum(w_img*image) + sum(w_a*alpha) + 0.1*sum(w_d*depth)  (w[2] already carries the 0.1): one fused,
        # deterministic reduction over the five planes (gs_weighted_sum); its gradient is the weights themselves, so the
        # backward pass of the loss launches nothing
        return fused_losses.weighted_sum_loss([out["image"], out["alpha"], out["depth"]], w, grads=w)

    def step_device():
        res = mv.multiview_step(model, rd, [cam], settings, lambda out, vid: loss_fn(out, w_dev), buffer=buf, reduce=world > 1)
        return res["losses"][0]

    cam_wv_host = cam.world_view_transform().clone().pin_memory()
    # two staging sets: step k computes from set k%2 while the copy stream already uploads step k+1's inputs
    stage_u8 = [[torch.empty_like(t, device=dev) for t in w_host] for _ in range(2)]
    stage = [[torch.empty(t.shape, dtype=torch.float32, device=dev) for t in w_host] for _ in range(2)]
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream(device=dev)
    e2e_state = {"k": 0, "primed": False}

    def upload(slot):
        # the copy stream has been told (wait_event) when the step that last read this slot ends
        with torch.cuda.stream(copy_stream):
            for s_, h_ in zip(stage_u8[slot], w_host):
                s_.copy_(h_, non_blocking=True)
            dequantise_weights(stage_u8[slot], out=stage[slot])  # on the copy stream, off the step's critical path
            uploaded[slot].record(copy_stream)

    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_losses = []

    def collect(k):
        """Host read of step k's loss: waits for ITS copy only (enqueued at the end of step k)."""
        loss_ready[k % 2].synchronize()
        e2e_losses.append(float(loss_host[k % 2][0]))

    def step_e2e():
        # this step's host inputs: camera pose + loss weights (the "ground truth" side of the step, 8-bit planes), 10.4 MB from
        # pinned memory, uploaded one step ahead on a copy stream (input double-buffering, as a data loader does).
        # This step's result: the loss, copied D2H (pinned, non-blocking) at the end of the step and read by the
        # host once the next step has been enqueued (asynchronous logging) -- so the host never drains the GPU, but every
        # step's inputs cross H2D and every step's result crosses D2H and is consumed inside the timed region.
        k = e2e_state["k"]
        if trace is not None:                                    # GS_BENCH_TRACE=1: where an end-to-end step spends its time
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev_a.record(torch.cuda.current_stream(dev))
            t_host0 = time.perf_counter()
        if not e2e_state["primed"]:
            upload(k % 2)
            e2e_state["primed"] = True
        c = gb.Camera(WIDTH, HEIGHT, cam._FoVx, cam._FoVy, world_view=cam_wv_host)   # pose: 64 B, passed by value to the kernels

        def loss_after_upload(out, vid):
            # The next step's inputs are sent from here -- after render() has read this frame's 24-byte counters back -- so that
            # the 10 MB H2D copy never shares the PCIe link with that read-back, the one transfer the host waits for (measured
            # at 8 GPUs: with the upload issued at the top of the step the GPU-side step was 2.04 ms instead of 1.83 ms).
            if k > 0:                                            # slot (k+1)%2 was last read by step k-1: wait for it ON THE GPU
                copy_stream.wait_event(loss_ready[(k - 1) % 2])
            upload((k + 1) % 2)
            torch.cuda.current_stream(dev).wait_event(uploaded[k % 2])
            return loss_fn(out, stage[k % 2])

        res = mv.multiview_step(model, rd, [c], settings, loss_after_upload, buffer=buf, reduce=world > 1)
        loss_host[k % 2].copy_(res["losses"][0].reshape(1), non_blocking=True)
        loss_ready[k % 2].record(torch.cuda.current_stream(dev))
        e2e_state["k"] = k + 1
        if trace is not None:
            ev_b.record(torch.cuda.current_stream(dev))
            t_host1 = time.perf_counter()
        if k > len(e2e_losses):
            collect(k - 1)                                       # host read of the previous step's result, this step already enqueued
        if trace is not None:
            trace.append((t_host0, t_host1, time.perf_counter(), ev_a, ev_b))

    trace = [] if os.environ.get("GS_BENCH_TRACE") == "1" else None

    def finish_e2e():
        if e2e_state["k"] > len(e2e_losses):
            collect(e2e_state["k"] - 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- multi-rank correctness of the exchange, inside the warm-up (SURVEY 8e acceptance) ------------------
    exchange_check = None
    if world > 1:
        def seg_tensors(b):
            return list(zip(("xyz", "features_dc", "scaling", "rotation", "opacity"), b.views)) + [
                ("grad_norm_sum", b.grad_norm_sum), ("vis_count", b.vis_count), ("max_radii", b.max_radii)]

        view_loss = lambda out, vid: loss_fn(out, w_dev)   # noqa: E731
        mv.multiview_step(model, rd, [cam], settings, view_loss, buffer=buf, reduce=False)
        ref = buf.storage.clone()                                   # this rank's own contribution ...
        dist.all_reduce(ref[:buf.sum_elems], op=dist.ReduceOp.SUM)  # ... reduced by NCCL
        dist.all_reduce(ref[buf.sum_elems:], op=dist.ReduceOp.MAX)
        buf.all_reduce()                                            # ... and by the product's exchange
        got = buf.storage
        stat = torch.stack([(got - ref).abs().max().double(), ref.abs().max().double()])
        dist.all_reduce(stat, op=dist.ReduceOp.MAX)
        r0 = got.clone()
        dist.broadcast(r0, 0)
        same = torch.tensor([int(torch.equal(r0, got))], dtype=torch.int32, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        exchange_check = {"max_abs_err_vs_nccl": float(stat[0]), "max_rel_err_vs_nccl": float(stat[0] / stat[1]),
                          "bitwise_same_on_all_ranks": bool(int(same.item()))}
        if rank == 0:
            # the same `world` views rendered and summed on ONE GPU through the same fused accumulation
            single = mv.FlatGradBuffer(model, peer=False)
            cams_all = [gb.Camera.orbit(r, world, WIDTH, HEIGHT) for r in range(world)]
            mv.multiview_step(model, rd, cams_all, settings, view_loss, buffer=single, reduce=False)
            per_seg, mags = {}, {}
            for (name, a), (_, b) in zip(seg_tensors(buf), seg_tensors(single)):
                per_seg[name] = float((a - b).abs().max() / (b.abs().max() + 1e-30))
                mags[name] = float(b.abs().max())
            # the ref-init scene is isotropic: the covariance does not depend on the rotation, so its "gradient" is rounding
            # noise (~1e-11) whose relative error means nothing; it is reported as an absolute error next to its
            # magnitude and left out of the headline figure
            noise = [k for k in per_seg if k == "rotation" and mags[k] < 1e-6 * mags["xyz"]]
            exchange_check["max_rel_err_vs_1gpu_sum"] = max(v for k, v in per_seg.items() if k not in noise)
            exchange_check["rel_err_vs_1gpu_sum_per_segment"] = {k: v for k, v in per_seg.items() if k not in noise}
            if noise:
                a, b = dict(seg_tensors(buf))["rotation"], dict(seg_tensors(single))["rotation"]
                exchange_check["rotation_gradient_is_rounding_noise"] = {
                    "max_abs_1gpu": mags["rotation"], "max_abs_err": float((a - b).abs().max()), "max_abs_xyz_gradient": mags["xyz"]}
            exchange_check["what"] = (f"{world} orbit views: (a) each rank's local buffer reduced by the product's exchange vs NCCL all_reduce "
                                      "SUM/MAX of the same inputs; (b) the exchanged result vs the same views rendered and accumulated on rank 0 "
                                      "alone (max|a-b|/max|b| per buffer segment); (c) torch.equal of the full buffer across ranks")
            del single
        del ref, r0
        barrier()

    # ---- timed region (device-resident inputs) ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.gs_kernel_launch_count()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                   # L2 flush, outside the per-step event pair
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_device()
        b.record()
        evs.append((a, b))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.gs_kernel_launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = world * 1000.0 / ms_per_step

    # ---- e2e (host inputs, through the public API) --------------------------------------------------
    e2e_warm = 4
    for _ in range(e2e_warm):
        step_e2e()
    finish_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(args.steps, 60)        # a wall-clock region: long enough (~0.1 s) that one host hiccup does not decide it
    for _ in range(e2e_steps):
        step_e2e()
    finish_e2e()                                         # the last step's loss is read inside the timed region too
    barrier()
    assert len(e2e_losses) == e2e_steps + e2e_warm and all(np.isfinite(v) for v in e2e_losses)
    if trace:
        tr = trace[-e2e_steps:]
        enq = [b - a for a, b, c_, _, _ in tr]
        col = [c_ - b for a, b, c_, _, _ in tr]
        period = [tr[i + 1][0] - tr[i][0] for i in range(len(tr) - 1)]
        gpu = [ea.elapsed_time(eb) for _, _, _, ea, eb in tr]
        gap = [tr[i][4].elapsed_time(tr[i + 1][3]) for i in range(len(tr) - 1)]
        sys.stderr.write(f"[trace rank {rank}] per e2e step: host enqueue {np.mean(enq) * 1e3:.3f} ms (max {np.max(enq) * 1e3:.3f}), "
                         f"collect {np.mean(col) * 1e3:.3f} ms, host period {np.mean(period) * 1e3:.3f} ms; GPU first-to-last kernel "
                         f"{np.mean(gpu):.3f} ms (max {np.max(gpu):.3f}), GPU gap between steps {np.mean(gap):.3f} ms (max {np.max(gap):.3f})\n")
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_steps / float(t.item())
    clocks = sampler.stop() if rank == 0 else None      # sampled every 20 ms over the timed region and the e2e region
    h2d = sum(x.numel() * x.element_size() for x in w_host) + 64
    d2h = 4 + 24           # loss scalar + the (V, D, visible) counters of the frame

    # ---- per-kernel events (extra, untimed steps) -> roofline -------------------------------------
    timer = rmod.StageTimer()
    rmod.stage_timer.active = timer
    for _ in range(3):
        flush.zero_()
        step_device()
    rmod.stage_timer.active = None
    per = {k: float(np.mean(v)) for k, v in timer.summary_ms().items()}
    allreduce_ms = None
    if world > 1:                                   # the exchange step alone: SUM of the flat buffer + MAX of the radii
        ts = []
        for _ in range(5):
            mv.multiview_step(model, rd, [cam], settings, lambda out, vid: loss_fn(out, w_dev), buffer=buf, reduce=False)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            buf.all_reduce()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t_ar = torch.tensor([float(np.mean(ts[1:]))], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
        allreduce_ms = float(t_ar.item())
    st = rd.last_stats
    V, D = st["num_visible"], st["tile_pairs"]
    T = rd._last_debug["tile_ranges"].shape[0]
    E = int(rd._last_debug["tile_consumed"].sum().item())
    P = WIDTH * HEIGHT
    # algorithmic bytes per launch of what actually runs
    Vb = st["num_binned"]
    lens = (rd._last_debug["tile_ranges"][:, 1] - rd._last_debug["tile_ranges"][:, 0]).long()
    stored = int(lens.clamp(max=int(rd.list_cap)).sum().item()) if rd.list_cap else D       # entries the scatter stores
    chunks, row_tiles = -(-Vb // 255), -(-T // 128) * 128
    sort_passes = 3                                            # 8-bit passes over the differing key bits (24 here)
    alg = {
        "project_fwd": n * (56 + 109),                       # xyz 12+scale 12+quat 16+opacity 4+dc 12 ; outputs 109
        # depth sort: key range pass 4N; pass 1 reads keys, later passes (key,id); every pass but the last writes (key,id);
        # the last gathers tiles_touched and writes sorted ids 4N + offsets 8N
        "bin_prepare": n * (4 + 4 + 8 * (sort_passes - 1) + 8 * (sort_passes - 1) + 4 + 4 + 8),
        # counting sort on the tile id: ids+offsets+rectangles of the binned splats, twice (walk, scatter); one byte per
        # pair written and read (count-so-far); the [chunks x tiles] count (1 B) and prefix (2 B) tables written and read;
        # 4 B per stored list entry; tile ranges
        "bin_sort": Vb * (16 + 8) + 2 * D + 2 * 3 * chunks * row_tiles + 4 * stored + 8 * T,
        "raster_fwd": T * 12 + E * 52 + P * 40,
        "raster_bwd": T * 12 + E * 52 + E * 44 + P * 40,
        "project_bwd": n * (56 + 44 + 56 + 17),              # parameters, incoming gradients, outgoing gradients, statistics
    }
    if args.sh_degree > 0:                                   # the 15 higher-order rows: read forward and backward, gradient written
        alg["project_fwd"] += n * 180
        alg["project_bwd"] += n * 360
    hbm_peak, peak_src = measured_peaks()
    kernels = {k: {"ms": per[k], "alg_bytes": alg[k], "gbs": alg[k] / (per[k] * 1e-3) / 1e9,
                   "frac_of_hbm_peak": alg[k] / (per[k] * 1e-3) / 1e9 / hbm_peak, "share_of_step": per[k] / ms_per_step}
               for k in per if k in alg}
    top = max(kernels, key=lambda k: kernels[k]["ms"])
    evals = None
    if top in ("raster_fwd", "raster_bwd"):
        with torch.no_grad():                   # one untimed debug render: per-pixel walk lengths
            rd.render(cam, model, gb.RenderSettings(HEIGHT, WIDTH, torch.zeros(3, device=dev), debug=True))
        ncons = rd._last_debug["n_consumed"]
        evals = float(ncons.sum().item())      # pixel x list-entry evaluations actually walked
    traffic, ncu_static = None, {}
    try:                                         # DRAM bytes per launch from the committed ncu --set full capture
        ncu_static = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top, {})
        traffic = ncu_static["dram_read"] + ncu_static["dram_write"]
    except Exception:
        pass
    roofline = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[top]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                "note": ("the compositing kernels are FP32/MUFU-issue bound, not HBM bound (SURVEY 8d): see `issue`"
                         if evals else "")}
    frame_bytes = float(sum(alg.values()))
    roofline["frame"] = {"alg_bytes": frame_bytes, "achieved": frame_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
                         "frac": frame_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                         "note": "all stages' algorithmic bytes / step time: the whole frame against the HBM roofline"}
    if evals:
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        lane_roof = 148 * 128 * sm_clock * 1e6
        roofline["issue"] = {"pixel_splat_evals": evals, "evals_per_s": evals / (kernels[top]["ms"] * 1e-3),
                             "fp32_lane_instr_per_s_roof": lane_roof,
                             "lane_instr_per_eval_at_roof": lane_roof / (evals / (kernels[top]["ms"] * 1e-3)),
                             "ncu_fma_pipe_active_pct": ncu_static.get("fma_pipe_active_pct"),
                             "ncu_issue_active_pct": ncu_static.get("issue_active_pct")}
        slots = ncu_static.get("issue_slots_per_entry")
        if slots:
            # profiles/r2_issue_model.md: a packed FP32 instruction costs two issue slots, everything else one; the
            # kernel's ceiling is its slot count per list entry x the entries each of the 592 schedulers walks
            roofline["issue"].update({
                "issue_slots_per_list_entry": slots,
                "frac_of_issue_slot_roof": slots * E / (148 * 4 * kernels[top]["ms"] * 1e-3 * sm_clock * 1e6),
                "model": "slots = 2 x packed FP32 instructions + 1 x every other instruction (tools/ubench_issue.cu, "
                         "profiles/r2_issue_model.md); live: slots x consumed entries / (592 schedulers x kernel time x SM clock)"})

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                sec, cores, sample = cpu_port_frame_seconds()
                cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            except Exception as e:   # the GPU number must not be lost to a host-side problem
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_gpu_per_step": 1, "splats": n, "resolution": [WIDTH, HEIGHT],
                       "sh_degree": args.sh_degree,
                       "colour": ("sigmoid of the DC feature row, as the reference renders it (renderer.py:88-92: the model stores "
                                  "15 higher-order rows, all zero after create_from_random, and render() never reads them)"
                                  if args.sh_degree == 0 else
                                  f"view-dependent real-SH colour of degree {args.sh_degree} (extension; --sh-degree)"),
                       "visible": V, "tile_pairs_D": D, "consumed_entries_E": E,
                       "l2": "flushed before every timed step (256 MiB memset); per-step working set > L2",
                       "renderer": {"binning": "flat counting sort, optimistic sizes, tile lists truncated to their first "
                                               f"{rd.list_cap} entries with a completion path (result identical to complete lists)",
                                    "tile_order": f"longest first (forward: {rd.fwd_tile_order}" + (" = this camera's exact work at its previous visit" if rd.fwd_tile_order == "camera" else "") + "; backward: exact work)",
                                    "loss": "fused weighted sum over the five planes (gs_weighted_sum), gradient = the weights; "
                                            "weights are 8-bit fixed point on the host (the step's 10.4 MB of inputs), dequantised on the device"},
                       "parallelism": (f"view-sharded dp{world}, replicated Gaussians, gradient/statistics exchange of 17N floats: "
                                       + ("one peer-memory kernel per rank over NVLink (gs_peer_allreduce)" if buf.peer is not None
                                          else f"NCCL all_reduce (peer path unavailable: {buf.peer_error})")) if world > 1 else "single GPU"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kernels,
            "wall_frames_per_s": world * args.steps / t_wall,
        }
        if allreduce_ms is not None:
            line["allreduce"] = {"ms": allreduce_ms, "bytes": int(buf.flat.numel() * 4 + buf.max_radii.numel() * 4),
                                 "what": ("gs_peer_allreduce (SUM of the flat gradient/statistics buffer + MAX of the radii, one kernel)"
                                          if buf.peer is not None else "NCCL all_reduce(SUM) of the flat buffer + all_reduce(MAX) of the radii")}
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
