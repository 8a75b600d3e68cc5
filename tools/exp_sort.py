#!/usr/bin/env python
"""Experiment: gs_bin_prepare (depth sort) alone on the bench scene's keys: CUDA-event time and, with a library built
with -DGS_SORT_TIMING=1, the %globaltimer stamps of CTA 0 at every phase boundary."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gsplat_b200 as gb
from importlib import import_module
_lib = import_module("mini-3d-gaussian-splatting_b200._lib")
lib = _lib.load()
W, H = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = gb.GaussianModel(device="cuda"); m.create_from_random(n, 1.0, seed=0)
rd = gb.GaussianRenderer()
with torch.no_grad():
    rd.render(gb.Camera.look_at_origin_c0(W, H), m, gb.RenderSettings(H, W, torch.zeros(3, device="cuda")))
keys, tt = rd._last_debug["depth_keys"], rd._last_debug["tiles_touched"]
P = _lib.ptr
counters = torch.empty(3, dtype=torch.int64, device="cuda")
sorted_ids = torch.empty(n, dtype=torch.int32, device="cuda")
offsets = torch.empty(n, dtype=torch.int64, device="cuda")
wsb = int(lib.gs_bin_workspace_bytes(n, 0, 8160))
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    _lib.check(lib.gs_bin_prepare(n, P(keys), P(tt), P(ws), wsb, P(sorted_ids), P(offsets), P(counters), st), "prep")
res = {"lib": os.environ.get("GSPLAT_B200_LIB", "default"), "n": n}
for mode in ("warm", "flushed"):
    ts = []
    for i in range(12):
        if mode == "flushed": flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    res[mode + "_us"] = [round(float(np.median(ts[2:])), 1), round(float(np.min(ts[2:])), 1)]
up = lambda x: (x + 255) // 256 * 256
off = 6 * up(n * 4) + up(1024 * 256 * 4) + up(1024 * 256 * 8) + 16384
stamps = ws[off:off + 48 * 8].view(torch.int64).cpu().numpy()
if stamps[0] > 0:
    names = ["start", "p0 done", "sync"] + [f"pass{p} {x}" for p in range(4) for x in ("A start", "A done", "sync", "scan done", "B done")] + ["end"] + ["-"] * 8 + [f"pass{p} B {x}" for p in range(4) for x in ("ranked", "reordered", "wscanned", "-")]
    t0 = stamps[0]
    res["stamps_us"] = {nm: round((int(s) - int(t0)) / 1e3, 2) for nm, s in zip(names, stamps) if s > 0}
print(json.dumps(res))
