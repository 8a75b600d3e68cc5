#!/usr/bin/env python
"""Times gs_bin_sort alone (both algorithms) on the bench scene; diagnostic for the binning kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gsplat_b200 as gb
from importlib import import_module
_lib = import_module("mini-3d-gaussian-splatting_b200._lib")
ptr, check = _lib.ptr, _lib.check
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W, H = 1920, 1080
dev = torch.device("cuda", 0)
model = gb.GaussianModel(device=dev); model.create_from_random(N, 1.0, seed=0)
rd = gb.GaussianRenderer()
st = gb.RenderSettings(H, W, torch.zeros(3, device=dev))
cam = gb.Camera.look_at_origin_c0(W, H)
with torch.no_grad():
    rd.render(cam, model, st)
dbg = rd._last_debug
stats = rd.last_stats
num_sorted, D = stats["num_binned"], stats["tile_pairs"]
tiles_x, tiles = (W + 15) // 16, ((W + 15) // 16) * ((H + 15) // 16)
sorted_ids = dbg["sorted_ids"].contiguous()
# offsets are recomputed exactly as gs_bin_prepare does
tt = dbg["tiles_touched"][sorted_ids.long()].to(torch.int64)
offsets = torch.cumsum(tt, 0) - tt
ws_bytes = int(lib.gs_bin_workspace_bytes(num_sorted, D, tiles))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
entry = torch.empty(D, dtype=torch.int32, device=dev)
ranges = torch.empty((tiles, 2), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
import ctypes
ref = None
for algo in (1, 2, 3):
    ts = []
    for r in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib.gs_bin_sort(N, num_sorted, D, ptr(sorted_ids), ptr(offsets), ptr(dbg["tile_rect"]), ptr(dbg["depth_keys"]),
                              tiles_x, tiles, algo, ptr(ws), ws.numel(), ptr(entry), ptr(ranges), None, None, 0, None, None, ctypes.c_void_p(stream)), "bin")
        b.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(a.elapsed_time(b))
    if ref is None: ref = (entry.clone(), ranges.clone())
    else: print("algos agree:", torch.equal(ref[0], entry), torch.equal(ref[1], ranges))
    print(f"algo {algo}: N={N} V={num_sorted} D={D}  {sum(ts)/len(ts)*1000:.1f} us  (min {min(ts)*1000:.1f})")
