#!/usr/bin/env python
"""torchrun tool: times the pieces of FlatGradBuffer.all_reduce's peer path (barriers alone, kernel alone)."""
import json, os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import gsplat_b200 as gb
from importlib import import_module
_lib = import_module("mini-3d-gaussian-splatting_b200._lib")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m = gb.GaussianModel(device=dev); m.create_from_random(1_000_000, 1.0, seed=0)
buf = gb.multiview.FlatGradBuffer(m)
h = buf.peer["handle"]; lib = _lib.load()
st = lambda: ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def kernel(mc, flags=0):
    _lib.check(lib.gs_peer_allreduce(buf.peer["ptrs"], mc, world, rank, 0, buf.sum_elems, buf.sum_elems, buf.max_elems, flags, st()), "k")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
out = {"world": world}
out["two_barriers_ms"] = timeit(lambda: (h.barrier(channel=0), h.barrier(channel=1)))
out["kernel_p2p_ms"] = timeit(lambda: kernel(0))
out["kernel_tma_ms"] = timeit(lambda: kernel(0, 1))
mc_ptr = buf.peer["multicast"]
if not mc_ptr:
    try:
        mc_ptr = int(buf.peer["handle"].multicast_ptr or 0)
    except Exception:
        mc_ptr = 0
if mc_ptr:
    for unroll in (4, 8, 16):
        for mult in (2, 4):
            os.environ["GS_PEER_MC_UNROLL"], os.environ["GS_PEER_GRID_MULT"] = str(unroll), str(mult)
            out[f"kernel_multicast_u{unroll}_g{mult}_ms"] = timeit(lambda: kernel(mc_ptr))
    os.environ["GS_PEER_GRID_MULT"] = "2"
for mult in (1, 4, 8):
    os.environ["GS_PEER_GRID_MULT"] = str(mult)
    out[f"kernel_p2p_grid{mult}_ms"] = timeit(lambda: kernel(0))
os.environ["GS_PEER_GRID_MULT"] = "2"
out["lib"] = os.environ.get("GSPLAT_B200_LIB", "default")
out["full_ms"] = timeit(buf.all_reduce)
buf.peer["flags"] = 1
out["full_tma_ms"] = timeit(buf.all_reduce)
buf.peer["flags"] = 0
x = torch.empty(68_000_000 // 4, device=dev); y = torch.empty_like(x)
out["local_copy_68MB_ms"] = timeit(lambda: y.copy_(x))
if rank == 0: print(json.dumps(out), flush=True)
dist.destroy_process_group()
