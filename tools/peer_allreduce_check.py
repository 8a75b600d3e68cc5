#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/peer_allreduce_check.py : gs_peer_allreduce against NCCL on the flat
gradient buffer of a 1 M-splat model (correctness, then CUDA-event timing of both).  Rank 0 prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import gsplat_b200 as gb
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = gb.GaussianModel(device=dev); m.create_from_random(n, 1.0, seed=0)
buf = gb.multiview.FlatGradBuffer(m)
out = {"world": world, "n": n, "peer": buf.peer is not None, "peer_error": buf.peer_error,
       "multicast": bool(buf.peer and buf.peer["multicast"])}
if buf.peer is not None:
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    data = torch.randn(buf.storage.numel(), generator=g, device=dev)
    data[buf.sum_elems:].abs_()
    ref_sum, ref_max = data[:buf.sum_elems].clone(), data[buf.sum_elems:].clone()
    dist.all_reduce(ref_sum, op=dist.ReduceOp.SUM); dist.all_reduce(ref_max, op=dist.ReduceOp.MAX)
    buf.storage.copy_(data)
    torch.cuda.synchronize(); dist.barrier()
    buf.all_reduce()
    torch.cuda.synchronize()
    out["max_abs_err_sum"] = float((buf.flat - ref_sum).abs().max())
    out["max_equal"] = bool(torch.equal(buf.storage[buf.sum_elems:], ref_max))
    # identical on all ranks?
    chk = buf.storage.double().sum().reshape(1)
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    out["bitwise_same_on_all_ranks"] = all(float(c) == float(allchk[0]) for c in allchk)

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
    nccl_flat, nccl_max = torch.zeros_like(ref_sum), torch.zeros_like(ref_max)
    def nccl():
        dist.all_reduce(nccl_flat, op=dist.ReduceOp.SUM); dist.all_reduce(nccl_max, op=dist.ReduceOp.MAX)
    out["peer_ms"] = timeit(buf.all_reduce)
    out["nccl_ms"] = timeit(nccl)
    out["bytes"] = buf.storage.numel() * 4
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
