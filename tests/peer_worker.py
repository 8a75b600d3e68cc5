"""torchrun worker of tests/test_peer_exchange.py: the NVLink peer-memory gradient exchange (FlatGradBuffer.all_reduce ->
gs_peer_allreduce between two symmetric-memory barriers) against NCCL on the same inputs.  Prints one JSON line (rank 0)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class _Params:
    """Five parameter tensors shaped like GaussianModel's; n chosen so that the slices are ragged."""

    def __init__(self, n, dev):
        self._xyz = torch.zeros(n, 3, device=dev)
        self._features_dc = torch.zeros(n, 1, 3, device=dev)
        self._scaling = torch.zeros(n, 3, device=dev)
        self._rotation = torch.zeros(n, 4, device=dev)
        self._opacity = torch.zeros(n, 1, device=dev)


def main():
    n = int(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from importlib import import_module
    mv = import_module("mini-3d-gaussian-splatting_b200.multiview")
    buf = mv.FlatGradBuffer(_Params(n, dev), peer=True)            # peer=True: a set-up failure raises instead of falling back
    res = {"world": world, "n": n, "sum_elems": buf.sum_elems, "max_elems": buf.max_elems,
           "multicast": bool(buf.peer["multicast"]), "rounds": []}
    for rnd in range(3):
        g = torch.Generator(device=dev).manual_seed(1000 * rnd + rank)
        buf.storage.copy_(torch.randn(buf.storage.numel(), generator=g, device=dev) * (10.0 ** (rnd - 1)))
        buf.storage[buf.sum_elems:].abs_()                          # the MAX region holds screen radii: non-negative
        if rnd == 2:
            buf.storage[buf.sum_elems:] *= (torch.rand(buf.max_elems, generator=g, device=dev) < 0.5)   # zeros where not visible
        ref = buf.storage.clone()
        dist.all_reduce(ref[:buf.sum_elems], op=dist.ReduceOp.SUM)
        dist.all_reduce(ref[buf.sum_elems:], op=dist.ReduceOp.MAX)
        buf.fresh = False
        buf.all_reduce()
        torch.cuda.synchronize()
        got = buf.storage
        scale = float(ref[:buf.sum_elems].abs().max())
        sum_err = float((got[:buf.sum_elems] - ref[:buf.sum_elems]).abs().max()) / scale
        max_equal = bool(torch.equal(got[buf.sum_elems:], ref[buf.sum_elems:]))      # MAX is exact in any order
        r0 = got.clone()
        dist.broadcast(r0, 0)
        flags = torch.tensor([int(torch.equal(r0, got)), int(max_equal)], dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        err = torch.tensor([sum_err], dtype=torch.float64, device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        res["rounds"].append({"sum_rel_err_vs_nccl": float(err), "bitwise_same_on_all_ranks": bool(int(flags[0])),
                              "max_region_equals_nccl": bool(int(flags[1]))})
    if rank == 0:
        print("PEER_RESULT " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
