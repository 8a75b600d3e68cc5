"""CPU, world_size 2, gloo: the host logic of the view-sharded step -- shard partition, gradients
accumulating in place into the flat buffer, SUM/MAX all-reduce -- with a small differentiable
stand-in for the renderer (the CUDA renderer itself is covered by the gpu tests).  The acceptance
criterion is SURVEY 8e's: N-rank reduced gradients == 1-rank sum over the same views."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gsplat_b200 as gb

mv = gb.multiview
N_SPLATS, N_VIEWS = 64, 5


class ToyRenderer:
    """Differentiable function of every parameter with the renderer's output contract."""

    def render(self, camera, model, settings):
        k = float(camera)
        means = model._xyz[:, :2] * (1.0 + 0.1 * k) + model._rotation[:, :2] * 0.01
        col = torch.sigmoid(model._features_dc[:, 0, :]) * torch.sigmoid(model._opacity)
        img = (means.sum(-1, keepdim=True) * col * torch.exp(model._scaling).mean(-1, keepdim=True)).t().reshape(3, 8, 8)
        vis = (model._xyz[:, 2] + 0.2 * k) > 0
        return {"image": img, "alpha": img[:1] * 0.5, "depth": img[1:2] * 2.0, "viewspace_points": means,
                "visibility_filter": vis, "radii": model._xyz[:, 0].detach().abs() * (k + 1), "conics": None}


def make_model():
    m = gb.GaussianModel(device="cpu")
    m.create_from_random(N_SPLATS, seed=3)
    return m


def loss_fn(out, vid):
    return (out["image"] * (vid + 1)).sum() + out["alpha"].sum() + 0.1 * out["depth"].sum()


def run_single():
    m = make_model()
    res = mv.multiview_step(m, ToyRenderer(), list(range(N_VIEWS)), None, loss_fn, view_ids=list(range(N_VIEWS)), reduce=False)
    return res["buffer"]


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = make_model()
    ids = mv.shard_views(N_VIEWS, rank, world)
    buf = mv.FlatGradBuffer(m)
    flat_ptr = buf.flat.data_ptr()
    mv.multiview_step(m, ToyRenderer(), ids, None, loss_fn, view_ids=ids, buffer=buf)
    # gradients must have accumulated in place into the flat buffer (no packing step)
    assert m._xyz.grad.data_ptr() == flat_ptr
    assert buf.flat.data_ptr() == flat_ptr
    q.put((rank, buf.flat.clone(), buf.max_radii.clone(), ids))
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_views_partitions_all_views():
    for n in (1, 5, 8, 13):
        for w in (1, 2, 3, 8):
            parts = [mv.shard_views(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_reduced_grads_equal_single_rank_sum():
    ref = run_single()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = sorted(sum((g[3] for g in got), []))
    assert ids == list(range(N_VIEWS))
    for rank, flat, max_radii, _ in got:
        assert torch.allclose(flat, ref.flat, rtol=1e-5, atol=1e-6), f"rank {rank}"
        assert torch.equal(max_radii, ref.max_radii)
    assert torch.equal(got[0][1], got[1][1])       # every rank ends with identical reduced buffers
