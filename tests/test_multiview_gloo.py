"""CPU, world_size 2, gloo: the host logic of the view-sharded step -- shard partition, gradients
accumulating in place into the flat buffer, SUM/MAX all-reduce -- once with a small differentiable
stand-in for the renderer and once with the CPU oracle of the render path on real orbit cameras (the
CUDA renderer itself is covered by the gpu tests and by bench.py's exchange_check).  The acceptance
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


class OracleRenderer:
    """The CPU oracle of the render path (oracle/splat_oracle.py, pinned to the literal reference) behind the renderer's
    call signature: the view-sharded step then reduces the gradients of the REAL render math, on real orbit cameras."""
    W, H = 48, 40

    def render(self, camera, model, settings):
        from oracle import splat_oracle as so
        cam = so.camera_orbit(int(camera), N_VIEWS, self.W, self.H)
        return so.render_from_params(cam, model._xyz, model._scaling, model._rotation, model._opacity, model._features_dc,
                                     torch.tensor([0.1, 0.2, 0.3]), self.H, self.W)


def make_model(oracle_scene: bool = False):
    m = gb.GaussianModel(device="cpu")
    if oracle_scene:
        import math
        from oracle import splat_oracle as so
        s = so.scene_aniso(120, 3)
        m.create_from_tensors(s["xyz"], s["features_dc"], s["scaling"] + math.log(8.0), s["rotation"], s["opacity"])
    else:
        m.create_from_random(N_SPLATS, seed=3)
    return m


def loss_fn(out, vid):
    return (out["image"] * (vid + 1)).sum() + out["alpha"].sum() + 0.1 * out["depth"].sum()


def run_single(oracle: bool = False):
    m = make_model(oracle)
    rd = OracleRenderer() if oracle else ToyRenderer()
    res = mv.multiview_step(m, rd, list(range(N_VIEWS)), None, loss_fn, view_ids=list(range(N_VIEWS)), reduce=False)
    return res["buffer"]


def worker(rank, world, port, q, oracle=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = make_model(oracle)
    ids = mv.shard_views(N_VIEWS, rank, world)
    buf = mv.FlatGradBuffer(m)
    flat_ptr = buf.flat.data_ptr()
    mv.multiview_step(m, OracleRenderer() if oracle else ToyRenderer(), ids, None, loss_fn, view_ids=ids, buffer=buf)
    # gradients must have accumulated in place into the flat buffer (no packing step)
    assert m._xyz.grad.data_ptr() == flat_ptr
    assert buf.flat.data_ptr() == flat_ptr
    # by value (numpy), not as torch tensors: a tensor travels as a handle to the sender's shared memory, and the sender
    # may have exited by the time the parent opens it
    q.put((rank, buf.flat.numpy().copy(), buf.max_radii.numpy().copy(), ids))
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


@pytest.mark.parametrize("oracle", [False, True], ids=["toy-renderer", "oracle-renderer"])
def test_two_rank_reduced_grads_equal_single_rank_sum(oracle):
    ref = run_single(oracle)
    if oracle:      # the scene is seen: every parameter group receives gradient, some splats are culled in some views
        assert all(float(v.abs().max()) > 0 for v in ref.views)
        assert float(ref.vis_count.max()) == N_VIEWS and float(ref.max_radii.max()) > 1.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q, oracle)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    got = [(r, torch.from_numpy(f), torch.from_numpy(m), i) for r, f, m, i in got]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = sorted(sum((g[3] for g in got), []))
    assert ids == list(range(N_VIEWS))
    for rank, flat, max_radii, _ in got:
        scale = float(ref.flat.abs().max())
        assert torch.allclose(flat, ref.flat, rtol=1e-5, atol=1e-6 * max(1.0, scale)), f"rank {rank}"
        assert torch.equal(max_radii, ref.max_radii)
    assert torch.equal(got[0][1], got[1][1])       # every rank ends with identical reduced buffers
