"""CPU: host-side logic around the render path -- flat gradient buffer layout, view sharding, and the bench.py
reference arm's output contract (one JSON line on stdout)."""
import json
import os
import subprocess
import sys

import pytest
import torch

import gsplat_b200 as gb
from tests.conftest import ROOT

mv = gb.multiview


def test_flat_grad_buffer_layout_is_16_byte_aligned_for_any_n():
    for n in (1, 5, 1001, 4096):
        m = gb.GaussianModel(device="cpu")
        m.create_from_random(n, seed=0)
        buf = mv.FlatGradBuffer(m)
        assert buf.peer is None                                     # no process group: plain tensor, NCCL/gloo path
        base = buf.storage.data_ptr()
        for v, name in zip(buf.views, mv.PARAM_ORDER):
            assert (v.data_ptr() - base) % 16 == 0, (n, name)
            assert v.shape == getattr(m, name).shape
        for t in (buf.grad_norm_sum, buf.vis_count, buf.max_radii):
            assert (t.data_ptr() - base) % 16 == 0 and t.shape == (n,)
        assert buf.sum_elems % 4 == 0 and buf.max_elems % 4 == 0
        assert buf.storage.numel() == buf.sum_elems + buf.max_elems
        # the views tile the storage without overlap
        spans = sorted((v.data_ptr() - base, v.numel() * 4) for v in buf.views + [buf.grad_norm_sum, buf.vis_count, buf.max_radii])
        for (a, la), (b, _) in zip(spans, spans[1:]):
            assert a + la <= b


def test_install_zeroes_unless_the_first_view_will_overwrite():
    m = gb.GaussianModel(device="cpu")
    m.create_from_random(10, seed=0)
    buf = mv.FlatGradBuffer(m)
    buf.storage.fill_(3.0)
    buf.install()
    assert float(buf.storage.abs().max()) == 0.0 and buf.fresh is False
    assert m._xyz.grad.data_ptr() == buf.views[0].data_ptr() and m._features_rest.grad is None
    buf.storage.fill_(3.0)
    buf.install(zero=False)
    assert float(buf.storage.min()) == 3.0 and buf.fresh is True


def test_shard_views_partitions_every_view_exactly_once():
    for views in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            parts = [mv.shard_views(views, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(views))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert all(p == list(range(p[0], p[0] + len(p))) for p in parts if p)     # contiguous blocks


def test_bench_reference_arm_prints_exactly_one_json_line():
    import time
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports to its workers: the arm must not obey it
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"] == "fwd+bwd frames/s at 1080p, 1M Gaussians" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("config[1]")
    # a whole frame per step, nothing extrapolated: the timed steps fit inside the run; all host cores at any N
    assert d["steps"] == 1 and d["ms_per_step"] * d["steps"] * 1e-3 < wall
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "extrapolat" not in d["cpu_baseline"]["sample"].replace("nothing extrapolated", "")


def test_list_cap_policy_follows_the_deepest_walk_and_doubles_when_tiles_were_flagged():
    """Truncated tile lists (renderer.py next_list_cap): any cap is safe, the policy only trades stored entries against
    completed tiles."""
    from importlib import import_module
    nxt = import_module("mini-3d-gaussian-splatting_b200.renderer").next_list_cap
    assert nxt(1024, 0, 492, 0, True) == (576, 492)            # config[1]: 492 * 1.125 + 16 = 569.5 -> next multiple of 64
    assert nxt(576, 0, 492, 492, True) == (576, 492)           # stable
    cap, hi = nxt(576, 0, 100, 492, True)                      # an easy frame: the memory of the deep walk fades by 2 % only
    assert (cap, hi) == (576, 482)
    assert nxt(576, 3, 700, 492, True) == (1152, 700)          # tiles were flagged: at least double, whatever the target says
    assert nxt(128, 5, 2000, 0, True) == (2304, 2000)          # ... or the target when it is larger
    assert nxt(512, 0, 3, 0, True)[0] == 128                   # never below 128
    assert nxt(512, 0, -1, 77, True) == (512, 77)              # nothing measured, nothing flagged: unchanged
    assert nxt(512, 2, -1, 0, False) == (1024, 0)              # policy off: doubling only
    assert nxt(512, 0, 300, 0, False) == (512, 0)
    assert nxt(1 << 30, 1, -1, 0, False)[0] == 1 << 30         # bounded


def test_fused_loss_object_behaves_like_a_scalar_loss_for_the_step_drivers():
    """losses.FusedLoss (what the fused loss heads return) is consumed by multiview_step / train_step through
    backward() / detach() / item(): its backward must route the precomputed gradients into the graph of its inputs."""
    import torch
    from importlib import import_module
    FusedLoss = import_module("mini-3d-gaussian-splatting_b200.losses").FusedLoss
    p = torch.arange(6, dtype=torch.float32, requires_grad=True)
    img, alpha = (2.0 * p).reshape(2, 3), (p * p).sum().reshape(1)            # two "rendered outputs" of one graph
    g_img, g_alpha = torch.full((2, 3), 0.5), torch.tensor([3.0])
    loss = FusedLoss(torch.tensor(7.0), [img, alpha], [g_img, g_alpha])
    assert loss.item() == 7.0 and float(loss.detach()) == 7.0 and float(loss) == 7.0
    loss.backward()
    assert torch.allclose(p.grad, 2.0 * 0.5 + 3.0 * 2.0 * p.detach())
    # inputs that need no gradient are skipped, an empty loss is a no-op
    FusedLoss(torch.tensor(0.0), [torch.zeros(3)], [torch.ones(3)]).backward()
    with __import__("pytest").raises(RuntimeError):
        import_module("mini-3d-gaussian-splatting_b200.losses").l1_loss(torch.zeros(3), torch.zeros(3))     # CPU tensors: no fallback


def test_camera_block_of_the_renderer_equals_the_oracles_camera():
    """The 20 host floats gs_project_fwd receives (renderer._camera_block): W2C rotation and translation, intrinsics
    rounded once to fp32 as renderer.py:140-152 does, and the camera centre -R^T t -- against the oracle's camera (pinned
    to the literal reference by the stage fixtures) and the C port's block, for every camera the tests and bench use."""
    import numpy as np
    from importlib import import_module
    from oracle import c_port, splat_oracle as so
    rmod = import_module("mini-3d-gaussian-splatting_b200.renderer")
    cams = [so.camera_c0(1920, 1080), so.camera_c0(256, 256), so.camera_orbit(5, 16, 1920, 1080), so.camera_orbit(1, 8, 1600, 1200),
            so.camera_orbit(3, 7, 50, 44)]
    for cam in cams:
        block = np.array(list(rmod._camera_block(gb.Camera(cam.width, cam.height, cam.fovx, cam.fovy, world_view=cam.world_view))),
                         dtype=np.float32)
        c16 = c_port.camera_block(cam.width, cam.height, cam.fovx, cam.fovy, cam.world_view.numpy())
        assert np.array_equal(block[:16].view(np.uint32), c16.view(np.uint32))
        fx, fy, cx, cy = (float(v) for v in cam.intrinsics())
        assert [float(v) for v in block[12:16]] == [fx, fy, cx, cy]
        wv = cam.world_view.double()
        centre = -(wv[:3, :3].T @ wv[:3, 3])
        assert np.allclose(block[16:19], centre.numpy(), rtol=0, atol=1e-6)
        assert block[19] == 0.0


def test_ssim_loss_equals_an_independent_statement_of_the_references_formula():
    """src/core/loss.py:9-41 (which does not run) restated with scipy's 1-D correlation: same window, zero padding,
    constants and clamp; D-SSIM is 0 for identical images, differentiable, and GaussianLoss mixes it as (1-l) L1 + l D-SSIM."""
    import numpy as np
    scipy_ndimage = pytest.importorskip("scipy.ndimage")
    from importlib import import_module
    losses = import_module("mini-3d-gaussian-splatting_b200.losses")
    g = torch.Generator().manual_seed(3)
    a = torch.rand(3, 40, 52, generator=g)
    b = (a + 0.2 * torch.randn(3, 40, 52, generator=g)).clamp(0, 1)
    K = 11
    x = np.arange(K) - (K - 1) / 2
    w = np.exp(-x ** 2 / (2 * (K / 6) ** 2)); w /= w.sum()

    def blur(img):
        out = scipy_ndimage.correlate1d(img, w, axis=2, mode="constant", cval=0.0)
        return scipy_ndimage.correlate1d(out, w, axis=1, mode="constant", cval=0.0)

    A, B = a.double().numpy(), b.double().numpy()
    mx, my = blur(A), blur(B)
    sx, sy, sxy = blur(A * A) - mx * mx, blur(B * B) - my * my, blur(A * B) - mx * my
    ssim = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx ** 2 + my ** 2 + 1e-4) * (sx + sy + 9e-4))
    want = 1.0 - np.clip(ssim, 0, 1).mean()
    loss = losses.SSIMLoss()
    assert float(loss(a, b)) == pytest.approx(want, abs=2e-6)
    assert float(loss(a.unsqueeze(0), b.unsqueeze(0))) == pytest.approx(want, abs=2e-6)
    assert float(loss(a, a)) == pytest.approx(0.0, abs=1e-6)
    pred = b.clone().requires_grad_(True)
    total, parts = losses.GaussianLoss(0.2)(pred, a)
    total.backward()
    assert parts["total_loss"] == pytest.approx(0.8 * parts["l1"] + 0.2 * parts["dssim"], rel=1e-6)
    assert parts["dssim"] == pytest.approx(want, abs=2e-6) and float(pred.grad.abs().max()) > 0
