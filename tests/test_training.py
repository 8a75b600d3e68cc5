"""The caller side (SURVEY 8f "next"): learning-rate schedule, densify schedule and split/clone/prune
bookkeeping on CPU tensors; a short optimisation run and a densification stress round on the GPU."""
import math
import os

import numpy as np

import pytest
import torch

import gsplat_b200 as gb
from oracle import splat_oracle as so
from tests import util


def test_learning_rate_schedule_matches_reference_formula():
    s = gb.LearningRateScheduler(1.6e-4, 1.6e-6, 300, 0.01, 30000)
    assert s.get_lr(0) == pytest.approx(1.6e-4 * 0.01)
    assert s.get_lr(300) == pytest.approx(1.6e-6 + (1.6e-4 - 1.6e-6) * 0.5 * (1 + math.cos(math.pi * 0.01)))
    assert s.get_lr(30000) == pytest.approx(1.6e-6)
    assert s.get_lr(10 ** 9) == pytest.approx(1.6e-6)
    assert gb.LearningRateScheduler(1.0, 0.5, 0, 1.0, 0).get_lr(5) == 0.5


def test_training_config_is_the_references_field_for_field_and_round_trips_through_yaml(tmp_path):
    import dataclasses
    cfg = gb.ConfigManager.get_default_config()
    path = str(tmp_path / "sub" / "config.yaml")
    changed = dataclasses.replace(cfg, iterations=123, densify_grad_threshold=5e-4, device="cpu", output_path="out/x")
    gb.ConfigManager.save_to_yaml(changed, path)
    assert gb.ConfigManager.load_from_yaml(path) == changed
    try:                                       # the literal reference dataclass, when the reference tree is present
        import contextlib, io, sys
        sys.path.insert(0, "/root/reference")
        with contextlib.redirect_stdout(io.StringIO()):
            from config.config import TrainingConfig as RefConfig
        assert dataclasses.asdict(RefConfig()) == dataclasses.asdict(cfg)
        assert [f.name for f in dataclasses.fields(RefConfig)] == [f.name for f in dataclasses.fields(cfg)]
    except ImportError:
        pass
    finally:
        if "/root/reference" in sys.path:
            sys.path.remove("/root/reference")


def test_densify_schedule():
    c = gb.DensityController(gb.TrainingConfig())
    assert not c.should_densify(400) and c.should_densify(500) and not c.should_densify(550)
    assert c.should_densify(15000) and not c.should_densify(15100)


def test_split_clone_prune_counts_cpu():
    """Pins the working subset of the reference's densification (tests/test_gaussian_model.py:91-140):
    a split removes the parent and adds two children, a clone adds one, pruning keeps the mask."""
    m = gb.GaussianModel(device="cpu")
    m.create_from_random(100, 1.0, seed=1)
    with torch.no_grad():
        m._scaling[:30] = math.log(0.05)      # large  -> split candidates
        m._scaling[30:50] = math.log(0.005)   # small  -> clone candidates
        m._opacity[90:] = -10.0               # transparent -> pruned
    g = torch.zeros(100, 3)
    g[:10] = 1.0                              # 10 large with gradient
    g[30:35] = 1.0                            # 5 small with gradient
    g[60:70] = 1.0                            # mid-sized: neither
    ctrl = gb.DensityController(gb.TrainingConfig())
    stats = ctrl.densify_and_prune(m, None, 1.0, grad=g, generator=torch.Generator().manual_seed(0))
    assert stats["split"] == 10 and stats["cloned"] == 5 and stats["pruned"] == 10
    assert m.get_num_points() == 100 + 5 + 10 - 10
    assert m._features_rest.shape == (m.get_num_points(), 15, 3) and m.denom.shape == (m.get_num_points(), 1)
    # children of a split are 0.75x the parent's size
    assert torch.allclose(m.get_scaling[-20:], torch.full((20, 3), 0.05 * 0.75), rtol=1e-5)


@pytest.mark.gpu
def test_train_steps_reduce_l1_loss():
    W, H = 160, 96
    target_scene = so.scene_aniso(3000, 50)
    target_scene["scaling"] = target_scene["scaling"] + math.log(2.5)
    tm = gb.GaussianModel(device="cuda")
    tm.create_from_tensors(target_scene["xyz"], target_scene["features_dc"], target_scene["scaling"],
                           target_scene["rotation"], target_scene["opacity"])
    cam = gb.Camera.orbit(1, 7, W, H)
    rd = gb.GaussianRenderer()
    st = gb.RenderSettings(H, W, torch.zeros(3, device="cuda"))
    with torch.no_grad():
        target = rd.render(cam, tm, st)["image"].clone()
    m = gb.GaussianModel(device="cuda")
    m.create_from_tensors(target_scene["xyz"] + 0.01 * torch.randn(3000, 3, generator=torch.Generator().manual_seed(2)),
                          target_scene["features_dc"] * 0.5, target_scene["scaling"], target_scene["rotation"],
                          target_scene["opacity"])
    cfg = gb.TrainingConfig(position_lr_init=1e-3, position_lr_final=1e-4, position_lr_max_steps=60)
    opt = gb.GaussianOptimizer(m, cfg)
    losses = [float(gb.train_step(m, rd, cam, target, opt, st, it)["loss"]) for it in range(40)]
    assert losses[-1] < 0.6 * losses[0], losses[::8]
    assert float(m.denom.max()) == 40 and float(m.max_radii2D.max()) > 0
    assert all(math.isfinite(v) for v in losses)


@pytest.mark.gpu
def test_densification_stress_grows_model():
    """Scaled-down BASELINE config[4]: repeated render -> backward -> split/clone/prune rounds."""
    s = so.scene_aniso(20000, 9)
    m = gb.GaussianModel(device="cuda")
    m.create_from_tensors(s["xyz"], s["features_dc"], s["scaling"], s["rotation"], s["opacity"])
    rd = gb.GaussianRenderer()
    W, H = 400, 304
    cams = [gb.Camera.orbit(k, 4, W, H) for k in range(4)]
    cfg = gb.TrainingConfig(densify_grad_threshold=1e-9)
    res = gb.training.densification_stress(m, rd, cams, gb.RenderSettings(H, W, torch.zeros(3, device="cuda")),
                                           gb.DensityController(cfg), 1.0, target_points=60000, max_rounds=12)
    assert res["points"] >= 40000 and len(res["history"]) >= 1
    n = m.get_num_points()
    assert m._xyz.shape == (n, 3) and m._rotation.shape == (n, 4) and m.max_radii2D.shape == (n,)
    with torch.no_grad():
        out = rd.render(cams[0], m, gb.RenderSettings(H, W, torch.zeros(3, device="cuda")))
    assert bool(torch.isfinite(out["image"]).all()) and out["radii"].shape == (n,)


@pytest.mark.gpu
def test_fused_gradient_sink_equals_autograd_accumulation():
    """multiview_step on the B200 renderer adds parameter gradients and densification statistics inside
    gs_project_bwd (accumulate + stat_*); it must equal autograd's `.grad +=` plus add_view_stats."""
    mv = gb.multiview
    s = so.scene_aniso(4001, 31)                 # odd count: float4 accumulations stay aligned
    s["scaling"] = s["scaling"] + math.log(2.5)
    W, H = 176, 112
    cams = [gb.Camera.orbit(k, 5, W, H) for k in (0, 2, 3)]
    st = gb.RenderSettings(H, W, torch.tensor([0.1, 0.0, 0.2], device="cuda"))
    wts = [tuple(t.cuda() for t in so.loss_weights(H, W, seed=10 + k)) for k in range(3)]

    def loss_fn(out, vid):
        return so.weighted_loss(out, wts[vid])

    def run(fused: bool):
        m = gb.GaussianModel(device="cuda")
        m.create_from_tensors(s["xyz"], s["features_dc"], s["scaling"], s["rotation"], s["opacity"])
        rd = gb.GaussianRenderer()
        res = mv.multiview_step(m, rd if fused else _NoSink(rd), cams, st, loss_fn, reduce=False)
        return res["buffer"]

    class _NoSink:
        """The same renderer without the sink entry point."""
        def __init__(self, rd):
            self._rd = rd

        def render(self, *a):
            return self._rd.render(*a)

    a, b = run(True), run(False)
    assert float(b.flat.abs().max()) > 0
    scale = float(b.flat[:b.param_elems].abs().max())
    assert float((a.flat[:a.param_elems] - b.flat[:b.param_elems]).abs().max()) <= 2e-6 * scale
    assert torch.allclose(a.grad_norm_sum, b.grad_norm_sum, rtol=1e-5, atol=1e-9 * scale)
    assert torch.equal(a.vis_count, b.vis_count) and float(a.vis_count.max()) == 3.0
    assert torch.equal(a.max_radii, b.max_radii) and float(a.max_radii.max()) > 0


@pytest.mark.gpu
def test_device_densification_equals_sequential_formulation():
    """gs_densify_plan / gs_densify_apply against density_and_clone -> density_and_split -> prune (tensor ops)."""
    n = 20011
    s = so.scene_aniso(n, 61)
    g = torch.Generator().manual_seed(7)
    sc = s["scaling"].clone()
    sc[: n // 3] = math.log(0.05) + 0.1 * torch.randn(n // 3, 3, generator=g)          # large  -> split candidates
    sc[n // 3: 2 * n // 3] = math.log(0.004) + 0.1 * torch.randn(n // 3, 3, generator=g)   # small  -> clone candidates
    op = s["opacity"].clone()
    op[::7] = -9.0                                         # transparent -> pruned (also as clone copies / children)
    op[5::11] = 8.0                                        # children's logit is clamped to 6
    grad = torch.zeros(n, 3)
    hot = torch.rand(n, generator=g) < 0.6
    grad[hot] = 1.0 + torch.rand(int(hot.sum()), 3, generator=g)
    rest = 0.1 * torch.randn(n, 15, 3, generator=g)

    def fresh():
        m = gb.GaussianModel(device="cuda")
        m.create_from_tensors(s["xyz"], s["features_dc"], sc, s["rotation"], op, rest)
        return m
    th, extent = 0.5, 1.0
    a, b = fresh(), fresh()
    gc = grad.cuda()
    # the sequential path draws randn(k,3) for its k clone candidates; the fused path draws the same block from
    # the same generator state (after its plan pass), so the two place the clones identically
    cfg = gb.TrainingConfig(densify_grad_threshold=th)
    ra = gb.DensityController(cfg, fused=False).densify_and_prune(a, None, extent, grad=gc, generator=torch.Generator().manual_seed(3))
    rb = b.densify_fused(gc, th, extent, 0.01, generator=torch.Generator().manual_seed(3))
    assert rb["points"] == ra["points"] == a.get_num_points() == b.get_num_points()
    assert rb["cloned"] > 1000 and rb["split"] > 1000 and rb["points"] != n
    for name in ("_features_dc", "_features_rest"):
        assert torch.equal(getattr(a, name).data, getattr(b, name).data), name
    for name, tol in (("_xyz", 1e-6), ("_scaling", 1e-6), ("_rotation", 1e-6), ("_opacity", 2e-5)):
        assert torch.allclose(getattr(a, name).data, getattr(b, name).data, rtol=0, atol=tol), name
    assert b.denom.shape == (rb["points"], 1) and b.max_radii2D.shape == (rb["points"],)
    # the controller takes the fused route on CUDA models by default
    c = fresh()
    rc = gb.DensityController(cfg).densify_and_prune(c, None, extent, grad=gc, generator=torch.Generator().manual_seed(3))
    assert rc["points"] == c.get_num_points() and rc["split"] == rb["split"] and rc["cloned"] == rb["cloned"]


# ----------------------------------------------------------------------------------------------
# density control pinned to the reference's own functions (tests/golden/densify_n1500.npz: gaussian_model.py:130-197
# run by tests/golden/make_golden.py with the `_append_points` patch of the reference's own test)
# ----------------------------------------------------------------------------------------------
DENSIFY_KEYS = (("_xyz", "xyz"), ("_features_dc", "features_dc"), ("_features_rest", "features_rest"), ("_scaling", "scaling"),
                ("_rotation", "rotation"), ("_opacity", "opacity"))


def _densify_fixture_model(d, device):
    m = gb.GaussianModel(device=device)
    m.create_from_tensors(*(torch.tensor(d["in_" + k]) for k in ("xyz", "features_dc", "scaling", "rotation", "opacity", "features_rest")))
    return m


def test_sequential_density_control_equals_the_reference_functions_row_for_row():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify_n1500.npz"))
    m = _densify_fixture_model(d, "cpu")
    cfg = gb.TrainingConfig(densify_grad_threshold=float(d["th"]))
    r = gb.DensityController(cfg, fused=False).densify_and_prune(m, None, float(d["extent"]), grad=torch.tensor(d["in_grad"]),
                                                                 generator=torch.Generator().manual_seed(3))
    assert r["cloned"] == int(d["n_clone_candidates"]) and r["split"] == int(d["n_split_candidates"])
    for attr, key in DENSIFY_KEYS:
        got, want = getattr(m, attr).data, torch.tensor(d["ref_" + key])
        assert got.shape == want.shape, (attr, got.shape, want.shape)
        assert torch.equal(got, want), (attr, float((got - want).abs().max()))


@pytest.mark.gpu
def test_device_density_control_equals_the_reference_functions_row_for_row():
    """gs_densify_plan / gs_densify_apply against the literal reference output: same rows, same order."""
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify_n1500.npz"))
    m = _densify_fixture_model(d, "cuda")
    r = m.densify_fused(torch.tensor(d["in_grad"]).cuda(), float(d["th"]), float(d["extent"]), float(d["min_opacity"]),
                        noise=torch.tensor(d["noise"]).cuda())
    assert r["points"] == d["ref_xyz"].shape[0]
    for attr, key in DENSIFY_KEYS:
        got, want = getattr(m, attr).data.cpu(), torch.tensor(d["ref_" + key])
        assert got.shape == want.shape, (attr, got.shape, want.shape)
        if attr in ("_features_dc", "_features_rest"):
            assert torch.equal(got, want), attr                          # copied rows
        else:
            tol = 2e-5 if attr == "_opacity" else 1e-6                   # expf / logf of the device vs the host's
            assert torch.allclose(got, want, rtol=0, atol=tol), (attr, float((got - want).abs().max()))
    # the same call with the generator the sequential path would use draws the same jitter block
    m2 = _densify_fixture_model(d, "cuda")
    m2.densify_fused(torch.tensor(d["in_grad"]).cuda(), float(d["th"]), float(d["extent"]), float(d["min_opacity"]),
                     generator=torch.Generator().manual_seed(3))
    assert torch.allclose(m2._xyz.data.cpu(), torch.tensor(d["ref_xyz"]), rtol=0, atol=1e-6)


# ----------------------------------------------------------------------------------------------
# optimiser state across a densification round and the train step against the oracle (rank 2)
# ----------------------------------------------------------------------------------------------
def _mixed_model(n, device, seed=5):
    s = so.scene_aniso(n, seed)
    g = torch.Generator().manual_seed(seed + 100)
    sc = s["scaling"].clone()
    sc[: n // 3] = math.log(0.05) + 0.1 * torch.randn(n // 3, 3, generator=g)
    sc[n // 3: 2 * n // 3] = math.log(0.004) + 0.1 * torch.randn(n // 3, 3, generator=g)
    op = s["opacity"].clone()
    op[::9] = -9.0
    m = gb.GaussianModel(device=device)
    m.create_from_tensors(s["xyz"], s["features_dc"], sc, s["rotation"], op)
    return m, g


def _adam_state(opt):
    out = {}
    for grp in opt.optimizer.param_groups:
        st = opt.optimizer.state[grp["params"][0]]
        out[grp["name"]] = (st["exp_avg"].clone(), st["exp_avg_sq"].clone(), float(st["step"]), grp["lr"])
    return out


def _check_carry(m, opt, before, src_row):
    src_row = src_row.long().cpu()
    kept = src_row >= 0
    assert int(kept.sum()) > 0 and int((~kept).sum()) > 0
    for grp in opt.optimizer.param_groups:
        st = opt.optimizer.state[grp["params"][0]]
        avg0, sq0, step0, lr0 = before[grp["name"]]
        assert st["exp_avg"].shape == grp["params"][0].shape
        assert torch.equal(st["exp_avg"].cpu()[kept], avg0.cpu()[src_row[kept]]), grp["name"]
        assert torch.equal(st["exp_avg_sq"].cpu()[kept], sq0.cpu()[src_row[kept]]), grp["name"]
        assert float(st["exp_avg"].cpu()[~kept].abs().max()) == 0.0 and float(st["exp_avg_sq"].cpu()[~kept].abs().max()) == 0.0
        assert float(st["step"]) == step0 and grp["lr"] == lr0


def _run_carry(device, fused):
    m, g = _mixed_model(900, device)
    cfg = gb.TrainingConfig(densify_grad_threshold=0.5)
    opt = gb.GaussianOptimizer(m, cfg)
    for it in range(3):
        opt.update_learning_rate(it)
        for name in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity"):
            p = getattr(m, name)
            p.grad = torch.randn(p.shape, generator=g).to(device)
        opt.step()
    before = _adam_state(opt)
    params_before = m._xyz.data.clone()
    grad = torch.zeros(900, 3)
    hot = torch.rand(900, generator=g) < 0.6
    grad[hot] = 1.0 + torch.rand(int(hot.sum()), 3, generator=g)
    r = gb.DensityController(cfg, fused=fused).densify_and_prune(m, opt, 1.0, grad=grad.to(device), generator=torch.Generator().manual_seed(1))
    src = r["src_row"]
    assert src.shape[0] == m.get_num_points() == r["points"]
    keep = src.long().cpu() >= 0
    assert torch.equal(m._xyz.data.cpu()[keep], params_before.cpu()[src.long().cpu()[keep]])     # rows really are where src_row says
    _check_carry(m, opt, before, src)
    for name in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity"):                 # and the next step runs on the new rows
        p = getattr(m, name)
        p.grad = torch.randn(p.shape, generator=g).to(device)
    opt.step()
    assert all(float(opt.optimizer.state[grp["params"][0]]["step"]) == 4.0 for grp in opt.optimizer.param_groups)
    return src.long().cpu()


def test_adam_moments_follow_their_rows_through_a_densification_round_cpu():
    _run_carry("cpu", fused=False)


def test_density_controller_can_reset_the_optimiser_like_the_reference():
    m, g = _mixed_model(300, "cpu")
    cfg = gb.TrainingConfig(densify_grad_threshold=0.5)
    opt = gb.GaussianOptimizer(m, cfg)
    m._xyz.grad = torch.ones_like(m._xyz)
    opt.step()
    gb.DensityController(cfg, fused=False, carry_state=False).densify_and_prune(m, opt, 1.0, grad=torch.ones(300, 3))
    assert all(len(opt.optimizer.state[grp["params"][0]]) == 0 for grp in opt.optimizer.param_groups)     # optimizer.py:67-71


@pytest.mark.gpu
def test_adam_moments_follow_their_rows_through_a_densification_round_device():
    a = _run_carry("cuda", fused=True)
    b = _run_carry("cuda", fused=False)
    assert torch.equal(a, b)                        # the device pass and the tensor-op formulation agree on every row's origin


@pytest.mark.gpu
def test_five_adam_steps_match_the_oracle_renderer_with_torch_adam():
    """train_step (CUDA renderer + fused Adam) against the same loop written with the CPU oracle renderer
    and torch.optim.Adam: losses and parameters after five steps."""
    W, H = 96, 64
    s = so.scene_aniso(300, 17)
    s["scaling"] = s["scaling"] + math.log(4.0)
    cam_o = so.camera_orbit(2, 9, W, H)
    bg = torch.tensor([0.1, 0.2, 0.3])
    g = torch.Generator().manual_seed(4)
    target = torch.rand(3, H, W, generator=g)
    cfg = gb.TrainingConfig(position_lr_init=1e-3, position_lr_final=1e-4, position_lr_max_steps=100)
    # CUDA
    m = util.cuda_model_from_params(s)
    rd = gb.GaussianRenderer()
    st = gb.RenderSettings(H, W, bg.cuda())
    opt = gb.GaussianOptimizer(m, cfg)
    cam = util.cuda_camera(cam_o)
    losses = [float(gb.train_step(m, rd, cam, target.cuda(), opt, st, it)["loss"]) for it in range(5)]
    # oracle + torch Adam, same groups / learning rates / eps
    leaf = {k: s[k].clone().requires_grad_(True) for k in util.PARAM_KEYS}
    sched = gb.LearningRateScheduler(cfg.position_lr_init, cfg.position_lr_final, int(0.01 * cfg.position_lr_max_steps),
                                     cfg.position_lr_delay_mult, cfg.position_lr_max_steps)
    ref_opt = torch.optim.Adam([{"params": [leaf["xyz"]], "lr": cfg.position_lr_init, "name": "xyz"},
                                {"params": [leaf["features_dc"]], "lr": cfg.feature_lr}, {"params": [leaf["opacity"]], "lr": cfg.opacity_lr},
                                {"params": [leaf["scaling"]], "lr": cfg.scaling_lr}, {"params": [leaf["rotation"]], "lr": cfg.rotation_lr}],
                               lr=0.0, eps=1e-15)
    ref_losses, min_abs_grad = [], {k: None for k in util.PARAM_KEYS}
    for it in range(5):
        ref_opt.param_groups[0]["lr"] = sched.get_lr(it)
        ref_opt.zero_grad(set_to_none=True)
        out = so.render_from_params(cam_o, leaf["xyz"], leaf["scaling"], leaf["rotation"], leaf["opacity"], leaf["features_dc"], bg, H, W)
        loss = (out["image"] - target).abs().mean()
        loss.backward()
        for k in util.PARAM_KEYS:
            a = leaf[k].grad.abs()
            min_abs_grad[k] = a if min_abs_grad[k] is None else torch.minimum(min_abs_grad[k], a)
        ref_opt.step()
        ref_losses.append(float(loss))
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-5 * abs(b), (losses, ref_losses)
    got = {"xyz": m._xyz, "scaling": m._scaling, "rotation": m._rotation, "opacity": m._opacity, "features_dc": m._features_dc}
    lrs = {"xyz": cfg.position_lr_init, "scaling": cfg.scaling_lr, "rotation": cfg.rotation_lr, "opacity": cfg.opacity_lr,
           "features_dc": cfg.feature_lr}
    for k in util.PARAM_KEYS:
        # Adam's update is lr * m / sqrt(v): where a gradient stays well above rounding noise in all five steps the two runs
        # move the parameter alike; where it is noise the update is +-lr whatever its size, so those entries are left out
        big = min_abs_grad[k] > 1e-3 * min_abs_grad[k].max()
        assert int(big.sum()) > 0.2 * big.numel() or k == "rotation", k
        delta = (got[k].detach().cpu() - leaf[k].detach()).abs()
        assert float(delta[big].max()) <= 0.02 * 5 * lrs[k], (k, float(delta[big].max()), lrs[k])
