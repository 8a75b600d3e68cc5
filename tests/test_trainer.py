"""CPU: the trainer surface of the reference (src/train/trainer.py:12-97, all stubs there) driven end to end with the
CPU oracle of the render path standing in for the CUDA renderer: the loop lowers the loss, density control runs on its
schedule with the optimiser's moments following their rows, a checkpoint resumes bit for bit."""
from __future__ import annotations

import math

import pytest
import torch

import gsplat_b200 as gb
from oracle import splat_oracle as so

W, H, N_VIEWS = 32, 24, 4


class OracleRenderer:
    """oracle/splat_oracle.py (pinned to the literal reference) behind GaussianRenderer.render's signature."""

    def render(self, camera, gaussians, settings):
        cam = so.OracleCamera(camera._width, camera._height, camera._FoVx, camera._FoVy, camera.world_view_transform())
        return so.render_from_params(cam, gaussians._xyz, gaussians._scaling, gaussians._rotation, gaussians._opacity,
                                     gaussians._features_dc, settings.bg_color, settings.image_height, settings.image_width)


def _scene(seed, n=60):
    s = so.scene_aniso(n, seed)
    m = gb.GaussianModel(device="cpu")
    m.create_from_tensors(s["xyz"], s["features_dc"], s["scaling"] + math.log(10.0), s["rotation"], s["opacity"] + 1.0)
    return m


def _cameras(truth):
    cams, rd = [], OracleRenderer()
    for k in range(N_VIEWS):
        c = gb.Camera.orbit(k, N_VIEWS, W, H)
        with torch.no_grad():
            img = rd.render(c, truth, gb.RenderSettings(H, W, torch.tensor([0.1, 0.1, 0.1])))["image"]
        cams.append(gb.Camera(W, H, c._FoVx, c._FoVy, world_view=c.world_view_transform(), uid=k, image=img.clone()))
    return cams


def _trainer(tmp_path, cams, **cfg):
    config = gb.TrainingConfig(device="cpu", output_path=str(tmp_path), iterations=12, position_lr_init=0.01, position_lr_final=0.001,
                               position_lr_max_steps=100, feature_lr=0.05, opacity_lr=0.05, scaling_lr=0.02, rotation_lr=0.01, **cfg)
    model = _scene(7)
    with torch.no_grad():                       # start away from the truth
        model._features_dc += 0.8
        model._xyz += 0.05
    t = gb.GaussianTrainer(config, cameras=cams, gaussians=model, renderer=OracleRenderer(), bg_color=(0.1, 0.1, 0.1), seed=5)
    t.setup()
    return t


@pytest.fixture(scope="module")
def cams():
    return _cameras(_scene(7))


def test_setup_builds_the_pieces_and_the_scene_extent_is_the_camera_rig_radius(tmp_path, cams):
    t = _trainer(tmp_path, cams)
    assert isinstance(t.optimizer, gb.GaussianOptimizer) and isinstance(t.density_controller, gb.DensityController)
    # orbit cameras sit on a circle of radius 3 at height 0.9 around the y axis: every centre is 3 from the centroid
    assert t.scene_extent == pytest.approx(1.1 * 3.0, rel=1e-5)
    assert t.get_scene_extent() == t.scene_extent
    with pytest.raises(ValueError, match="_image"):
        t.train_step(gb.Camera.orbit(0, 4, W, H))


def test_training_lowers_the_loss_and_validation_reports_l1_and_psnr(tmp_path, cams):
    t = _trainer(tmp_path, cams, densify_from_iter=1000)
    before = t.validate()
    t.train(iterations=12, val_every=6)
    after = t.validate()
    assert t.iteration == 12 and len(t.train_losses) == 12 and len(t.val_losses) == 2
    assert after["l1"] < 0.8 * before["l1"] and after["psnr"] > before["psnr"] and after["views"] == N_VIEWS
    assert all(math.isfinite(v) for v in t.train_losses)


def test_density_control_runs_on_schedule_and_the_optimiser_follows_the_rows(tmp_path, cams):
    t = _trainer(tmp_path, cams, densify_from_iter=4, densify_interval=4, densify_until_iter=8, densify_grad_threshold=1e-7)
    seen = []
    for _ in range(9):
        cam = t.cameras[t.iteration % N_VIEWS]
        stats = t.train_step(cam)
        if "split" in stats:
            seen.append((int(stats["iteration"]), int(stats["split"]), int(stats["cloned"]), int(stats["pruned"]), int(stats["points"])))
    assert [s[0] for s in seen] == [4, 8]                               # optimizer.py:39-41 schedule
    assert sum(s[1] + s[2] for s in seen) > 0                             # something was densified
    n = t.gaussians.get_num_points()
    assert n == seen[-1][4]
    for group in t.optimizer.optimizer.param_groups:                      # moments were carried, row for row
        p = group["params"][0]
        assert p.shape[0] == n
        st = t.optimizer.optimizer.state.get(p)
        if st:
            assert st["exp_avg"].shape == p.shape and st["exp_avg_sq"].shape == p.shape
    assert t.gaussians.xyz_gradient_accum.shape[0] == n


def test_checkpoint_resumes_bit_for_bit(tmp_path, cams):
    a = _trainer(tmp_path, cams, densify_from_iter=1000)
    a.train(iterations=6)
    a.save_checkpoint(6)
    a.train(iterations=10)
    b = gb.GaussianTrainer(a.config, cameras=cams, renderer=OracleRenderer(), bg_color=(0.1, 0.1, 0.1), seed=99)
    b.load_checkpoint(6)
    assert b.iteration == 6 and b.train_losses == a.train_losses[:6]
    b.train(iterations=10)
    assert b.train_losses == a.train_losses                              # same cameras drawn, same arithmetic
    for name in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity"):
        assert torch.equal(getattr(a.gaussians, name), getattr(b.gaussians, name)), name
