"""Generate golden fixtures by running the *literal, unmodified* reference renderer.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [case ...]        # config0 / saturating / aniso_n1000_128x96_orbit only when named

For every case it
  1. builds a seeded synthetic scene (oracle.splat_oracle generators = SURVEY 8d scenes),
  2. renders it with /root/reference's GaussianRenderer.render through the duck-typed adapters the
     reference's own test uses (tests/test_renderer.py:7-53), forward + autograd backward,
  3. stores inputs, outputs and gradients as tests/golden/<case>.npz,
  4. prints how far the oracle restatement is from the literal reference.

The fixtures are small (N <= 400, <= 128x96) because the reference's pixel loop costs ~190 us per
(pixel x list entry) and its autograd backward is super-linear.
"""
from __future__ import annotations

import contextlib
import io
import math
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import splat_oracle as so  # noqa: E402


def import_reference():
    sys.path.insert(0, "/root/reference")
    with contextlib.redirect_stdout(io.StringIO()):      # config/config.py prints on import
        from src.core.renderer import GaussianRenderer, RenderSettings
        from src.core.gaussian_model import GaussianModel
        from config.config import TrainingConfig
    return GaussianRenderer, RenderSettings, GaussianModel, TrainingConfig


class RefCamera:
    """4 attributes + a callable, exactly what renderer.py:140-152 reads."""

    def __init__(self, cam: so.OracleCamera):
        self._width, self._height = cam.width, cam.height
        self._FoVx, self._FoVy = cam.fovx, cam.fovy
        self._WV = cam.world_view

    def world_view_transform(self):
        return self._WV


class RefGaussians:
    """Forwards to a real reference GaussianModel; get_covariance -> compute_3d_covariance()
    because the model's own property is broken (gaussian_model.py:124-128)."""

    def __init__(self, model):
        self.m = model

    get_xyz = property(lambda s: s.m.get_xyz)
    get_opacity = property(lambda s: s.m.get_opacity)
    get_features = property(lambda s: s.m.get_features)
    get_covariance = property(lambda s: s.m.compute_3d_covariance())

    @property
    def _features_dc(self):
        return self.m._features_dc


class CovGaussians:
    """The reference test's DummyGaussians shape: covariance handed over directly."""

    def __init__(self, xyz, cov3d, features, opacity):
        self._xyz, self._cov, self._features, self._opacity = xyz, cov3d, features, opacity

    get_xyz = property(lambda s: s._xyz)
    get_opacity = property(lambda s: s._opacity)
    get_features = property(lambda s: s._features)
    get_covariance = property(lambda s: s._cov)


# ---------------------------------------------------------------------------------------------
# case table
# ---------------------------------------------------------------------------------------------
CASES = {
    # name: dict(scene, n, seed, W, H, cam, bg, scale_boost (added to log-scale), opacity_boost)
    "aniso_n120_48x40_orbit": dict(scene="aniso", n=120, seed=3, W=48, H=40, cam=("orbit", 1, 8),
                                   bg=(0.25, 0.1, 0.4), scale_boost=math.log(8.0), opacity_boost=0.0),
    "refinit_n300_64x64_saturating": dict(scene="ref", n=300, seed=0, W=64, H=64, cam=("c0",),
                                          bg=(0.0, 0.0, 0.0), scale_boost=math.log(6.0), opacity_boost=3.0),
    "aniso_n80_40x40_rot": dict(scene="aniso", n=80, seed=5, W=40, H=40, cam=("orbit", 3, 7),
                                bg=(0.25, 0.1, 0.4), scale_boost=math.log(10.0), opacity_boost=1.0),
    "aniso_n200_96x64_bigsplats": dict(scene="aniso", n=200, seed=11, W=96, H=64, cam=("c0",),
                                       bg=(0.6, 0.6, 0.6), scale_boost=math.log(25.0), opacity_boost=-1.0),
    # the largest frame the reference's autograd finishes here: 1.2 M pixel x entry evaluations, ~5 GB of graph; no pixel
    # saturates (max alpha 0.98), so no termination flip can blur the gradient comparison
    "aniso_n1000_128x96_orbit": dict(scene="aniso", n=1000, seed=37, W=128, H=96, cam=("orbit", 2, 7),
                                     bg=(0.1, 0.2, 0.3), scale_boost=math.log(3.0), opacity_boost=-0.4),
    # other tile sizes (renderer.py:24 takes any): which splats a pixel sees depends on its tile's list (3-sigma rectangles
    # against tiles, renderer.py:263-298), so the image itself changes with the tile size
    "aniso_n100_48x40_tile8": dict(scene="aniso", n=100, seed=13, W=48, H=40, cam=("orbit", 2, 9), tile=8,
                                   bg=(0.1, 0.3, 0.2), scale_boost=math.log(6.0), opacity_boost=1.0),
    "aniso_n90_50x44_tile12": dict(scene="aniso", n=90, seed=17, W=50, H=44, cam=("orbit", 5, 9), tile=12,
                                   bg=(0.3, 0.1, 0.2), scale_boost=math.log(7.0), opacity_boost=1.5),
    "aniso_n100_72x56_tile32": dict(scene="aniso", n=100, seed=19, W=72, H=56, cam=("c0",), tile=32,
                                    bg=(0.2, 0.2, 0.5), scale_boost=math.log(6.0), opacity_boost=2.0),
}


def build_scene(spec):
    gen = so.scene_aniso if spec["scene"] == "aniso" else so.scene_ref_init
    s = gen(spec["n"], spec["seed"])
    s["scaling"] = s["scaling"] + spec["scale_boost"]
    s["opacity"] = s["opacity"] + spec["opacity_boost"]
    if spec["cam"][0] == "c0":
        cam = so.camera_c0(spec["W"], spec["H"])
    else:
        cam = so.camera_orbit(spec["cam"][1], spec["cam"][2], spec["W"], spec["H"])
    return s, cam


def run_reference(spec, s, cam):
    GaussianRenderer, RenderSettings, GaussianModel, TrainingConfig = import_reference()
    m = GaussianModel(TrainingConfig())
    P = torch.nn.Parameter
    m._xyz, m._features_dc, m._features_rest = P(s["xyz"].clone()), P(s["features_dc"].clone()), P(s["features_rest"].clone())
    m._scaling, m._rotation, m._opacity = P(s["scaling"].clone()), P(s["rotation"].clone()), P(s["opacity"].clone())
    H, W = spec["H"], spec["W"]
    settings = RenderSettings(image_height=H, image_width=W, bg_color=torch.tensor(spec["bg"], dtype=torch.float32))
    rd = GaussianRenderer(tile_size=spec.get("tile", 16), radius_min=0.01, radius_max=50.0)
    t0 = time.time()
    out = rd.render(RefCamera(cam), RefGaussians(m), settings)
    t_fwd = time.time() - t0
    out["viewspace_points"].retain_grad()
    weights = so.loss_weights(H, W)
    loss = so.weighted_loss(out, weights)
    t0 = time.time()
    loss.backward()
    t_bwd = time.time() - t0
    # the reference's own stages again, for the intermediate tensors render() does not return
    with torch.no_grad():
        proj = rd._project_gaussians_3d_to_2d(RefCamera(cam), RefGaussians(m))
        vis = rd._frustum_culling(proj["means2D"], proj["depths"], proj["radii"], settings)
        sorted_idx = rd._sort_gaussians_by_depth(vis, proj["depths"]) if int(vis.sum()) else torch.zeros(0, dtype=torch.int64)
    d = proj["depths"][vis]
    n_ties = int(d.numel() - torch.unique(d).numel())
    res = {
        "image": out["image"], "alpha": out["alpha"], "depth": out["depth"],
        "means2D": out["viewspace_points"], "conics": out["conics"], "radii": out["radii"],
        "vis": out["visibility_filter"], "depths": proj["depths"], "cov2D": proj["cov2D"],
        "sorted_idx": sorted_idx, "loss": loss.detach(),
        "g_xyz": m._xyz.grad, "g_scaling": m._scaling.grad, "g_rotation": m._rotation.grad,
        "g_opacity": m._opacity.grad, "g_features_dc": m._features_dc.grad,
        "g_features_rest_absmax": m._features_rest.grad.abs().max(),
        "g_means2D": out["viewspace_points"].grad,
    }
    return res, t_fwd, t_bwd, n_ties


def run_oracle(spec, s, cam):
    leaf = {k: s[k].clone().requires_grad_(True) for k in ("xyz", "scaling", "rotation", "opacity", "features_dc")}
    H, W = spec["H"], spec["W"]
    out = so.render_from_params(cam, leaf["xyz"], leaf["scaling"], leaf["rotation"], leaf["opacity"],
                                leaf["features_dc"], torch.tensor(spec["bg"]), H, W, return_stats=True,
                                tile_size=spec.get("tile", 16))
    out["viewspace_points"].retain_grad()
    loss = so.weighted_loss(out, so.loss_weights(H, W))
    loss.backward()
    return out, leaf, loss


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_case(name):
    spec = CASES[name]
    s, cam = build_scene(spec)
    ref, t_fwd, t_bwd, n_ties = run_reference(spec, s, cam)
    assert n_ties == 0, f"{name}: {n_ties} visible depth ties; the reference sort order would be unspecified"
    o, leaf, oloss = run_oracle(spec, s, cam)
    print(f"[{name}] reference fwd {t_fwd:.1f}s bwd {t_bwd:.1f}s  visible {int(ref['vis'].sum())}/{spec['n']}"
          f"  saturated px {int((ref['alpha'] >= 0.995).sum())}  clamped ch {int((ref['image'] >= 1).sum())}")
    print("   oracle vs literal:  image %.2e  alpha %.2e  depth %.2e  means2D %.2e (bit-equal %.4f)  depths bit-equal %.4f" % (
        float((o["image"] - ref["image"]).abs().max()), float((o["alpha"] - ref["alpha"]).abs().max()),
        float((o["depth"] - ref["depth"]).abs().max()), float((o["viewspace_points"] - ref["means2D"]).abs().max()),
        float((o["viewspace_points"].detach().view(torch.int32) == ref["means2D"].detach().view(torch.int32)).float().mean()),
        float((o["depths"].detach().view(torch.int32) == ref["depths"].view(torch.int32)).float().mean())))
    print("   radii max rel %.2e  int(radii) mismatches %d  vis mismatches %d  sorted order equal %s" % (
        rel(o["radii"], ref["radii"]), int((o["radii"].detach().int() != ref["radii"].detach().int()).sum()),
        int((o["visibility_filter"] != ref["vis"]).sum()),
        bool(torch.equal(so.sort_by_depth(o["visibility_filter"], o["depths"]), ref["sorted_idx"]))))
    print("   grads rel: xyz %.2e scaling %.2e rotation %.2e opacity %.2e dc %.2e means2D %.2e" % (
        rel(leaf["xyz"].grad, ref["g_xyz"]), rel(leaf["scaling"].grad, ref["g_scaling"]),
        rel(leaf["rotation"].grad, ref["g_rotation"]), rel(leaf["opacity"].grad, ref["g_opacity"]),
        rel(leaf["features_dc"].grad, ref["g_features_dc"]), rel(o["viewspace_points"].grad, ref["g_means2D"])))
    save = {"in_" + k: s[k].numpy() for k in ("xyz", "scaling", "rotation", "opacity", "features_dc")}
    save.update({"cam_WV": cam.world_view.numpy(), "cam_fov": np.array([cam.fovx, cam.fovy], dtype=np.float64),
                 "size_WH": np.array([spec["W"], spec["H"]]), "bg": np.array(spec["bg"], dtype=np.float32),
                 "tile_size": np.array(spec.get("tile", 16))})
    save.update({"ref_" + k: v.detach().numpy() for k, v in ref.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)


def make_kat_two_splats():
    """The reference's own known-answer scene (tests/test_renderer.py:127-161), through the
    covariance-input path, with gradients added."""
    GaussianRenderer, RenderSettings, _, _ = import_reference()
    H = W = 64
    xyz = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 2.0]], requires_grad=True)
    sig = torch.full((2, 3), 0.01)
    cov = torch.diag_embed(sig ** 2).requires_grad_(True)
    feats = torch.zeros(2, 16, 3)
    feats[0, 0, 0] = 1.0
    feats[1, 0, 1] = 1.0
    feats.requires_grad_(True)
    op = torch.tensor([[0.5], [0.5]], requires_grad=True)
    cam = so.OracleCamera(W, H, math.radians(60.0), math.radians(60.0), torch.eye(4))
    settings = RenderSettings(image_height=H, image_width=W, bg_color=torch.zeros(3))
    out = GaussianRenderer().render(RefCamera(cam), CovGaussians(xyz, cov, feats, op), settings)
    out["viewspace_points"].retain_grad()
    loss = so.weighted_loss(out, so.loss_weights(H, W))
    loss.backward()
    save = dict(in_xyz=xyz.detach().numpy(), in_cov3d=cov.detach().numpy(), in_features=feats.detach().numpy(),
                in_opacity=op.detach().numpy(), cam_WV=np.eye(4, dtype=np.float32),
                cam_fov=np.array([cam.fovx, cam.fovy]), size_WH=np.array([W, H]), bg=np.zeros(3, np.float32),
                ref_image=out["image"].detach().numpy(), ref_alpha=out["alpha"].detach().numpy(),
                ref_depth=out["depth"].detach().numpy(), ref_means2D=out["viewspace_points"].detach().numpy(),
                ref_conics=out["conics"].detach().numpy(), ref_radii=out["radii"].detach().numpy(),
                ref_vis=out["visibility_filter"].numpy(), ref_loss=loss.detach().numpy(),
                ref_g_xyz=xyz.grad.numpy(), ref_g_cov3d=cov.grad.numpy(), ref_g_features=feats.grad.numpy(),
                ref_g_opacity=op.grad.numpy(), ref_g_means2D=out["viewspace_points"].grad.numpy())
    np.savez_compressed(os.path.join(HERE, "kat_two_splats_64x64.npz"), **save)
    cy = cx = 32
    print("[kat_two_splats] centre alpha %.6f rgb %s depth %.6f" % (
        float(out["alpha"][0, cy, cx].detach()), out["image"][:, cy, cx].tolist(), float(out["depth"][0, cy, cx].detach())))


def make_stage_fixture(n=200000, seed=0):
    """Stages 1-3 of the literal reference on the bench scene geometry (1080p, C0 and one orbit
    view), reduced to what pins the integer outputs: int radii, visibility mask, depth bits,
    tile rectangles.  Stored for a strided subset so the file stays small."""
    GaussianRenderer, RenderSettings, GaussianModel, TrainingConfig = import_reference()
    W, H = 1920, 1080
    for tag, cam in (("c0", so.camera_c0(W, H)), ("orbit5of16", so.camera_orbit(5, 16, W, H))):
        s = so.scene_aniso(n, seed)
        m = GaussianModel(TrainingConfig())
        P = torch.nn.Parameter
        m._xyz, m._features_dc, m._features_rest = P(s["xyz"]), P(s["features_dc"]), P(s["features_rest"])
        m._scaling, m._rotation, m._opacity = P(s["scaling"]), P(s["rotation"]), P(s["opacity"])
        rd = GaussianRenderer()
        settings = RenderSettings(image_height=H, image_width=W, bg_color=torch.zeros(3))
        with torch.no_grad():
            proj = rd._project_gaussians_3d_to_2d(RefCamera(cam), RefGaussians(m))
            vis = rd._frustum_culling(proj["means2D"], proj["depths"], proj["radii"], settings)
            cov3d = so.covariance_3d(s["scaling"], s["rotation"])
            o = so.project(s["xyz"], cov3d, cam)
            ovis = so.cull(o["means2D"], o["depths"], o["radii"], H, W)
        print(f"[stages_{tag}] N={n}: means2D bit-equal {float((o['means2D'].view(torch.int32) == proj['means2D'].view(torch.int32)).float().mean()):.6f}"
              f"  depths bit-equal {float((o['depths'].view(torch.int32) == proj['depths'].view(torch.int32)).float().mean()):.6f}"
              f"  radii rel {rel(o['radii'], proj['radii']):.2e} closed-form rel {rel(so.radii_closed_form(o['cov2D']), proj['radii']):.2e}"
              f"  int(radii) mism {int((o['radii'].int() != proj['radii'].int()).sum())}"
              f"  closed-form int mism {int((so.radii_closed_form(o['cov2D']).int() != proj['radii'].int()).sum())}"
              f"  vis mism {int((ovis != vis).sum())}  conics rel {rel(o['conics'], proj['conics']):.2e}")
        tx0, tx1, ty0, ty1, cnt = so.tile_rects(proj["means2D"], proj["radii"], H, W)
        np.savez_compressed(
            os.path.join(HERE, f"stages_aniso_n{n}_1080p_{tag}.npz"),
            n=np.array(n), seed=np.array(seed), cam_WV=cam.world_view.numpy(), cam_fov=np.array([cam.fovx, cam.fovy]),
            size_WH=np.array([W, H]),
            ref_means2D_bits=proj["means2D"].numpy().view(np.uint32), ref_depth_bits=proj["depths"].numpy().view(np.uint32),
            ref_radii=proj["radii"].numpy(), ref_vis=np.packbits(vis.numpy()),
            ref_conics_sub=proj["conics"].numpy()[::16],
            ref_rect=torch.stack([tx0, tx1, ty0, ty1], 1).numpy().astype(np.int16), ref_cnt=cnt.numpy().astype(np.int16))


def make_densify_fixture(n=1500, seed=61):
    """Density control through the reference's OWN functions -- GaussianModel.density_and_clone, density_and_split and
    prune_points (src/core/gaussian_model.py:130-197) -- on a seeded model, with the `_append_points` patch of the
    reference's own test (tests/test_gaussian_model.py:103-111: the stock method reads a non-existent `_scaling_log`).
    Driver order = the one mini-3d-gaussian-splatting_b200.training.DensityController uses: both masks from the same
    pre-densification gradient, clone first (appended rows do not disturb the split indices), then split, then the opacity
    prune of optimizer.py:64-66 (whose `get_opacity()` call is itself a bug: the property is read here)."""
    import types
    import torch.nn as nn
    GaussianRenderer, RenderSettings, GaussianModel, TrainingConfig = import_reference()
    s = so.scene_aniso(n, seed)
    g = torch.Generator().manual_seed(7)
    sc = s["scaling"].clone()
    sc[: n // 3] = math.log(0.05) + 0.1 * torch.randn(n // 3, 3, generator=g)              # large  -> split candidates
    sc[n // 3: 2 * n // 3] = math.log(0.004) + 0.1 * torch.randn(n // 3, 3, generator=g)   # small  -> clone candidates
    op = s["opacity"].clone()
    op[::7] = -9.0                                         # transparent -> pruned (also as clone copies / children)
    op[5::11] = 8.0                                        # children's logit is clamped to 6
    grad = torch.zeros(n, 3)
    hot = torch.rand(n, generator=g) < 0.6
    grad[hot] = 1.0 + torch.rand(int(hot.sum()), 3, generator=g)
    rest = 0.1 * torch.randn(n, 15, 3, generator=g)
    th, extent, min_opacity, noise_seed = 0.5, 1.0, 0.01, 3

    m = GaussianModel(TrainingConfig())
    P = nn.Parameter
    m._xyz, m._features_dc, m._features_rest = P(s["xyz"].clone()), P(s["features_dc"].clone()), P(rest.clone())
    m._scaling, m._rotation, m._opacity = P(sc.clone()), P(s["rotation"].clone()), P(op.clone())

    def _append_points_patch(self, xyz, fdc, frest, scaling_log, rot, op_):       # tests/test_gaussian_model.py:103-111
        self._xyz = nn.Parameter(torch.cat([self._xyz.data, xyz], dim=0))
        self._features_dc = nn.Parameter(torch.cat([self._features_dc.data, fdc], dim=0))
        self._features_rest = nn.Parameter(torch.cat([self._features_rest.data, frest], dim=0))
        self._scaling = nn.Parameter(torch.cat([self._scaling.data, scaling_log], dim=0))
        self._rotation = nn.Parameter(torch.cat([self._rotation.data, rot], dim=0))
        self._opacity = nn.Parameter(torch.cat([self._opacity.data, op_], dim=0))
    m._append_points = types.MethodType(_append_points_patch, m)

    size = m.get_scaling.mean(dim=-1)
    clone_mask = (grad.norm(dim=-1) > th) & (size < 0.01 * extent)
    split_mask = (grad.norm(dim=-1) > th) & (size > 0.03 * extent)
    k = int(clone_mask.sum())
    torch.manual_seed(noise_seed)
    noise = torch.randn(k, 3)                              # the block density_and_clone's randn_like draws next
    torch.manual_seed(noise_seed)
    m._xyz.grad = grad.clone()
    m.density_and_clone(th, extent)                        # reference code
    assert m.get_num_points() == n + k
    g2 = torch.cat([torch.where(split_mask.unsqueeze(-1), grad, torch.zeros_like(grad)), torch.zeros(k, 3)])
    m._xyz.grad = g2
    m.density_and_split(th, extent)                        # reference code
    n_split = int(split_mask.sum())
    assert m.get_num_points() == n + k + n_split
    keep = m.get_opacity.squeeze(1) > min_opacity          # optimizer.py:64 (property, not a call)
    m.prune_points(keep)                                   # reference code
    print(f"[densify] n={n}: {k} clone candidates, {n_split} split candidates, {int((~keep).sum())} rows pruned -> {m.get_num_points()} rows")
    np.savez_compressed(
        os.path.join(HERE, f"densify_n{n}.npz"),
        th=np.array(th), extent=np.array(extent), min_opacity=np.array(min_opacity),
        in_xyz=s["xyz"].numpy(), in_features_dc=s["features_dc"].numpy(), in_features_rest=rest.numpy(), in_scaling=sc.numpy(),
        in_rotation=s["rotation"].numpy(), in_opacity=op.numpy(), in_grad=grad.numpy(), noise=noise.numpy(),
        n_clone_candidates=np.array(k), n_split_candidates=np.array(n_split),
        ref_xyz=m._xyz.data.numpy(), ref_features_dc=m._features_dc.data.numpy(), ref_features_rest=m._features_rest.data.numpy(),
        ref_scaling=m._scaling.data.numpy(), ref_rotation=m._rotation.data.numpy(), ref_opacity=m._opacity.data.numpy())


def _forward_fixture(name, s, cam, bg, extra):
    """Forward frame of the literal reference (no autograd: its backward does not fit memory beyond a few hundred
    splats) + the stage outputs that pin the integer work: pixel-centre and depth bits, radii, visibility,
    depth order."""
    GaussianRenderer, RenderSettings, GaussianModel, TrainingConfig = import_reference()
    W, H = cam.width, cam.height
    n = s["xyz"].shape[0]
    m = GaussianModel(TrainingConfig())
    P = torch.nn.Parameter
    m._xyz, m._features_dc, m._features_rest = P(s["xyz"].clone()), P(s["features_dc"].clone()), P(s["features_rest"].clone())
    m._scaling, m._rotation, m._opacity = P(s["scaling"].clone()), P(s["rotation"].clone()), P(s["opacity"].clone())
    settings = RenderSettings(image_height=H, image_width=W, bg_color=torch.tensor(bg, dtype=torch.float32))
    rd = GaussianRenderer()
    t0 = time.time()
    with torch.no_grad():
        out = rd.render(RefCamera(cam), RefGaussians(m), settings)
        t_fwd = time.time() - t0
        proj = rd._project_gaussians_3d_to_2d(RefCamera(cam), RefGaussians(m))
        vis = rd._frustum_culling(proj["means2D"], proj["depths"], proj["radii"], settings)
        sorted_idx = rd._sort_gaussians_by_depth(vis, proj["depths"])
        d = proj["depths"][vis]
        n_ties = int(d.numel() - torch.unique(d).numel())
        o = so.render_from_params(cam, s["xyz"], s["scaling"], s["rotation"], s["opacity"], s["features_dc"],
                                  torch.tensor(bg), H, W, return_stats=True)
    print(f"[{name}] literal reference forward {t_fwd:.1f}s  visible {int(vis.sum())}/{n}  visible depth ties {n_ties}"
          f"  saturated px {int((out['alpha'] >= 0.995).sum())}  clamped ch {int((out['image'] >= 1).sum())}")
    print("   oracle vs literal:  image %.2e  alpha %.2e  depth %.2e  means2D bit-equal %.4f  depths bit-equal %.4f"
          "  int(radii) mismatches %d  vis mismatches %d" % (
              float((o["image"] - out["image"]).abs().max()), float((o["alpha"] - out["alpha"]).abs().max()),
              float((o["depth"] - out["depth"]).abs().max()),
              float((o["viewspace_points"].view(torch.int32) == proj["means2D"].view(torch.int32)).float().mean()),
              float((o["depths"].view(torch.int32) == proj["depths"].view(torch.int32)).float().mean()),
              int((o["radii"].int() != proj["radii"].int()).sum()), int((o["visibility_filter"] != vis).sum())))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        n=np.array(n), cam_WV=cam.world_view.numpy(), cam_fov=np.array([cam.fovx, cam.fovy]),
        size_WH=np.array([W, H]), bg=np.array(bg, dtype=np.float32), n_depth_ties=np.array(n_ties),
        ref_image=out["image"].numpy(), ref_alpha=out["alpha"].numpy(), ref_depth=out["depth"].numpy(),
        ref_means2D=out["viewspace_points"].numpy(), ref_depths=proj["depths"].numpy(), ref_radii=out["radii"].numpy(),
        ref_conics=out["conics"].numpy(), ref_vis=out["visibility_filter"].numpy(), ref_sorted_idx=sorted_idx.numpy(),
        ref_seconds=np.array(t_fwd), **extra)


def make_config0_fixture(n=10000, seed=0, W=256, H=256):
    """BASELINE.json configs[0] -- the reference's own CPU-runnable case (examples/simple_scene.py: 10 k random-init
    Gaussians, 256x256, one synthetic camera) -- rendered FORWARD by the literal reference (~13 minutes of its Python
    pixel loop)."""
    _forward_fixture(f"config0_refinit_n{n}_{W}x{H}_c0", so.scene_ref_init(n, seed), so.camera_c0(W, H), (0.0, 0.0, 0.0),
                     dict(seed=np.array(seed)))


SATURATING = dict(n=3000, seed=23, W=192, H=128, cam=(3, 11), bg=(0.2, 0.3, 0.1), scale_boost=math.log(4.0), opacity_boost=2.0)


def saturating_scene(spec=SATURATING):
    s = so.scene_aniso(spec["n"], spec["seed"])
    s["scaling"] = s["scaling"] + spec["scale_boost"]
    s["opacity"] = s["opacity"] + spec["opacity_boost"]
    return s, so.camera_orbit(spec["cam"][0], spec["cam"][1], spec["W"], spec["H"])


def make_saturating_fixture():
    """The termination rule at scale against the literal reference: 3 000 anisotropic, enlarged, mostly opaque splats
    seen from an orbit camera at 192x128 over a non-zero background -- two thirds of the pixels reach A >= 0.995 and
    stop early (renderer.py:352), hundreds of channels clamp (renderer.py:359), lists of ~340 entries per tile."""
    s, cam = saturating_scene()
    sp = SATURATING
    _forward_fixture(f"aniso_n{sp['n']}_{sp['W']}x{sp['H']}_orbit_saturating_fwd", s, cam, sp["bg"],
                     dict(seed=np.array(sp["seed"]), scale_boost=np.array(sp["scale_boost"]), opacity_boost=np.array(sp["opacity_boost"]),
                          orbit=np.array(sp["cam"])))


if __name__ == "__main__":
    torch.set_num_threads(4)
    # named only: config0 (13 min), saturating (6 min), aniso_n1000_128x96_orbit (10 min, 25 GB of autograd graph)
    want = sys.argv[1:] or (["kat", "stages", "densify"] + [c for c in CASES if c != "aniso_n1000_128x96_orbit"])
    for c in want:
        if c == "kat":
            make_kat_two_splats()
        elif c == "stages":
            make_stage_fixture()
        elif c == "densify":
            make_densify_fixture()
        elif c == "config0":
            make_config0_fixture()
        elif c == "saturating":
            make_saturating_fixture()
        else:
            make_case(c)
