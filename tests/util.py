"""Shared helpers for the test-suite: golden fixtures, scene builders, comparisons."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from oracle import splat_oracle as so

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAM_KEYS = ("xyz", "scaling", "rotation", "opacity", "features_dc")
RENDER_CASES = ["aniso_n80_40x40_rot", "aniso_n120_48x40_orbit", "refinit_n300_64x64_saturating",
                "aniso_n200_96x64_bigsplats", "aniso_n100_48x40_tile8", "aniso_n90_50x44_tile12", "aniso_n100_72x56_tile32"]
# the largest frame the literal reference's autograd finishes (1 000 splats, 128x96: 289 s forward, 310 s backward, 25 GB);
# same layout as RENDER_CASES, checked on the GPU by tests/test_zz_gpu_config0.py
RENDER_CASE_LARGE = "aniso_n1000_128x96_orbit"


def golden_tile_size(d) -> int:
    return int(d["tile_size"]) if "tile_size" in d.files else 16


def golden_available(name: str) -> bool:
    return os.path.exists(os.path.join(GOLDEN_DIR, name + ".npz"))


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def golden_camera(d) -> so.OracleCamera:
    return so.OracleCamera(int(d["size_WH"][0]), int(d["size_WH"][1]), float(d["cam_fov"][0]), float(d["cam_fov"][1]),
                           torch.tensor(d["cam_WV"], dtype=torch.float32))


def golden_params(d):
    return {k: torch.tensor(d["in_" + k]) for k in PARAM_KEYS}


def is_isotropic(scaling) -> bool:
    """All three log-scales equal for every splat: the covariance does not depend on the rotation,
    so rotation gradients are rounding noise and are not compared."""
    sc = np.asarray(scaling)
    return bool(np.all(sc.max(axis=1) == sc.min(axis=1)))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the gradient metric of BASELINE.json's north_star."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def oracle_render_with_grads(cam: so.OracleCamera, params, bg, weights=None, tile_size: int = 16):
    """Oracle forward + autograd backward with the SURVEY 8d fixed-weight loss."""
    leaf = {k: params[k].clone().requires_grad_(True) for k in PARAM_KEYS}
    H, W = cam.height, cam.width
    out = so.render_from_params(cam, leaf["xyz"], leaf["scaling"], leaf["rotation"], leaf["opacity"], leaf["features_dc"],
                                bg, H, W, return_stats=True, tile_size=tile_size)
    out["viewspace_points"].retain_grad()
    loss = so.weighted_loss(out, weights if weights is not None else so.loss_weights(H, W))
    loss.backward()
    grads = {k: leaf[k].grad for k in PARAM_KEYS}
    grads["means2D"] = out["viewspace_points"].grad
    return out, grads, loss.detach()


def cuda_model_from_params(params, device="cuda"):
    import gsplat_b200 as gb
    m = gb.GaussianModel(device=device)
    m.create_from_tensors(params["xyz"], params["features_dc"], params["scaling"], params["rotation"], params["opacity"],
                          params.get("features_rest"))
    return m


def cuda_camera(cam: so.OracleCamera):
    import gsplat_b200 as gb
    return gb.Camera(cam.width, cam.height, cam.fovx, cam.fovy, world_view=cam.world_view)


def cuda_render_with_grads(cam: so.OracleCamera, params, bg, weights=None, renderer=None):
    import gsplat_b200 as gb
    m = cuda_model_from_params(params)
    H, W = cam.height, cam.width
    rd = renderer or gb.GaussianRenderer()
    out = rd.render(cuda_camera(cam), m, gb.RenderSettings(H, W, bg.cuda(), debug=True))   # debug: per-pixel n_consumed
    out["viewspace_points"].retain_grad()
    w = weights if weights is not None else so.loss_weights(H, W)
    loss = so.weighted_loss(out, tuple(t.cuda() for t in w))
    loss.backward()
    grads = {"xyz": m._xyz.grad, "scaling": m._scaling.grad, "rotation": m._rotation.grad, "opacity": m._opacity.grad,
             "features_dc": m._features_dc.grad, "features_rest": m._features_rest.grad,
             "means2D": out["viewspace_points"].grad}
    return out, grads, loss.detach(), rd, m


def assert_same_ranges(got: torch.Tensor, want: torch.Tensor) -> None:
    """[begin,end) per tile: lengths equal everywhere, positions equal for non-empty tiles (an
    empty tile's begin is unspecified: the kernel leaves (0,0), a cumulative sum leaves (k,k))."""
    got, want = got.cpu().long(), want.cpu().long()
    assert torch.equal(got[:, 1] - got[:, 0], want[:, 1] - want[:, 0])
    nz = (want[:, 1] - want[:, 0]) > 0
    assert torch.equal(got[nz], want[nz])


IMG_TOL = 1e-4          # BASELINE.json north_star: image / alpha / depth, absolute
FLIP_TOL = 1e-2         # one contribution gained or lost at the A >= 0.995 test (<= 5e-3 on colour/alpha)
FLIP_MAX_FRAC = 2e-3


def assert_images_close(got, want, ncons_got, ncons_want, what=""):
    """image/alpha/depth within IMG_TOL on every pixel whose walk ended at the same list entry as the
    oracle's.  A pixel whose accumulated opacity crosses 0.995 one splat earlier or later (the one
    discontinuity of the path; exp() differs in the last bits between CPU libm and the GPU's MUFU)
    gains or loses a single contribution: those pixels are counted, bounded in number and deviation,
    and reported -- the policy SURVEY 8c prescribes."""
    ncg, ncw = ncons_got.cpu().long(), ncons_want.cpu().long()
    same = (ncg == ncw)
    flips = int((~same).sum())
    assert flips <= max(2, int(FLIP_MAX_FRAC * same.numel())), f"{what}: {flips} termination flips of {same.numel()} pixels"
    if flips:
        assert int((ncg - ncw).abs().max()) <= 2, f"{what}: a pixel stops {int((ncg - ncw).abs().max())} entries away"
    worst, worst_flip = {}, {}
    for k in ("image", "alpha", "depth"):
        d = (got[k].detach().cpu().double() - want[k].detach().cpu().double()).abs()
        m = same.unsqueeze(0).expand_as(d)
        worst[k] = float(d[m].max()) if bool(m.any()) else 0.0
        worst_flip[k] = float(d[~m].max()) if flips else 0.0
        assert worst[k] < IMG_TOL, f"{what}: {k} differs by {worst[k]:.3e}"
        scale = 1.0 if k != "depth" else float(want[k].abs().max()) + 1.0
        assert worst_flip[k] < FLIP_TOL * scale, f"{what}: {k} differs by {worst_flip[k]:.3e} on a flipped pixel"
    return {"flips": flips, "worst": worst, "worst_flip": worst_flip}


CONFIG0 = "config0_refinit_n10000_256x256_c0"     # BASELINE configs[0], forward of the literal reference (make_golden.py config0)


def assert_depth_order_equal_up_to_ties(got_ids, ref_ids, depths) -> int:
    """The reference's argsort is unstable (renderer.py:235): ids must agree wherever the depth is unique, and the
    sequence of depths must agree everywhere.  Returns the number of positions inside tie groups."""
    got_ids, ref_ids, depths = np.asarray(got_ids, np.int64), np.asarray(ref_ids, np.int64), np.asarray(depths)
    assert got_ids.shape == ref_ids.shape
    dg, dr = depths[got_ids], depths[ref_ids]
    assert np.array_equal(dg.view(np.uint32), dr.view(np.uint32)), "depth sequences differ"
    tied = np.zeros(dg.shape[0], bool)
    eq = dg[1:] == dg[:-1]
    tied[1:] |= eq
    tied[:-1] |= eq
    assert np.array_equal(got_ids[~tied], ref_ids[~tied]), "depth order differs outside tie groups"
    assert np.array_equal(np.sort(got_ids[tied]), np.sort(ref_ids[tied]))
    return int(tied.sum())


SATURATING = "aniso_n3000_192x128_orbit_saturating_fwd"   # literal reference forward, 2/3 of the pixels terminate early


def saturating_scene(d):
    """The scene make_golden.py `saturating` rendered, rebuilt from the parameters stored in the fixture."""
    s = so.scene_aniso(int(d["n"]), int(d["seed"]))
    s["scaling"] = s["scaling"] + float(d["scale_boost"])
    s["opacity"] = s["opacity"] + float(d["opacity_boost"])
    return s
