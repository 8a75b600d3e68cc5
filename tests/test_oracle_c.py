"""CPU: the plain-C port (oracle/splat_oracle.c) against the torch oracle (itself pinned to the literal
reference) and against the literal-reference fixtures -- forward, bins and every gradient."""
import math

import numpy as np
import pytest
import torch

from oracle import c_port, splat_oracle as so
from tests import util


def _c_render(cam, params, bg):
    cam16 = c_port.camera_block(cam.width, cam.height, cam.fovx, cam.fovy, cam.world_view.numpy())
    p = {k: v.numpy() for k, v in params.items() if isinstance(v, torch.Tensor)}
    w = tuple(t.numpy() for t in so.loss_weights(cam.height, cam.width))
    return c_port.render_fwd_bwd(cam16, cam.width, cam.height, p, np.asarray(bg, np.float32), w)


@pytest.mark.parametrize("name", ["aniso_n80_40x40_rot", "aniso_n120_48x40_orbit", "refinit_n300_64x64_saturating",
                                  "aniso_n200_96x64_bigsplats", "aniso_n1000_128x96_orbit"])
def test_c_port_matches_literal_reference_fixture(name):
    if not util.golden_available(name):
        pytest.skip("fixture not generated")
    d = util.load_golden(name)
    cam = util.golden_camera(d)
    res = _c_render(cam, util.golden_params(d), d["bg"])
    pr = res["proj"]
    assert np.array_equal(pr["means2D"].view(np.uint32), d["ref_means2D"].view(np.uint32))
    assert np.array_equal(pr["depths"].view(np.uint32), d["ref_depths"].view(np.uint32))
    assert np.array_equal(pr["vis"].astype(bool), d["ref_vis"])
    assert np.array_equal(pr["radii"].astype(np.int64), d["ref_radii"].astype(np.int64))
    assert np.array_equal(res["sorted_ids"].astype(np.int64), d["ref_sorted_idx"])
    for k in ("image", "alpha", "depth"):
        assert np.abs(res[k] - d["ref_" + k]).max() < 1e-5, k
    g = res["grads"]
    pairs = [("xyz", g["xyz"]), ("scaling", g["scaling"]), ("opacity", g["opacity"].reshape(-1, 1)),
             ("features_dc", g["feat0"].reshape(-1, 1, 3)), ("means2D", res["g_raster"]["means2D"])]
    if not util.is_isotropic(d["in_scaling"]):
        pairs.append(("rotation", g["rotation"]))
    for k, got in pairs:
        assert util.rel_err(torch.tensor(got), torch.tensor(d["ref_g_" + k])) < 1e-4, k


def test_c_port_matches_torch_oracle_midsize():
    s = so.scene_aniso(1500, 41)
    s["scaling"] = s["scaling"] + math.log(3.0)
    s["opacity"] = s["opacity"] + 1.0
    cam = so.camera_orbit(4, 9, 112, 80)
    bg = torch.tensor([0.3, 0.2, 0.1])
    o_out, o_grads, _ = util.oracle_render_with_grads(cam, s, bg)
    res = _c_render(cam, s, bg.numpy())
    assert np.array_equal(res["entry_ids"].astype(np.int64), o_out["sort_ids"].numpy())
    util.assert_same_ranges(torch.tensor(res["ranges"]), o_out["tile_ranges"])
    assert np.array_equal(res["n_consumed"].astype(np.int64), o_out["n_consumed"].numpy())
    for k in ("image", "alpha", "depth"):
        assert util.max_abs(torch.tensor(res[k]), o_out[k]) < 1e-5, k
    g = res["grads"]
    for k, got in [("xyz", g["xyz"]), ("scaling", g["scaling"]), ("rotation", g["rotation"]),
                   ("opacity", g["opacity"].reshape(-1, 1)), ("features_dc", g["feat0"].reshape(-1, 1, 3)),
                   ("means2D", res["g_raster"]["means2D"])]:
        assert util.rel_err(torch.tensor(got), o_grads[k]) < 1e-4, k


@pytest.mark.parametrize("tag", ["c0", "orbit5of16"])
def test_c_port_stage_fixture_integer_outputs_1080p(tag):
    """The checker of the full-size GPU tests, itself checked at scale: stages 1-3 of the literal reference at 1080p
    on 200k anisotropic splats -- pixel centres and depths bit-equal, visibility, int(radii), tile rectangles and tile
    counts exact (rectangles of splats with an empty AABB are undefined and left out)."""
    name = f"stages_aniso_n200000_1080p_{tag}"
    if not util.golden_available(name):
        pytest.skip("fixture not generated")
    d = util.load_golden(name)
    s = so.scene_aniso(int(d["n"]), int(d["seed"]))
    cam = util.golden_camera(d)
    cam16 = c_port.camera_block(cam.width, cam.height, cam.fovx, cam.fovy, cam.world_view.numpy())
    pr = c_port.project(cam16, cam.width, cam.height, s["xyz"].numpy(), s["scaling"].numpy(), s["rotation"].numpy(), None,
                        s["opacity"].numpy(), True, s["features_dc"].numpy().reshape(-1, 3))
    assert np.array_equal(pr["means2D"].view(np.uint32), d["ref_means2D_bits"])
    assert np.array_equal(pr["depths"].view(np.uint32), d["ref_depth_bits"])
    assert np.array_equal(np.packbits(pr["vis"].astype(bool)), d["ref_vis"])
    assert np.array_equal(pr["radii"].astype(np.int32), d["ref_radii"].astype(np.int32))
    assert np.abs(pr["radii"] - d["ref_radii"]).max() <= 5e-7 * np.abs(d["ref_radii"]).max()
    v = pr["vis"].astype(bool)
    assert np.array_equal(pr["tiles_touched"][v], d["ref_cnt"][v].astype(np.int32))
    binned = v & (pr["tiles_touched"] > 0)
    # the port stores (x0, y0, x1, y1), the fixture (x0, x1, y0, y1)
    assert np.array_equal(pr["rect"][binned][:, [0, 2, 1, 3]], d["ref_rect"][binned].astype(np.int32))


def test_c_port_matches_the_literal_reference_on_config0():
    """BASELINE configs[0] through the C port against the literal reference's recorded forward frame."""
    if not util.golden_available(util.CONFIG0):
        pytest.skip("fixture not generated")
    d = util.load_golden(util.CONFIG0)
    s = so.scene_ref_init(int(d["n"]), int(d["seed"]))
    cam = util.golden_camera(d)
    cam16 = c_port.camera_block(cam.width, cam.height, cam.fovx, cam.fovy, cam.world_view.numpy())
    p = {k: s[k].numpy() for k in ("xyz", "scaling", "rotation", "opacity", "features_dc")}
    res = c_port.render_fwd_bwd(cam16, cam.width, cam.height, p, d["bg"], None, backward=False)
    pr = res["proj"]
    assert np.array_equal(pr["means2D"].view(np.uint32), d["ref_means2D"].view(np.uint32))
    assert np.array_equal(pr["depths"].view(np.uint32), d["ref_depths"].view(np.uint32))
    assert np.array_equal(pr["vis"].astype(bool), d["ref_vis"])
    assert np.array_equal(pr["radii"].astype(np.int64), d["ref_radii"].astype(np.int64))
    util.assert_depth_order_equal_up_to_ties(res["sorted_ids"], d["ref_sorted_idx"], d["ref_depths"])
    for k in ("image", "alpha", "depth"):
        assert np.abs(res[k] - d["ref_" + k]).max() < 1e-5, k


def test_c_port_matches_the_literal_reference_on_the_saturating_midsize_frame():
    """The C port (its own expf) against the literal reference where two thirds of the pixels terminate early: pixels
    may stop one entry apart, counted and bounded; all others within 1e-5."""
    if not util.golden_available(util.SATURATING):
        pytest.skip("fixture not generated")
    d = util.load_golden(util.SATURATING)
    s = util.saturating_scene(d)
    cam = util.golden_camera(d)
    cam16 = c_port.camera_block(cam.width, cam.height, cam.fovx, cam.fovy, cam.world_view.numpy())
    p = {k: s[k].numpy() for k in ("xyz", "scaling", "rotation", "opacity", "features_dc")}
    res = c_port.render_fwd_bwd(cam16, cam.width, cam.height, p, d["bg"], None, backward=False)
    pr = res["proj"]
    assert np.array_equal(pr["means2D"].view(np.uint32), d["ref_means2D"].view(np.uint32))
    assert np.array_equal(pr["depths"].view(np.uint32), d["ref_depths"].view(np.uint32))
    assert np.array_equal(pr["vis"].astype(bool), d["ref_vis"])
    assert np.array_equal(pr["radii"].astype(np.int64), d["ref_radii"].astype(np.int64))
    with torch.no_grad():        # the oracle's per-pixel walk lengths stand for the literal reference's (test_oracle_golden)
        o = so.render_from_params(cam, s["xyz"], s["scaling"], s["rotation"], s["opacity"], s["features_dc"],
                                  torch.tensor(d["bg"]), cam.height, cam.width, return_stats=True)
    got = {k: torch.tensor(res[k]) for k in ("image", "alpha", "depth")}
    ref = {k: torch.tensor(d["ref_" + k]) for k in ("image", "alpha", "depth")}
    util.assert_images_close(got, ref, torch.tensor(res["n_consumed"]), o["n_consumed"], "C port, saturating frame")
