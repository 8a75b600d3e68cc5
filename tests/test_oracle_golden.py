"""CPU: the oracle restatement against fixtures produced by the literal reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the
CUDA path with the oracle and with the same fixtures."""
import math

import numpy as np
import pytest
import torch

from oracle import splat_oracle as so
from tests import util


@pytest.mark.parametrize("name", util.RENDER_CASES + [util.RENDER_CASE_LARGE])
def test_oracle_matches_literal_reference_forward_and_grads(name):
    if not util.golden_available(name):
        pytest.skip("fixture not generated")
    d = util.load_golden(name)
    cam = util.golden_camera(d)
    out, grads, loss = util.oracle_render_with_grads(cam, util.golden_params(d), torch.tensor(d["bg"]),
                                                     tile_size=util.golden_tile_size(d))
    # stage outputs: identical bits where the arithmetic is the same IEEE ops
    assert np.array_equal(out["viewspace_points"].detach().numpy().view(np.uint32), d["ref_means2D"].view(np.uint32))
    assert np.array_equal(out["depths"].detach().numpy().view(np.uint32), d["ref_depths"].view(np.uint32))
    assert np.array_equal(out["visibility_filter"].numpy(), d["ref_vis"])
    assert np.array_equal(out["radii"].detach().numpy().astype(np.int64), d["ref_radii"].astype(np.int64))
    assert util.rel_err(out["conics"], torch.tensor(d["ref_conics"])) < 1e-6
    assert torch.equal(so.sort_by_depth(out["visibility_filter"], out["depths"]), torch.tensor(d["ref_sorted_idx"]))
    # images: BASELINE tolerance is 1e-4; the restatement is ~1e-7
    assert util.max_abs(out["image"], torch.tensor(d["ref_image"])) < 2e-6
    assert util.max_abs(out["alpha"], torch.tensor(d["ref_alpha"])) < 2e-6
    assert util.max_abs(out["depth"], torch.tensor(d["ref_depth"])) < 2e-5
    # gradients: tolerance 1e-3 relative; the restatement is ~1e-6
    for k in ("xyz", "scaling", "opacity", "features_dc"):
        assert util.rel_err(grads[k], torch.tensor(d["ref_g_" + k])) < 2e-5, k
    assert util.rel_err(grads["means2D"], torch.tensor(d["ref_g_means2D"])) < 2e-5
    g_rot_ref = torch.tensor(d["ref_g_rotation"])
    if not util.is_isotropic(d["in_scaling"]):       # isotropic splats: rotation grads are pure rounding noise
        assert util.rel_err(grads["rotation"], g_rot_ref) < 2e-5


def test_oracle_known_answer_two_splats():
    """The reference's own KAT (tests/test_renderer.py:127-161): alpha 0.75, depth 4/3."""
    d = util.load_golden("kat_two_splats_64x64")
    cam = util.golden_camera(d)
    feats = torch.tensor(d["in_features"])
    out = so.render(cam, torch.tensor(d["in_xyz"]), torch.tensor(d["in_cov3d"]), feats[:, 0, :],
                    torch.tensor(d["in_opacity"]).reshape(-1), torch.tensor(d["bg"]), 64, 64)
    assert abs(float(out["alpha"][0, 32, 32]) - 0.75) < 1e-3
    exp_rgb = 0.5 * torch.sigmoid(torch.tensor([1.0, 0, 0])) + 0.25 * torch.sigmoid(torch.tensor([0, 1.0, 0]))
    assert torch.allclose(out["image"][:, 32, 32], exp_rgb, atol=1e-3)
    assert abs(float(out["depth"][0, 32, 32]) - 4 / 3) < 2e-2
    assert util.max_abs(out["image"], torch.tensor(d["ref_image"])) < 1e-6
    assert util.max_abs(out["alpha"], torch.tensor(d["ref_alpha"])) < 1e-6


def test_oracle_all_behind_returns_background_once():
    """renderer.py:74-83 and the reference's test_culling_all_behind."""
    cam = so.OracleCamera(32, 32, math.radians(60), math.radians(60), torch.eye(4))
    xyz = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, -2.0]])
    cov = torch.diag_embed(torch.full((2, 3), 1e-4))
    bg = torch.tensor([0.2, 0.3, 0.4])
    out = so.render(cam, xyz, cov, torch.zeros(2, 3), torch.tensor([0.5, 0.5]), bg, 64, 64)
    assert torch.allclose(out["image"], bg.view(3, 1, 1).expand(3, 64, 64))
    assert int(torch.count_nonzero(out["alpha"])) == 0 and int(out["visibility_filter"].sum()) == 0


def test_oracle_background_counted_twice_when_something_is_visible():
    """renderer.py:273 + :359: an empty pixel shows 2*bg once any splat passed culling."""
    cam = so.OracleCamera(32, 32, math.radians(60), math.radians(60), torch.eye(4))
    xyz = torch.tensor([[0.0, 0.0, 1.0]])
    cov = torch.diag_embed(torch.full((1, 3), 1e-6))
    bg = torch.tensor([0.2, 0.1, 0.3])
    out = so.render(cam, xyz, cov, torch.zeros(1, 3), torch.tensor([0.5]), bg, 32, 32)
    assert torch.allclose(out["image"][:, 0, 0], 2 * bg)


@pytest.mark.parametrize("tag", ["c0", "orbit5of16"])
def test_oracle_stage_fixture_integer_outputs(tag):
    """Stages 1-3 of the literal reference at 1080p on 200k anisotropic splats: pixel centres and
    depths bit-equal, radii / visibility / tile rectangles exact."""
    name = f"stages_aniso_n200000_1080p_{tag}"
    if not util.golden_available(name):
        pytest.skip("fixture not generated")
    d = util.load_golden(name)
    n = int(d["n"])
    s = so.scene_aniso(n, int(d["seed"]))
    cam = util.golden_camera(d)
    with torch.no_grad():
        o = so.project(s["xyz"], so.covariance_3d(s["scaling"], s["rotation"]), cam)
        vis = so.cull(o["means2D"], o["depths"], o["radii"], cam.height, cam.width)
    assert np.array_equal(o["means2D"].numpy().view(np.uint32), d["ref_means2D_bits"])
    assert np.array_equal(o["depths"].numpy().view(np.uint32), d["ref_depth_bits"])
    assert np.array_equal(np.packbits(vis.numpy()), d["ref_vis"])
    assert np.array_equal(o["radii"].numpy().astype(np.int32), d["ref_radii"].astype(np.int32))
    cf = so.radii_closed_form(o["cov2D"])
    assert np.array_equal(cf.numpy().astype(np.int32), d["ref_radii"].astype(np.int32))
    assert util.rel_err(cf, torch.tensor(d["ref_radii"])) < 5e-7
    tx0, tx1, ty0, ty1, cnt = so.tile_rects(o["means2D"], o["radii"], cam.height, cam.width)
    rect = torch.stack([tx0, tx1, ty0, ty1], 1).numpy().astype(np.int16)
    v = vis.numpy()
    assert np.array_equal(rect[v], d["ref_rect"][v])
    assert np.array_equal(cnt.numpy().astype(np.int16)[v], d["ref_cnt"][v])


def test_bin_tiles_equals_append_loop():
    """bin_tiles (flat arrays) against a direct transcription of the reference's append loop
    semantics on a small scene: per-tile lists in global depth order."""
    s = so.scene_aniso(300, 7)
    s["scaling"] = s["scaling"] + math.log(6.0)
    cam = so.camera_orbit(2, 9, 80, 56)
    with torch.no_grad():
        o = so.project(s["xyz"], so.covariance_3d(s["scaling"], s["rotation"]), cam)
        vis = so.cull(o["means2D"], o["depths"], o["radii"], 56, 80)
        ids = so.sort_by_depth(vis, o["depths"])
        keys, flat, ranges = so.bin_tiles(ids, o["means2D"], o["radii"], o["depths"], 56, 80)
    tiles_x, tiles_y = 5, 4
    lists = [[] for _ in range(tiles_x * tiles_y)]
    for i in ids.tolist():
        r = int(o["radii"][i]); cx = int(o["means2D"][i, 0]); cy = int(o["means2D"][i, 1])
        x0, x1 = max(cx - r, 0), min(cx + 1 + r, 80)
        y0, y1 = max(cy - r, 0), min(cy + 1 + r, 56)
        if x0 >= x1 or y0 >= y1:
            continue
        for ty in range(y0 // 16, (y1 - 1) // 16 + 1):
            for tx in range(x0 // 16, (x1 - 1) // 16 + 1):
                lists[ty * tiles_x + tx].append(i)
    for t in range(tiles_x * tiles_y):
        assert flat[int(ranges[t, 0]):int(ranges[t, 1])].tolist() == lists[t]
    assert bool((keys[1:] >= keys[:-1]).all())


def test_sh_basis_of_the_oracle_is_orthonormal_on_the_sphere():
    """The SH extension is unpinned by the reference (its evaluator is a stub), so the oracle's basis is checked
    against the definition instead: Y_1..Y_15 (with Y_0 = 1/(2 sqrt(pi))) are orthonormal under the sphere integral."""
    import numpy as np
    import torch
    from oracle import splat_oracle as so
    nodes, weights = np.polynomial.legendre.leggauss(24)          # cos(theta) quadrature, exact for degree <= 47
    phis = (np.arange(48) + 0.5) * (2 * np.pi / 48)               # uniform phi: exact for trigonometric degree < 48
    ct, ph = np.meshgrid(nodes, phis, indexing="ij")
    st = np.sqrt(1 - ct ** 2)
    d = torch.tensor(np.stack([st * np.cos(ph), st * np.sin(ph), ct], axis=-1).reshape(-1, 3), dtype=torch.float64)
    w = torch.tensor((weights[:, None] * np.full_like(ph, 2 * np.pi / 48)).reshape(-1), dtype=torch.float64)
    Y = so.sh_basis(3, d)                                          # [Q, 15]
    Y = torch.cat([torch.full((Y.shape[0], 1), 0.28209479177387814, dtype=torch.float64), Y], dim=1)
    gram = (Y * w[:, None]).T @ Y
    assert torch.allclose(gram, torch.eye(16, dtype=torch.float64), atol=1e-9)
    for deg, terms in ((1, 3), (2, 8)):                            # lower degrees are prefixes of the same basis
        assert torch.equal(so.sh_basis(deg, d), so.sh_basis(3, d)[:, :terms])


def test_sh_basis_of_the_oracle_equals_scipy_spherical_harmonics():
    """Third-party pin of the SH extension (the reference's own evaluator is a stub, math_utils.py:44-49): the
    oracle's Y_1..Y_15 are the real combinations of scipy.special.sph_harm_y (complex harmonics, Condon-Shortley phase
    kept) -- sqrt(2) Im Y_l^|m| for m < 0, Y_l^0, sqrt(2) Re Y_l^m for m > 0 -- in the order m = -l..l within each
    degree, which is the convention of the original 3DGS code the north star names."""
    import numpy as np
    import torch
    scipy_special = pytest.importorskip("scipy.special")
    if not hasattr(scipy_special, "sph_harm_y"):
        pytest.skip("scipy.special.sph_harm_y needs scipy >= 1.15")
    from oracle import splat_oracle as so
    rng = np.random.default_rng(0)
    d = rng.normal(size=(2000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    theta, phi = np.arccos(d[:, 2]), np.arctan2(d[:, 1], d[:, 0])     # polar, azimuth
    cols = []
    for l in (1, 2, 3):
        for m in range(-l, l + 1):
            Y = scipy_special.sph_harm_y(l, abs(m), theta, phi)
            cols.append(Y.real if m == 0 else np.sqrt(2.0) * (Y.real if m > 0 else Y.imag))
    want = np.stack(cols, axis=1)
    got = so.sh_basis(3, torch.tensor(d, dtype=torch.float64)).numpy()
    assert np.abs(got - want).max() < 1e-13


def test_oracle_matches_the_literal_reference_on_config0():
    """BASELINE configs[0] (10 k random-init Gaussians, 256x256, C0): the literal reference's forward frame
    (~12 minutes of its pixel loop, recorded by make_golden.py config0) against the oracle."""
    import numpy as np
    import torch
    from oracle import splat_oracle as so
    if not util.golden_available(util.CONFIG0):
        pytest.skip("fixture not generated")
    d = util.load_golden(util.CONFIG0)
    s = so.scene_ref_init(int(d["n"]), int(d["seed"]))
    cam = util.golden_camera(d)
    with torch.no_grad():
        o = so.render_from_params(cam, s["xyz"], s["scaling"], s["rotation"], s["opacity"], s["features_dc"],
                                  torch.tensor(d["bg"]), cam.height, cam.width, return_stats=True)
    assert np.array_equal(o["viewspace_points"].numpy().view(np.uint32), d["ref_means2D"].view(np.uint32))
    assert np.array_equal(o["depths"].numpy().view(np.uint32), d["ref_depths"].view(np.uint32))
    assert np.array_equal(o["visibility_filter"].numpy(), d["ref_vis"])
    assert np.array_equal(o["radii"].numpy().astype(np.int64), d["ref_radii"].astype(np.int64))
    vis = torch.tensor(d["ref_vis"])
    util.assert_depth_order_equal_up_to_ties(so.sort_by_depth(vis, o["depths"]).numpy(), d["ref_sorted_idx"], d["ref_depths"])
    assert util.rel_err(o["conics"], torch.tensor(d["ref_conics"])) < 1e-5
    for k in ("image", "alpha", "depth"):
        assert util.max_abs(o[k], torch.tensor(d["ref_" + k])) < 1e-5, k


def test_oracle_matches_the_literal_reference_on_the_saturating_midsize_frame():
    """3 000 enlarged, mostly opaque anisotropic splats at 192x128: two thirds of the pixels stop at A >= 0.995
    (renderer.py:352).  The oracle follows the reference's fp32 operation order with the same libm, so not a single
    pixel may stop at a different entry: every pixel within 1e-5."""
    import numpy as np
    import torch
    from oracle import splat_oracle as so
    if not util.golden_available(util.SATURATING):
        pytest.skip("fixture not generated")
    d = util.load_golden(util.SATURATING)
    s = util.saturating_scene(d)
    cam = util.golden_camera(d)
    with torch.no_grad():
        o = so.render_from_params(cam, s["xyz"], s["scaling"], s["rotation"], s["opacity"], s["features_dc"],
                                  torch.tensor(d["bg"]), cam.height, cam.width, return_stats=True)
    assert int(d["n_depth_ties"]) == 0
    assert np.array_equal(o["viewspace_points"].numpy().view(np.uint32), d["ref_means2D"].view(np.uint32))
    assert np.array_equal(o["depths"].numpy().view(np.uint32), d["ref_depths"].view(np.uint32))
    assert np.array_equal(o["visibility_filter"].numpy(), d["ref_vis"])
    assert np.array_equal(o["radii"].numpy().astype(np.int64), d["ref_radii"].astype(np.int64))
    assert np.array_equal(so.sort_by_depth(torch.tensor(d["ref_vis"]), o["depths"]).numpy(), d["ref_sorted_idx"])
    assert int((torch.tensor(d["ref_alpha"]) >= 0.995).sum()) > 10000          # the case is what it claims to be
    for k in ("image", "alpha", "depth"):
        assert util.max_abs(o[k], torch.tensor(d["ref_" + k])) < 1e-5, k
