"""CPU: the scene container behind the renderer's fused input path, checked the way the reference checks its own
(/root/reference/tests/test_gaussian_model.py:35-140 -- initialisation shapes, activated accessors, the covariance
R diag(sigma^2) R^T with [w,x,y,z] quaternions, and the point counts of split / clone), here WITHOUT the monkey-patches
that file needs to get the reference's densification to run (gaussian_model.py:229 reads a missing `_scaling_log`)."""
from __future__ import annotations

import math

import torch

import gsplat_b200 as gb


def _rotation_from_wxyz(q: torch.Tensor) -> torch.Tensor:
    """Independent statement of math_utils.py:9-26 through the axis-angle form: R = I + sin(t) K + (1 - cos t) K^2."""
    q = torch.nn.functional.normalize(q.double(), dim=-1)
    w, v = q[:, 0].clamp(-1, 1), q[:, 1:]
    s = v.norm(dim=-1)
    axis = torch.where(s[:, None] > 1e-12, v / s[:, None].clamp_min(1e-300), torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64).expand_as(v))
    theta = 2.0 * torch.atan2(s, w)
    K = torch.zeros(q.shape[0], 3, 3, dtype=torch.float64)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    eye = torch.eye(3, dtype=torch.float64).expand_as(K)
    return eye + torch.sin(theta)[:, None, None] * K + (1 - torch.cos(theta))[:, None, None] * (K @ K)


def _model(n: int, extent: float = 1.0) -> "gb.GaussianModel":
    m = gb.GaussianModel(device="cpu")
    m.create_from_random(n, extent, seed=11)
    return m


def test_initialisation_shapes_and_statistics_buffers():
    n = 256
    m = _model(n)
    assert m.get_num_points() == n
    assert m.get_xyz.shape == (n, 3) and m._features_dc.shape == (n, 1, 3) and m._features_rest.shape == (n, 15, 3)
    assert m._scaling.shape == (n, 3) and m._rotation.shape == (n, 4) and m._opacity.shape == (n, 1)
    assert m.xyz_gradient_accum.shape == (n, 3) and m.denom.shape == (n, 1) and m.max_radii2D.shape[0] == n
    # gaussian_model.py:78-98: positions in the cube of the scene extent, sigma = 0.02 * extent, opacity = sigmoid(-2)
    assert float(m.get_xyz.detach().abs().max()) <= 1.0
    assert torch.allclose(m.get_scaling, torch.full((n, 3), 0.02), rtol=1e-6)
    assert torch.allclose(m.get_opacity, torch.full((n, 1), 1.0 / (1.0 + math.exp(2.0))), rtol=1e-6)
    assert float(m._features_rest.detach().abs().max()) == 0.0


def test_activated_accessors():
    n = 64
    m = _model(n)
    assert m.get_scaling.shape == (n, 3) and bool((m.get_scaling > 0).all())
    qn = torch.linalg.norm(m.get_rotation, dim=-1)
    assert torch.allclose(qn, torch.ones_like(qn), atol=1e-5)
    assert bool(((m.get_opacity > 0) & (m.get_opacity < 1)).all())
    feats = m.get_features
    assert feats.shape == (n, 16, 3) and torch.equal(feats, torch.cat([m._features_dc, m._features_rest], dim=1))
    # the three activations are the very functions the renderer looks for before fusing them (renderer._is_parameter_model)
    assert m.scaling_activation is torch.exp and m.opacity_activation is torch.sigmoid
    assert m.rotation_activation is torch.nn.functional.normalize


def test_covariance_is_r_diag_sigma_squared_rt_with_wxyz_quaternions():
    n = 32
    m = _model(n)
    with torch.no_grad():
        m._scaling += 0.5 * torch.randn(n, 3, generator=torch.Generator().manual_seed(1))      # anisotropic: R matters
    cov = m.compute_3d_covariance()
    assert cov.shape == (n, 3, 3)
    R = _rotation_from_wxyz(m._rotation.detach())
    want = R @ torch.diag_embed(m.get_scaling.detach().double() ** 2) @ R.transpose(-1, -2)
    assert torch.allclose(cov.double(), want, rtol=1e-5, atol=1e-9)
    assert torch.equal(m.get_covariance, cov)                  # the reference's property raises (gaussian_model.py:124-128)
    assert bool((torch.linalg.eigvalsh(cov.double()) > 0).all())
    # a quarter turn about z, written [w,x,y,z]: x -> y
    q = torch.tensor([[math.cos(math.pi / 4), 0.0, 0.0, math.sin(math.pi / 4)]])
    assert torch.allclose(gb.scene.quaternion_to_rotation(q)[0] @ torch.tensor([1.0, 0.0, 0.0]), torch.tensor([0.0, 1.0, 0.0]), atol=1e-6)


def test_split_adds_one_net_point_per_large_splat_and_clone_one_per_small_splat():
    n, extent, k = 64, 1.0, 8
    m = _model(n, extent)
    with torch.no_grad():
        m._scaling[:k] = math.log(0.06 * extent)               # > 0.03 * extent: split
        m._scaling[k:2 * k] = math.log(0.005 * extent)         # < 0.01 * extent: clone
    m._xyz.grad = torch.ones_like(m._xyz)
    n0 = m.get_num_points()
    m.density_and_split(grad_threshold=0.5, scene_extent=extent)
    assert m.get_num_points() == n0 + k                        # k originals leave, 2k children arrive
    sig = m.get_scaling
    assert int((sig.mean(-1) > 0.03 * extent).sum()) == 2 * k  # children: sigma / 1.6 = 0.0375
    m._xyz.grad = torch.ones_like(m._xyz)
    n1 = m.get_num_points()
    m.density_and_clone(grad_threshold=0.5, scene_extent=extent)
    assert m.get_num_points() == n1 + k
    for p in (m._xyz, m._features_dc, m._features_rest, m._scaling, m._rotation, m._opacity):
        assert p.shape[0] == n1 + k and p.requires_grad
    assert m.xyz_gradient_accum.shape[0] == n1 + k and m.denom.shape[0] == n1 + k and m.max_radii2D.shape[0] == n1 + k


def test_math_utils_carry_the_references_names_and_the_sh_rows_of_the_device_path():
    """math_utils.py:7-49: rotation and covariance equal the model's own; degree 0 returns the DC row (the reference's
    behaviour); degrees 1..3 equal DC + the oracle's basis (pinned to scipy in tests/test_oracle_golden.py) times the
    higher-order rows."""
    from oracle import splat_oracle as so
    n = 50
    m = _model(n)
    with torch.no_grad():
        m._scaling += 0.4 * torch.randn(n, 3, generator=torch.Generator().manual_seed(2))
    assert torch.equal(gb.MathUtils.build_rotation_matrix(m._rotation), gb.scene.quaternion_to_rotation(m._rotation))
    assert torch.allclose(gb.MathUtils.build_covariance_3d(m.get_scaling, m._rotation), m.compute_3d_covariance(), rtol=1e-5, atol=1e-9)
    g = torch.Generator().manual_seed(4)
    coeffs = torch.randn(n, 16, 3, generator=g, dtype=torch.float64)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)
    assert torch.equal(gb.MathUtils.spherical_harmonics_eval(0, dirs, coeffs), coeffs[:, 0])
    for deg in (1, 2, 3):
        Y = so.sh_basis(deg, dirs)
        want = coeffs[:, 0] + (Y.unsqueeze(-1) * coeffs[:, 1:1 + Y.shape[1]]).sum(dim=1)
        assert torch.allclose(gb.MathUtils.spherical_harmonics_eval(deg, dirs, coeffs), want, rtol=1e-12, atol=1e-12)
