"""Scene I/O and camera helpers: the reference's working behaviour -- io_utils.py:17-85,
camera.py:79-141, gaussian_model.py:43-76 -- and round trips of the formats added here.  CPU only."""
import math
import sys

import numpy as np
import pytest
import torch

import gsplat_b200 as gb
from gsplat_b200 import CameraUtils, IOUtils


def test_world_view_matrix_matches_reference_function():
    rng = np.random.default_rng(0)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], dtype=np.float32)
    C = rng.normal(size=3).astype(np.float32)
    v = CameraUtils.build_world_view_matrix(R, C, from_c2w=True)
    assert torch.allclose(v[:3, :3], torch.from_numpy(R).T) and torch.allclose(v[:3, 3], -torch.from_numpy(R).T @ torch.from_numpy(C))
    assert torch.allclose(v @ torch.tensor([*C, 1.0]), torch.tensor([0.0, 0.0, 0.0, 1.0]), atol=1e-6)   # the centre maps to the origin
    v2 = CameraUtils.build_world_view_matrix(R, C, from_c2w=False)
    assert torch.equal(v2[:3, :3], torch.from_numpy(R)) and torch.equal(v2[:3, 3], torch.from_numpy(C))
    try:                                       # the literal reference function, when the reference tree is present
        sys.path.insert(0, "/root/reference")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            from src.core.camera import CameraUtils as RefCU
        assert torch.equal(RefCU.build_world_view_matrix(R, C, True), v)
        assert torch.equal(RefCU.build_world_view_matrix(R, C, False), v2)
    except ImportError:
        pass
    finally:
        if "/root/reference" in sys.path:
            sys.path.remove("/root/reference")
    cam = gb.Camera(64, 48, math.radians(60), world_view=v)
    assert torch.allclose(cam.camera_center, torch.from_numpy(C), atol=1e-6)
    f = CameraUtils.fov_to_focal(math.radians(60), 64)
    assert CameraUtils.focal_to_fov(f, 64) == pytest.approx(math.radians(60))


def test_projection_matrix_is_the_references_in_both_of_its_forms():
    """tests/test_camera.py of the reference wants its two definitions (half-angle tangents; focal lengths and image
    size) to agree; only the second survives in its class (same method name twice), so that one is called live when the
    reference tree is present, and the first is restated here."""
    cases = [(60, 45, 0.1, 1000.0, 640, 480), (90, 67.5, 0.1, 500.0, 1920, 1080), (30, 22.5, 0.5, 2000.0, 800, 600)]
    ref = None
    try:
        sys.path.insert(0, "/root/reference")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            from src.core.camera import CameraUtils as RefCU
        ref = RefCU.build_projection_matrix
    except ImportError:
        pass
    finally:
        if "/root/reference" in sys.path:
            sys.path.remove("/root/reference")
    for fx_deg, fy_deg, zn, zf, w, h in cases:
        fx, fy = math.radians(fx_deg), math.radians(fy_deg)
        P = CameraUtils.build_projection_matrix(zn, zf, fx, fy)
        assert torch.equal(P, CameraUtils.build_projection_matrix(zn, zf, fx, fy, w, h))
        tangents = torch.tensor([[1 / math.tan(fx / 2), 0, 0, 0], [0, 1 / math.tan(fy / 2), 0, 0],
                                 [0, 0, -(zf + zn) / (zf - zn), -2 * zf * zn / (zf - zn)], [0, 0, -1, 0]], dtype=torch.float32)
        assert torch.allclose(P, tangents, rtol=1e-6, atol=0)
        if ref is not None:
            assert torch.allclose(P, ref(zn, zf, fx, fy, w, h), rtol=1e-6, atol=0)
        # -z forward: the near and far planes land on NDC z = -1 and +1
        for z, want in ((-zn, -1.0), (-zf, 1.0)):
            clip = P @ torch.tensor([0.0, 0.0, z, 1.0])
            assert float(clip[2] / clip[3]) == pytest.approx(want, abs=1e-4)


@pytest.mark.parametrize("ext", [".npz", ".npy", ".ply", ".xyz"])
def test_point_cloud_round_trip(tmp_path, ext):
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(57, 3)).astype(np.float32)
    cols = rng.uniform(size=(57, 3)).astype(np.float32)
    for c in (cols, None):
        path = str(tmp_path / f"cloud{'' if c is None else '_c'}{ext}")
        IOUtils.save_point_cloud(pts, c, path)
        p2, c2 = IOUtils.load_point_cloud(path)
        assert p2.dtype == np.float32 and p2.shape == (57, 3)
        assert np.allclose(p2, pts, atol=1e-6)
        if c is None:
            assert c2 is None
        else:
            assert np.allclose(c2, c, atol=(1 / 255 if ext == ".ply" else 1e-6))


def test_colmap_points3d_and_ascii_ply(tmp_path):
    p = tmp_path / "points3D.txt"
    p.write_text("# 3D point list\n1 0.5 -1.0 2.0 255 0 127 0.3 1 2 3 4\n2 1.5 1.0 0.0 0 255 0 0.1 5 6\nbad line\n")
    pts, cols = IOUtils.load_point_cloud(str(p))
    assert pts.tolist() == [[0.5, -1.0, 2.0], [1.5, 1.0, 0.0]]
    assert np.allclose(cols, [[1.0, 0.0, 127 / 255], [0.0, 1.0, 0.0]])
    q = tmp_path / "a.ply"
    q.write_text("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n0 1 2 255 128 0\n3 4 5 0 0 255\n")
    pts, cols = IOUtils.load_point_cloud(str(q))
    assert pts.tolist() == [[0, 1, 2], [3, 4, 5]] and np.allclose(cols[0], [1.0, 128 / 255, 0.0])


def test_create_from_pcd_follows_reference_initialisation(tmp_path):
    pts = np.array([[0, 0, 0], [2, 0, 0], [0, 4, 0], [0, 0, 6]], dtype=np.float32)
    cols = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0.5, 0.5]], dtype=np.float32)
    path = str(tmp_path / "pc.npz")
    IOUtils.save_point_cloud(pts, cols, path)
    m = gb.GaussianModel(device="cpu")
    m.create_from_pcd(path, spatial_lr_scale=2.0, seed=0)
    assert m.get_num_points() == 4 and m._features_rest.shape == (4, 15, 3)
    assert torch.equal(m._features_dc[:, 0, :], torch.from_numpy(cols))
    extent = (2 + 4 + 6) / 3                                   # mean side of the bounding box (gaussian_model.py:63)
    assert torch.allclose(m._scaling, torch.full((4, 3), math.log(0.01 * extent * 2.0)))
    assert torch.allclose(m._rotation.norm(dim=-1), torch.ones(4)) and torch.all(m._opacity == 0.5)
    with pytest.raises(ValueError):
        IOUtils.save_point_cloud(np.zeros((0, 3), np.float32), None, str(tmp_path / "e.npz"))
        m.create_from_pcd(str(tmp_path / "e.npz"))


def test_model_and_image_round_trip(tmp_path):
    m = gb.GaussianModel(device="cpu")
    m.create_from_random(33, 1.0, seed=4)
    IOUtils.save_gaussians(m, str(tmp_path / "m.npz"))
    m2 = IOUtils.load_gaussians(gb.GaussianModel(device="cpu"), str(tmp_path / "m.npz"))
    for k in IOUtils.PARAMS:
        assert torch.equal(getattr(m, k).data, getattr(m2, k).data), k
    img = torch.rand(3, 12, 20)
    for name in ("a.png", "b.npy"):
        IOUtils.save_image(img, str(tmp_path / "out" / name))
        back = IOUtils.load_image(str(tmp_path / "out" / name))
        assert back.shape == (3, 12, 20) and float((back - img).abs().max()) <= 1 / 255 + 1e-6
