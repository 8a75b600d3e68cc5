"""Scene container and camera that feed the renderer.

`GaussianModel` keeps the reference's parameter names, shapes and activations
(src/core/gaussian_model.py:15-122,200-236) so the renderer's fused path, an optimiser and the
densification routines see what they would see on the reference model.  The reference pieces
that do not run are fixed, not reproduced:
  * `get_covariance` (gaussian_model.py:124-128 raises) returns compute_3d_covariance();
  * `_append_points` (gaussian_model.py:229 reads a non-existent `_scaling_log`) concatenates
    `_scaling`;
behaviour is otherwise the working subset pinned by the reference's tests/test_gaussian_model.py.

`Camera` holds what renderer.py:140-152 reads: `_width, _height, _FoVx, _FoVy` and a *callable*
`world_view_transform()`; the reference's own Camera exposes that as a broken property
(camera.py:45-50), so the renderer is duck-typed on the callable, as the reference's tests are.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Camera:
    def __init__(self, width: int, height: int, FoVx: float, FoVy: Optional[float] = None,
                 world_view: Optional[torch.Tensor] = None, uid: int = 0, image: Optional[torch.Tensor] = None,
                 image_name: str = ""):
        self._uid = uid
        self._width, self._height = int(width), int(height)
        self._FoVx = float(FoVx)
        # square pixels unless told otherwise
        self._FoVy = float(FoVy) if FoVy is not None else 2.0 * math.atan(math.tan(self._FoVx / 2) * height / width)
        self._wv = torch.eye(4, dtype=torch.float32) if world_view is None else world_view.detach().to("cpu", torch.float32)
        self._image = image
        self._image_name = image_name

    def world_view_transform(self) -> torch.Tensor:
        """4x4 world-to-camera matrix, X_c = R_wc X_w + t_wc (camera.py:79-141 convention)."""
        return self._wv

    @property
    def camera_center(self) -> torch.Tensor:
        R, t = self._wv[:3, :3], self._wv[:3, 3]
        return -(R.T @ t)

    @staticmethod
    def from_c2w(R_cw: np.ndarray, C_w: np.ndarray, **kw) -> "Camera":
        """Camera from a camera-to-world rotation and centre: R_wc = R_cw^T, t = -R_cw^T C
        (camera.py:128-131)."""
        R = torch.as_tensor(np.asarray(R_cw), dtype=torch.float32)
        C = torch.as_tensor(np.asarray(C_w), dtype=torch.float32).reshape(3)
        wv = torch.eye(4, dtype=torch.float32)
        wv[:3, :3] = R.T
        wv[:3, 3] = -(R.T @ C)
        return Camera(world_view=wv, **kw)

    @staticmethod
    def look_at_origin_c0(width: int, height: int, fov_deg: float = 60.0) -> "Camera":
        """SURVEY 8d camera C0: identity rotation, t = (0,0,3)."""
        wv = torch.eye(4, dtype=torch.float32)
        wv[2, 3] = 3.0
        return Camera(width, height, math.radians(fov_deg), world_view=wv)

    @staticmethod
    def orbit(k: int, M: int, width: int, height: int, fov_deg: float = 60.0) -> "Camera":
        """SURVEY 8d orbit view k of M: centre 3*(sin t, 0.3, -cos t) looking at the origin."""
        th = 2.0 * math.pi * k / M
        C = 3.0 * np.array([math.sin(th), 0.3, -math.cos(th)])
        f = -C / np.linalg.norm(C)
        r = np.cross(np.array([0.0, 1.0, 0.0]), f)
        r /= np.linalg.norm(r)
        u = np.cross(f, r)
        R_wc = np.stack([r, u, f])
        wv = torch.eye(4, dtype=torch.float32)
        wv[:3, :3] = torch.tensor(R_wc, dtype=torch.float32)
        wv[:3, 3] = torch.tensor(-R_wc @ C, dtype=torch.float32)
        return Camera(width, height, math.radians(fov_deg), world_view=wv)


def quaternion_to_rotation(q: torch.Tensor) -> torch.Tensor:
    """[N,4] (w,x,y,z) -> [N,3,3]; normalises first (math_utils.py:9-26)."""
    w, x, y, z = F.normalize(q, dim=-1).unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.view(-1, 3, 3)


class GaussianModel(nn.Module):
    """Parameter container with the reference's layout: `_xyz [N,3]`, `_features_dc [N,1,3]`,
    `_features_rest [N,15,3]`, `_scaling [N,3]` (log sigma), `_rotation [N,4]` (w,x,y,z),
    `_opacity [N,1]` (logit); statistics buffers `xyz_gradient_accum`, `denom`, `max_radii2D`."""

    def __init__(self, config=None, device: Optional[str] = None):
        """`config` is accepted for signature parity with the reference (a TrainingConfig whose
        `.device` is honoured, config/config.py:67); only the device is used."""
        super().__init__()
        self.config = config
        self.max_sh_degree = 3
        if device is None and isinstance(config, (str, torch.device)):
            device, self.config = config, None
        if device is None and getattr(config, "device", None) == "cpu":
            device = "cpu"
        dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._xyz = nn.Parameter(torch.empty(0, 3, device=dev))
        self._features_dc = nn.Parameter(torch.empty(0, 1, 3, device=dev))
        self._features_rest = nn.Parameter(torch.empty(0, 15, 3, device=dev))
        self._scaling = nn.Parameter(torch.empty(0, 3, device=dev))
        self._rotation = nn.Parameter(torch.empty(0, 4, device=dev))
        self._opacity = nn.Parameter(torch.empty(0, 1, device=dev))
        self.register_buffer("xyz_gradient_accum", torch.zeros(0, 3, device=dev))
        self.register_buffer("denom", torch.zeros(0, 1, device=dev))
        self.register_buffer("max_radii2D", torch.zeros(0, device=dev))
        # the three activations the projection kernel fuses (gaussian_model.py:34-40)
        self.scaling_activation = torch.exp
        self.scaling_inverse_activation = torch.log
        self.opacity_activation = torch.sigmoid
        self.opacity_inverse_activation = self.inverse_sigmoid
        self.rotation_activation = F.normalize

    # ---- initialisation ---------------------------------------------------------------------
    @torch.no_grad()
    def create_from_random(self, num_points: int, scene_extent: float = 1.0, seed: Optional[int] = None) -> None:
        """gaussian_model.py:78-98.  With `seed`, the draws come from a CPU generator in the
        reference's order, so the scene is identical on every device and rank."""
        dev = self._xyz.device
        if seed is None:
            xyz = (torch.rand(num_points, 3, device=dev) - 0.5) * (2.0 * scene_extent)
            dc = torch.rand(num_points, 1, 3, device=dev)
            rot = F.normalize(torch.randn(num_points, 4, device=dev), dim=-1)
        else:
            g = torch.Generator().manual_seed(seed)
            xyz = ((torch.rand(num_points, 3, generator=g) - 0.5) * (2.0 * scene_extent)).to(dev)
            dc = torch.rand(num_points, 1, 3, generator=g).to(dev)
            rot = F.normalize(torch.randn(num_points, 4, generator=g), dim=-1).to(dev)
        self._set(xyz, dc, torch.zeros(num_points, 15, 3, device=dev),
                  torch.full((num_points, 3), math.log(0.02 * scene_extent), device=dev), rot,
                  torch.full((num_points, 1), -2.0, device=dev))

    @torch.no_grad()
    def create_from_pcd(self, pcd_path: str, spatial_lr_scale: float = 1.0, seed: Optional[int] = None) -> None:
        """Initialise from a point cloud file (gaussian_model.py:43-76): colours -> features_dc (white when the
        file has none), log-scale = log(0.01 * max(extent, 0.01) * spatial_lr_scale), random unit quaternions,
        opacity parameter 0.5."""
        from .io_utils import IOUtils
        points, colors = IOUtils.load_pcd(pcd_path)
        if points.size == 0:
            raise ValueError("No points found in the PCD file.")
        n = points.shape[0]
        dev = self._xyz.device
        xyz = torch.from_numpy(points).float()
        cols = torch.ones(n, 3) if colors is None else torch.from_numpy(colors).float()
        extent = float((xyz.max(dim=0).values - xyz.min(dim=0).values).mean())
        base_scale = 0.01 * max(extent, 1e-2) * spatial_lr_scale
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        rot = F.normalize(torch.randn(n, 4, generator=g), dim=-1)
        self._set(xyz.to(dev), cols[:, None, :].contiguous().to(dev), torch.zeros(n, 15, 3, device=dev),
                  torch.full((n, 3), math.log(base_scale), device=dev), rot.to(dev), torch.full((n, 1), 0.5, device=dev))

    def create_from_tensors(self, xyz, features_dc, scaling, rotation, opacity, features_rest=None) -> None:
        dev = self._xyz.device
        n = xyz.shape[0]
        rest = torch.zeros(n, 15, 3) if features_rest is None else features_rest
        self._set(*(t.detach().to(dev, torch.float32).contiguous().clone()
                    for t in (xyz, features_dc.reshape(n, 1, 3), rest, scaling, rotation, opacity.reshape(n, 1))))

    def _set(self, xyz, dc, rest, scaling, rot, opacity) -> None:
        self._xyz, self._features_dc, self._features_rest = nn.Parameter(xyz), nn.Parameter(dc), nn.Parameter(rest)
        self._scaling, self._rotation, self._opacity = nn.Parameter(scaling), nn.Parameter(rot), nn.Parameter(opacity)
        self._reset_stats()

    def _reset_stats(self) -> None:
        n, dev = self._xyz.shape[0], self._xyz.device
        self.xyz_gradient_accum = torch.zeros(n, 3, device=dev)
        self.denom = torch.zeros(n, 1, device=dev)
        self.max_radii2D = torch.zeros(n, device=dev)

    # ---- accessors (gaussian_model.py:101-128) --------------------------------------------------
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        if self._features_rest.numel() == 0:
            return self._features_dc
        return torch.cat([self._features_dc, self._features_rest], dim=1)

    @property
    def get_scaling(self):
        return self.scaling_activation(self._scaling)

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_opacity(self):
        return self.opacity_activation(self._opacity)

    @property
    def get_covariance(self):
        return self.compute_3d_covariance()

    def compute_3d_covariance(self) -> torch.Tensor:
        """R diag(sigma^2) R^T -> [N,3,3] (gaussian_model.py:200-207)."""
        R = quaternion_to_rotation(self.get_rotation)
        return (R * (self.get_scaling ** 2).unsqueeze(1)) @ R.transpose(-1, -2)

    def get_num_points(self) -> int:
        return int(self._xyz.shape[0])

    @staticmethod
    def inverse_sigmoid(x):
        return torch.log(x / (1 - x))

    @torch.no_grad()
    def reset_opacity(self, new_opacity: float = 0.01) -> None:
        val = min(max(new_opacity, 1e-4), 1 - 1e-4)
        self._opacity.data.fill_(math.log(val / (1 - val)))

    # ---- densification (gaussian_model.py:130-197,224-236) ----------------------------------------
    @torch.no_grad()
    def add_densification_stats(self, viewspace_grad: torch.Tensor, visibility: torch.Tensor, radii: torch.Tensor) -> None:
        """Accumulate what the reference allocates but never fills (gaussian_model.py:29-31):
        per-splat sum of |dL/d means2D| over views where visible, visit count, max screen radius."""
        gn = viewspace_grad.norm(dim=-1, keepdim=True)
        self.xyz_gradient_accum[:, :1] += torch.where(visibility.unsqueeze(-1), gn, torch.zeros_like(gn))
        self.denom += visibility.unsqueeze(-1).to(self.denom.dtype)
        self.max_radii2D = torch.maximum(self.max_radii2D, torch.where(visibility, radii, torch.zeros_like(radii)))

    @torch.no_grad()
    def prune_points(self, keep_mask: torch.Tensor) -> None:
        """Keep rows where mask is True and rebuild the parameters (gaussian_model.py:181-197)."""
        self._set(*(p.data[keep_mask] for p in (self._xyz, self._features_dc, self._features_rest, self._scaling,
                                                self._rotation, self._opacity)))

    @torch.no_grad()
    def _append_points(self, xyz, fdc, frest, scaling_log, rot, op) -> None:
        cat = lambda a, b: torch.cat([a.data, b], dim=0)  # noqa: E731
        self._set(cat(self._xyz, xyz), cat(self._features_dc, fdc), cat(self._features_rest, frest),
                  cat(self._scaling, scaling_log), cat(self._rotation, rot), cat(self._opacity, op))

    @torch.no_grad()
    def density_and_split(self, grad_threshold: float, scene_extent: float, grad: Optional[torch.Tensor] = None) -> int:
        """Large, high-gradient splats -> two children at +-0.5*mean(sigma) along the first
        principal axis, sigma x0.75 (gaussian_model.py:130-156).  Returns the number split."""
        g = self._xyz.grad if grad is None else grad
        if g is None:
            return 0
        sig = self.get_scaling
        mask = (g.norm(dim=-1) > grad_threshold) & (sig.mean(dim=-1) > 0.03 * scene_extent)
        k = int(mask.sum())
        if k == 0:
            return 0
        xyz, s, rot = self._xyz.data[mask], sig[mask], self.get_rotation[mask]
        offset = quaternion_to_rotation(rot)[:, :, 0] * (s.mean(dim=-1, keepdim=True) * 0.5)
        children = (torch.cat([xyz - offset, xyz + offset], 0),
                    self._features_dc.data[mask].repeat(2, 1, 1), self._features_rest.data[mask].repeat(2, 1, 1),
                    torch.log(s * 0.75).repeat(2, 1), rot.repeat(2, 1),
                    torch.logit(self.get_opacity[mask])[:, 0:1].clamp(-6, 6).repeat(2, 1))
        self.prune_points(~mask)
        self._append_points(*children)
        return k

    @torch.no_grad()
    def density_and_clone(self, grad_threshold: float, scene_extent: float, grad: Optional[torch.Tensor] = None,
                          generator: Optional[torch.Generator] = None) -> int:
        """Small, high-gradient splats -> a jittered copy (gaussian_model.py:159-179)."""
        g = self._xyz.grad if grad is None else grad
        if g is None:
            return 0
        sig = self.get_scaling
        mask = (g.norm(dim=-1) > grad_threshold) & (sig.mean(dim=-1) < 0.01 * scene_extent)
        k = int(mask.sum())
        if k == 0:
            return 0
        noise = torch.randn(k, 3, generator=generator, device=self._xyz.device if generator is None else generator.device)
        jitter = noise.to(self._xyz.device) * (sig[mask].mean(dim=-1, keepdim=True) * 0.5)
        self._append_points(self._xyz.data[mask] + jitter, self._features_dc.data[mask], self._features_rest.data[mask],
                            self._scaling.data[mask], self._rotation.data[mask], self._opacity.data[mask])
        return k

    @torch.no_grad()
    def densify_fused(self, grad: torch.Tensor, grad_threshold: float, scene_extent: float, min_opacity: float = 0.01,
                      noise: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> dict:
        """Clone, split and prune in one plan + one apply pass on the device (gs_densify_plan /
        gs_densify_apply): same result, row for row, as density_and_clone -> density_and_split ->
        prune_points(opacity > min_opacity) driven by DensityController, with one host read (the four
        counts) instead of one per step.  `noise` [k,3] is the jitter of the k splats that meet the clone
        criterion, in index order; when omitted it is drawn AFTER the plan pass as randn(k, 3) -- the draw
        density_and_clone makes -- so the same generator places the clones where the sequential path does."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        n = self.get_num_points()
        dev = self._xyz.device
        if dev.type != "cuda":
            raise RuntimeError("densify_fused runs on CUDA tensors only")
        f32c = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731
        grad = f32c(grad)
        params = [f32c(p.data) for p in (self._xyz, self._features_dc, self._features_rest, self._scaling, self._rotation,
                                          self._opacity)]
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            ws_bytes = int(lib.gs_densify_workspace_bytes(n))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            counts = torch.empty(5, dtype=torch.int64, device=dev)
            _lib.check(lib.gs_densify_plan(n, _lib.ptr(params[3]), _lib.ptr(params[5]), _lib.ptr(grad), float(grad_threshold),
                                           float(0.01 * scene_extent), float(0.03 * scene_extent), float(min_opacity),
                                           _lib.ptr(ws), ws_bytes, _lib.ptr(counts), stream), "gs_densify_plan")
            kept, cloned, split, total, candidates = (int(v) for v in counts.tolist())
            if noise is None:
                noise = torch.randn(candidates, 3, generator=generator, device=dev if generator is None else generator.device)
            noise = f32c(noise.to(dev))
            if noise.shape[0] < candidates:
                raise ValueError(f"noise has {noise.shape[0]} rows, {candidates} splats meet the clone criterion")
            out = [torch.empty((total,) + tuple(p.shape[1:]), dtype=torch.float32, device=dev) for p in params]
            src_row = torch.empty(total, dtype=torch.int32, device=dev)
            _lib.check(lib.gs_densify_apply(n, _lib.ptr(ws), kept, cloned, split, *[_lib.ptr(p) for p in params], _lib.ptr(noise),
                                            *[_lib.ptr(o) for o in out], _lib.ptr(src_row), stream), "gs_densify_apply")
        self._set(*out)
        return {"kept": kept, "cloned": cloned, "split": split, "points": total, "src_row": src_row}

