"""Host-side math helpers with the reference's names (src/utils/math_utils.py:7-49).  Utilities for callers and tests:
the renderer itself evaluates all of this inside `project_fwd_kernel` (csrc/project.cu) and never calls into here."""
from __future__ import annotations

import torch

from .scene import quaternion_to_rotation

_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


class MathUtils:
    @staticmethod
    def build_rotation_matrix(quaternion: torch.Tensor) -> torch.Tensor:
        """[N,4] quaternions (w,x,y,z), normalised first -> [N,3,3] (math_utils.py:9-26)."""
        return quaternion_to_rotation(quaternion)

    @staticmethod
    def build_covariance_3d(scaling: torch.Tensor, rotation: torch.Tensor) -> torch.Tensor:
        """R diag(scaling^2) R^T with ACTIVATED scales (math_utils.py:28-34)."""
        R = quaternion_to_rotation(rotation)
        return (R * (scaling ** 2).unsqueeze(1)) @ R.transpose(-1, -2)

    @staticmethod
    def spherical_harmonics_eval(degrees: int, dirs: torch.Tensor, coeffs: torch.Tensor) -> torch.Tensor:
        """coeffs [N,K,3] evaluated at unit directions dirs [N,3] -> [N,3].  Degree 0 is the reference's behaviour
        (math_utils.py:44-49 returns coeffs[:,0], which is also what its renderer uses as the colour logit); degrees
        1..3 add the real spherical-harmonics rows 1..(degrees+1)^2-1 in the order and sign convention of the original
        3DGS code -- the same terms `GaussianRenderer(sh_degree=...)` evaluates on the device."""
        out = coeffs[:, 0]
        if degrees <= 0:
            return out
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        basis = [-_C1 * y, _C1 * z, -_C1 * x]
        if degrees >= 2:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            basis += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
        if degrees >= 3:
            basis += [_C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
                      _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
                      _C3[6] * x * (xx - 3 * yy)]
        for k, b in enumerate(basis):
            out = out + b * coeffs[:, k + 1]
        return out
