"""CPU oracle for the GaussianRenderer.render hot path -- TEST INFRASTRUCTURE ONLY.

This module is a from-scratch restatement, in vectorised torch-CPU fp32, of what the
reference renderer computes (Loveof1ife7/mini-3d-gaussian-splatting,
src/core/renderer.py:31-367 plus src/core/gaussian_model.py:101-122,200-207 and
src/utils/math_utils.py:9-26).  It exists so that the CUDA path can be checked on the GPU
box, where /root/reference is not available.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product package never does: it fails loudly when its CUDA library is absent.

Parity pin: every stage below is compared against the *literal* reference in
tests/golden/make_golden.py (run in the build container, where /root/reference is importable);
the resulting fixtures are committed under tests/golden/ and re-checked by the CPU test-suite.

Differences from the reference are deliberate and limited to:
  * the per-pixel Python loops (renderer.py:302-355) are evaluated for all pixels of a tile at
    once, keeping the *same sequential recurrence* and the same fp32 operation order;
  * the depth sort is stable (ties -> ascending Gaussian index); the reference's argsort
    (renderer.py:235) leaves tie order unspecified.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

F32 = torch.float32
F64 = torch.float64


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _fma(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """fp32 fused multiply-add, emulated through fp64 (the fp32 x fp32 product is exact in
    fp64).  Matches the FMA accumulation MKL sgemm uses for the K=3 products in the
    reference (verified bit-for-bit in make_golden.py)."""
    return (a.to(F64) * b.to(F64) + c.to(F64)).to(F32)


def _dot3_fma(x0, x1, x2, r0, r1, r2):
    """x0*r0, then two FMAs -- the order MKL uses for `Xw @ Rv.T` (renderer.py:154)."""
    acc = x0 * r0
    acc = _fma(x1, r1, acc)
    acc = _fma(x2, r2, acc)
    return acc


@dataclass
class OracleCamera:
    """What the reference reads from a camera: renderer.py:140-152."""
    width: int
    height: int
    fovx: float
    fovy: float
    world_view: torch.Tensor  # [4,4] world->camera

    def intrinsics(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        # python float64 arithmetic, then one cast to fp32 (renderer.py:142-147)
        fx = torch.tensor(0.5 * self.width / math.tan(self.fovx * 0.5), dtype=F32)
        fy = torch.tensor(0.5 * self.height / math.tan(self.fovy * 0.5), dtype=F32)
        cx = torch.tensor(self.width * 0.5, dtype=F32)
        cy = torch.tensor(self.height * 0.5, dtype=F32)
        return fx, fy, cx, cy


# --------------------------------------------------------------------------------------
# row M: activations + 3D covariance  (gaussian_model.py:101-122,200-207; math_utils.py:9-26)
# --------------------------------------------------------------------------------------
def rotation_matrix(rotation: torch.Tensor) -> torch.Tensor:
    """[N,4] (w,x,y,z), not necessarily unit -> [N,3,3].  The reference normalises in
    get_rotation and again inside build_rotation_matrix; both are kept."""
    q = torch.nn.functional.normalize(rotation, dim=-1)
    q = torch.nn.functional.normalize(q, dim=-1)
    w, x, y, z = q.unbind(-1)
    xx, yy, zz = x * x, y * y, z * z
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    rows = [1 - 2 * (yy + zz), 2 * (xy - wz), 2 * (xz + wy),
            2 * (xy + wz), 1 - 2 * (xx + zz), 2 * (yz - wx),
            2 * (xz - wy), 2 * (yz + wx), 1 - 2 * (xx + yy)]
    return torch.stack(rows, dim=-1).view(-1, 3, 3)


def covariance_3d(scaling_log: torch.Tensor, rotation: torch.Tensor) -> torch.Tensor:
    """Sigma = R diag(exp(s)^2) R^T  (gaussian_model.py:200-207)."""
    sig = torch.exp(scaling_log)
    R = rotation_matrix(rotation)
    d = sig * sig                                    # [N,3]
    # (R * d) @ R^T written out: Sigma_ij = sum_k R_ik d_k R_jk
    RD = R * d.unsqueeze(1)
    return RD @ R.transpose(-1, -2)


# --------------------------------------------------------------------------------------
# row P: projection (renderer.py:117-200)
# --------------------------------------------------------------------------------------
def project(xyz: torch.Tensor, cov3d: torch.Tensor, cam: OracleCamera,
            radius_min: float = 0.01, radius_max: float = 50.0) -> Dict[str, torch.Tensor]:
    fx, fy, cx, cy = cam.intrinsics()
    WV = cam.world_view.to(F32)
    Rv, Tv = WV[:3, :3], WV[:3, 3]

    x0, x1, x2 = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    X = _dot3_fma(x0, x1, x2, Rv[0, 0], Rv[0, 1], Rv[0, 2]) + Tv[0]
    Y = _dot3_fma(x0, x1, x2, Rv[1, 0], Rv[1, 1], Rv[1, 2]) + Tv[1]
    Z = _dot3_fma(x0, x1, x2, Rv[2, 0], Rv[2, 1], Rv[2, 2]) + Tv[2]

    xpix = fx * X / Z + cx                       # ((fx*X)/Z)+cx          renderer.py:161
    ypix = -fy * Y / Z + cy                      # (((-fy)*Y)/Z)+cy      renderer.py:162
    means2D = torch.stack([xpix, ypix], dim=-1)

    # Sigma_cam = Rv Sigma Rv^T                                             renderer.py:168
    cov_cam = Rv @ cov3d @ Rv.T
    invZ = 1.0 / Z
    j00 = fx * invZ
    j02 = -fx * X * invZ * invZ
    j11 = -fy * invZ
    j12 = fy * Y * invZ * invZ
    # J assembled without in-place writes so autograd stays simple (renderer.py:171-177)
    zero = torch.zeros_like(j00)
    J = torch.stack([torch.stack([j00, zero, j02], -1), torch.stack([zero, j11, j12], -1)], -2)
    cov2d = J @ cov_cam @ J.transpose(-1, -2)
    cov2d = cov2d + torch.eye(2, dtype=F32).unsqueeze(0) * 1e-6            # renderer.py:182-183

    conics = torch.linalg.inv(cov2d)                                        # renderer.py:186
    lam = torch.linalg.eigvalsh(cov2d)[:, 1]                                # renderer.py:188
    radii = (3.0 * torch.sqrt(lam)).clamp(radius_min, radius_max)           # renderer.py:190-192
    return {"means2D": means2D, "cov2D": cov2d, "conics": conics, "depths": Z, "radii": radii}


def radii_closed_form(cov2d: torch.Tensor, radius_min=0.01, radius_max=50.0) -> torch.Tensor:
    """lambda_max = (a+c)/2 + sqrt(((a-c)/2)^2 + b^2): what the CUDA kernel evaluates instead
    of LAPACK eigvalsh (int(radii) identical, float within 2 ulp)."""
    a, b, c = cov2d[:, 0, 0], cov2d[:, 0, 1], cov2d[:, 1, 1]
    mid = 0.5 * (a + c)
    hd = 0.5 * (a - c)
    lam = mid + torch.sqrt(hd * hd + b * b)
    return (3.0 * torch.sqrt(lam)).clamp(radius_min, radius_max)


# --------------------------------------------------------------------------------------
# row C: culling (renderer.py:201-220)   row S: depth sort (renderer.py:222-239)
# --------------------------------------------------------------------------------------
def cull(means2D, depths, radii, H: int, W: int) -> torch.Tensor:
    x, y = means2D[:, 0], means2D[:, 1]
    return (depths > 0) & (x >= -radii) & (x < W + radii) & (y >= -radii) & (y < H + radii) & (radii > 0)


def sort_by_depth(vis: torch.Tensor, depths: torch.Tensor) -> torch.Tensor:
    idx = torch.nonzero(vis, as_tuple=False).flatten()
    order = torch.argsort(depths[idx].detach(), stable=True)
    return idx[order]


# --------------------------------------------------------------------------------------
# row B: tile binning (renderer.py:263-298)
# --------------------------------------------------------------------------------------
def tile_rects(means2D, radii, H: int, W: int, T: int = 16):
    """Integer AABB -> inclusive tile rectangle per Gaussian.  `.to(int)` truncates toward
    zero exactly as Python int() does (renderer.py:278-286).  count==0 <=> empty AABB."""
    m = means2D.detach()
    big = 2 ** 30   # guard: inf/NaN centres are never visible; keep the casts defined
    ix = torch.nan_to_num(m[:, 0], nan=0.0, posinf=big, neginf=-big).clamp(-big, big).to(torch.int64)
    iy = torch.nan_to_num(m[:, 1], nan=0.0, posinf=big, neginf=-big).clamp(-big, big).to(torch.int64)
    ir = radii.detach().to(torch.int64)
    x0 = (ix - ir).clamp(min=0)
    x1 = (ix + 1 + ir).clamp(max=W)
    y0 = (iy - ir).clamp(min=0)
    y1 = (iy + 1 + ir).clamp(max=H)
    empty = (x0 >= x1) | (y0 >= y1)
    tx0 = x0 // T
    tx1 = (x1 - 1) // T
    ty0 = y0 // T
    ty1 = (y1 - 1) // T
    cnt = torch.where(empty, torch.zeros_like(tx0), (tx1 - tx0 + 1) * (ty1 - ty0 + 1))
    return tx0, tx1, ty0, ty1, cnt


def depth_key_bits(depths: torch.Tensor) -> torch.Tensor:
    """fp32 bit pattern as uint32 held in int64; order-preserving for Z > 0."""
    return depths.detach().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF


def bin_tiles(sorted_ids, means2D, radii, depths, H: int, W: int, T: int = 16):
    """Per-tile lists in global depth order, as flat arrays.

    Returns (keys[D] int64 = tile_id<<32 | depth_bits, ids[D] int64, ranges[num_tiles,2]).
    Equivalent to the append loop at renderer.py:277-298.
    """
    tiles_x = math.ceil(W / T)
    tiles_y = math.ceil(H / T)
    tx0, tx1, ty0, ty1, cnt = tile_rects(means2D, radii, H, W, T)
    ids = sorted_ids
    c = cnt[ids]
    D = int(c.sum())
    rank_of_entry = torch.repeat_interleave(torch.arange(ids.numel()), c)
    start = torch.cumsum(c, 0) - c
    k = torch.arange(D) - start[rank_of_entry]
    g = ids[rank_of_entry]
    wdt = (tx1 - tx0 + 1)[g]
    tile = (ty0[g] + k // wdt) * tiles_x + (tx0[g] + k % wdt)
    order = torch.argsort(tile, stable=True)          # stable: keeps depth order within a tile
    tile_s, g_s = tile[order], g[order]
    keys = (tile_s << 32) | depth_key_bits(depths)[g_s]
    counts = torch.bincount(tile_s, minlength=tiles_x * tiles_y)
    ends = torch.cumsum(counts, 0)
    ranges = torch.stack([ends - counts, ends], dim=1)
    return keys, g_s, ranges


# --------------------------------------------------------------------------------------
# row R: compositing (renderer.py:300-367), all pixels of a tile at once
# --------------------------------------------------------------------------------------
def composite_tile(tid: int, ids_list, ranges, means2D, conics, qs_all, depths, colors, opacities, bg,
                   H: int, W: int, T: int = 16):
    """One tile of the reference pixel loop (renderer.py:302-355), all its pixels at once.

    `alive` plays the role of the `break` at renderer.py:352, `act` the three `continue`s at
    :336,:340,:345.  Returns (C[3,P] incl. the initial bg, A[P], Dsum[P], n_consumed[P], (y0,y1,x0,x1)).
    """
    tiles_x = math.ceil(W / T)
    ty, tx = divmod(tid, tiles_x)
    x0, x1 = tx * T, min(tx * T + T, W)
    y0, y1 = ty * T, min(ty * T + T, H)
    ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
    xs = xs.reshape(-1).to(F32)
    ys = ys.reshape(-1).to(F32)
    P = xs.numel()
    A = torch.zeros(P, dtype=F32)
    C = bg.view(3, 1).expand(3, P).clone()          # out_rgb starts at bg (:273)
    Ds = torch.zeros(P, dtype=F32)
    alive = torch.ones(P, dtype=torch.bool)
    ncons = torch.zeros(P, dtype=torch.int64)
    s0, s1 = int(ranges[tid, 0]), int(ranges[tid, 1])
    for pos in range(s0, s1):
        if not bool(alive.any()):
            break
        i = ids_list[pos]
        dx = xs - means2D[i, 0]
        dy = ys - means2D[i, 1]
        s = dx * dx * conics[i, 0, 0] + qs_all[i] * dx * dy + dy * dy * conics[i, 1, 1]
        w = torch.exp(-0.5 * s).clamp(0.0, 1.0)
        act = alive & ~(w < 1e-5)
        a = (opacities[i] * w).clamp(0.0, 1.0)
        act = act & ~(a <= 0.0)
        contrib = (1.0 - A) * a
        act = act & ~(contrib <= 0.0)
        C = torch.where(act.unsqueeze(0), C + contrib.unsqueeze(0) * colors[i].view(3, 1), C)
        A = torch.where(act, A + contrib, A)
        Ds = torch.where(act, Ds + contrib * depths[i], Ds)
        ncons = torch.where(alive, torch.full_like(ncons, pos - s0 + 1), ncons)
        alive = alive & ~(act & (A.detach() >= 0.995))
    return C, A, Ds, ncons, (y0, y1, x0, x1)


def finish_pixels(C, A, Ds, bg):
    """Epilogue renderer.py:359-367 on flat [3,P]/[P] tensors: background a second time, depth
    normalisation, clamps."""
    rgb = C + (1.0 - A).unsqueeze(0) * bg.view(3, 1)
    return rgb.clamp(0, 1), A.clamp(0, 1), Ds / (A + 1e-6)


def rasterize(ids_flat, ranges, means2D, conics, depths, colors, opacities, bg,
              H: int, W: int, T: int = 16, return_stats: bool = False):
    """Differentiable (torch autograd) restatement of the reference tile loop; fp32 operation
    order is the reference's."""
    tiles_x = math.ceil(W / T)
    tiles_y = math.ceil(H / T)
    bg = bg.to(F32).view(3)
    rgb_rows: List[List[torch.Tensor]] = []
    a_rows: List[List[torch.Tensor]] = []
    d_rows: List[List[torch.Tensor]] = []
    n_consumed = torch.zeros((H, W), dtype=torch.int64)
    tile_consumed = torch.zeros(tiles_x * tiles_y, dtype=torch.int64)
    ids_list = ids_flat.tolist()
    qs_all = conics[:, 0, 1] + conics[:, 1, 0]
    for ty in range(tiles_y):
        rgb_row, a_row, d_row = [], [], []
        for tx in range(tiles_x):
            tid = ty * tiles_x + tx
            C, A, Ds, ncons, (y0, y1, x0, x1) = composite_tile(tid, ids_list, ranges, means2D, conics, qs_all,
                                                               depths, colors, opacities, bg, H, W, T)
            tile_consumed[tid] = int(ncons.max()) if ncons.numel() else 0
            n_consumed[y0:y1, x0:x1] = ncons.view(y1 - y0, x1 - x0)
            rgb_row.append(C.view(3, y1 - y0, x1 - x0))
            a_row.append(A.view(1, y1 - y0, x1 - x0))
            d_row.append(Ds.view(1, y1 - y0, x1 - x0))
        rgb_rows.append(torch.cat(rgb_row, dim=2))
        a_rows.append(torch.cat(a_row, dim=2))
        d_rows.append(torch.cat(d_row, dim=2))
    out_rgb = torch.cat(rgb_rows, dim=1)
    out_a = torch.cat(a_rows, dim=1)
    out_d = torch.cat(d_rows, dim=1)
    out_rgb = out_rgb + (1.0 - out_a) * bg.view(3, 1, 1)      # background a second time (:359)
    out_d = out_d / (out_a + 1e-6)                            # :362
    res = {"image": out_rgb.clamp(0, 1), "alpha": out_a.clamp(0, 1), "depth": out_d}
    if return_stats:
        res["n_consumed"] = n_consumed
        res["tile_consumed"] = tile_consumed
    return res


# --------------------------------------------------------------------------------------
# row O: the whole render() (renderer.py:31-114)
# --------------------------------------------------------------------------------------
def render(cam: OracleCamera, xyz, cov3d, features0, opacity_act, bg, H: int, W: int,
           tile_size: int = 16, radius_min: float = 0.01, radius_max: float = 50.0,
           return_stats: bool = False) -> Dict[str, torch.Tensor]:
    """`features0` = get_features[:,0,:] *before* the sigmoid (renderer.py:88-92);
    `opacity_act` = get_opacity.squeeze(1), already activated (renderer.py:94)."""
    proj = project(xyz, cov3d, cam, radius_min, radius_max)
    means2D, conics, depths, radii = proj["means2D"], proj["conics"], proj["depths"], proj["radii"]
    vis = cull(means2D, depths, radii, H, W)
    bgv = bg.to(F32).view(3, 1, 1)
    if int(vis.sum()) == 0:                                                  # renderer.py:74-83
        out = {"image": bgv.repeat(1, H, W), "alpha": torch.zeros((1, H, W)),
               "depth": torch.zeros((1, H, W))}
    else:
        sorted_ids = sort_by_depth(vis, depths)
        keys, ids_flat, ranges = bin_tiles(sorted_ids, means2D, radii, depths, H, W, tile_size)
        colors = torch.sigmoid(features0)
        out = rasterize(ids_flat, ranges, means2D, conics, depths, colors, opacity_act, bg,
                        H, W, tile_size, return_stats=return_stats)
        out["sort_keys"], out["sort_ids"], out["tile_ranges"] = keys, ids_flat, ranges
    out.update({"viewspace_points": means2D, "visibility_filter": vis, "radii": radii,
                "conics": conics, "depths": depths, "cov2D": proj["cov2D"]})
    return out


def sh_basis(degree: int, d: torch.Tensor) -> torch.Tensor:
    """Real spherical-harmonics basis Y_1..Y_{(deg+1)^2-1} at unit directions d [N,3] -> [N,terms].
    NOT part of the reference (its colour is DC-only; math_utils.py:44-49 is a stub): this is the
    oracle of the renderer's optional `sh_degree` extension, so its parity is unpinned by the reference."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    out = [-0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x]
    if degree >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.31539156525252005 * (2 * zz - xx - yy),
                -1.0925484305920792 * xz, 0.5462742152960396 * (xx - yy)]
    if degree >= 3:
        out += [-0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * xy * z,
                -0.4570457994644658 * y * (4 * zz - xx - yy), 0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy),
                -0.4570457994644658 * x * (4 * zz - xx - yy), 1.445305721320277 * z * (xx - yy),
                -0.5900435899266435 * x * (xx - 3 * yy)]
    return torch.stack(out, dim=1)


def colour_logits(cam: OracleCamera, xyz, features_dc, features_rest=None, sh_degree: int = 0) -> torch.Tensor:
    """What goes into the sigmoid: features[:,0,:] (renderer.py:88-92), plus -- extension -- the SH terms."""
    f0 = features_dc.reshape(-1, 3)
    if sh_degree <= 0 or features_rest is None:
        return f0
    WV = cam.world_view.to(torch.float64)
    centre = (-(WV[:3, :3].T @ WV[:3, 3])).to(F32)
    d = torch.nn.functional.normalize(xyz - centre, dim=-1)
    Y = sh_basis(sh_degree, d)                                        # [N, terms]
    return f0 + (Y.unsqueeze(-1) * features_rest[:, :Y.shape[1], :]).sum(dim=1)


def render_from_params(cam: OracleCamera, xyz, scaling_log, rotation, opacity_logit, features_dc,
                       bg, H: int, W: int, features_rest=None, sh_degree: int = 0, **kw) -> Dict[str, torch.Tensor]:
    """render() fed the way a real GaussianModel would feed it (adapter of SURVEY 8c)."""
    cov3d = covariance_3d(scaling_log, rotation)
    op = torch.sigmoid(opacity_logit).reshape(-1)
    f0 = colour_logits(cam, xyz, features_dc, features_rest, sh_degree)
    return render(cam, xyz, cov3d, f0, op, bg, H, W, **kw)


# --------------------------------------------------------------------------------------
# synthetic scenes and cameras of SURVEY 8d (shared by tests and bench so inputs are identical)
# --------------------------------------------------------------------------------------
def scene_ref_init(n: int, seed: int = 0, scene_extent: float = 1.0) -> Dict[str, torch.Tensor]:
    """Same draws, in the same order, as GaussianModel.create_from_random
    (gaussian_model.py:78-98) on the CPU generator."""
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(n, 3, generator=g) - 0.5) * (2.0 * scene_extent)
    features_dc = torch.rand(n, 1, 3, generator=g)
    features_rest = torch.zeros(n, 15, 3)
    scaling = torch.full((n, 3), math.log(0.02 * scene_extent))
    rot = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    opacity = torch.full((n, 1), -2.0)
    return {"xyz": xyz, "features_dc": features_dc, "features_rest": features_rest,
            "scaling": scaling, "rotation": rot, "opacity": opacity, "_gen": g}


def scene_aniso(n: int, seed: int = 0, scene_extent: float = 1.0) -> Dict[str, torch.Tensor]:
    s = scene_ref_init(n, seed, scene_extent)
    g = s["_gen"]
    s["scaling"] = s["scaling"] + 0.5 * torch.randn(n, 3, generator=g)
    s["opacity"] = s["opacity"] + 1.5 * torch.randn(n, 1, generator=g)
    s["features_dc"] = 1.5 * torch.randn(n, 1, 3, generator=g)
    return s


def camera_c0(W: int, H: int, fovx_deg: float = 60.0, square_pixels: bool = True) -> OracleCamera:
    fovx = math.radians(fovx_deg)
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W) if square_pixels else fovx
    WV = torch.eye(4, dtype=F32)
    WV[2, 3] = 3.0
    return OracleCamera(W, H, fovx, fovy, WV)


def camera_orbit(k: int, M: int, W: int, H: int, fovx_deg: float = 60.0) -> OracleCamera:
    th = 2.0 * math.pi * k / M
    C = 3.0 * np.array([math.sin(th), 0.3, -math.cos(th)])
    f = -C / np.linalg.norm(C)
    r = np.cross(np.array([0.0, 1.0, 0.0]), f)
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    R = np.stack([r, u, f])
    t = -R @ C
    WV = torch.eye(4, dtype=F32)
    WV[:3, :3] = torch.tensor(R, dtype=F32)
    WV[:3, 3] = torch.tensor(t, dtype=F32)
    fovx = math.radians(fovx_deg)
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
    return OracleCamera(W, H, fovx, fovy, WV)


def loss_weights(H: int, W: int, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(3, H, W, generator=g), torch.rand(1, H, W, generator=g),
            torch.rand(1, H, W, generator=g))


def weighted_loss(out, weights):
    wi, wa, wd = weights
    return (wi * out["image"]).sum() + (wa * out["alpha"]).sum() + 0.1 * (wd * out["depth"]).sum()
