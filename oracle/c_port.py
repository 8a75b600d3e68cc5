"""numpy/ctypes front-end of the plain-C oracle (oracle/splat_oracle.c -> oracle/_ref/liboracle.so).
TEST INFRASTRUCTURE ONLY -- see the header of splat_oracle.c.  Built by `make oracle`."""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from ctypes import c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "liboracle.so")
_lib = None


def build() -> None:
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-fopenmp",
                    "-o", LIB_PATH, os.path.join(HERE, "splat_oracle.c"), "-lm"], check=True)


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "splat_oracle.c")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_bin.restype = c_int64
        _lib.oracle_num_threads.restype = c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def camera_block(width, height, fovx, fovy, world_view) -> np.ndarray:
    wv = np.asarray(world_view, dtype=np.float32)
    fx = np.float32(0.5 * width / math.tan(fovx * 0.5))
    fy = np.float32(0.5 * height / math.tan(fovy * 0.5))
    return np.concatenate([wv[:3, :3].reshape(-1), wv[:3, 3], [fx, fy, np.float32(width * 0.5), np.float32(height * 0.5)]]).astype(np.float32)


def num_threads() -> int:
    return int(load().oracle_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the port from now on (a launcher such as torchrun exports OMP_NUM_THREADS=1)."""
    load().oracle_set_num_threads(c_int(int(n)))
    return num_threads()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:      # pragma: no cover
        return os.cpu_count() or 1


def project(cam16, W, H, xyz, scaling=None, rotation=None, cov3d=None, opacity=None, opacity_is_logit=True, feat0=None,
            rmin=0.01, rmax=50.0):
    lib = load()
    xyz, scaling, rotation, cov3d = _f32(xyz), _f32(scaling), _f32(rotation), _f32(cov3d)
    opacity = _f32(opacity).reshape(-1)
    feat0 = _f32(feat0).reshape(xyz.shape[0], -1)
    n = xyz.shape[0]
    out = dict(means2D=np.empty((n, 2), np.float32), depths=np.empty(n, np.float32), conics=np.empty((n, 2, 2), np.float32),
               radii=np.empty(n, np.float32), colors=np.empty((n, 3), np.float32), opacities=np.empty(n, np.float32),
               vis=np.empty(n, np.uint8), tiles_touched=np.empty(n, np.int32), rect=np.empty((n, 4), np.int32))
    lib.oracle_project(c_int64(n), _p(xyz), _p(scaling), _p(rotation), _p(cov3d), _p(opacity), c_int(int(opacity_is_logit)),
                       _p(feat0), c_int64(feat0.shape[1]), _p(cam16), c_int(W), c_int(H), c_float(rmin), c_float(rmax),
                       _p(out["means2D"]), _p(out["depths"]), _p(out["conics"]), _p(out["radii"]), _p(out["colors"]),
                       _p(out["opacities"]), _p(out["vis"]), _p(out["tiles_touched"]), _p(out["rect"]))
    out["_inputs"] = (xyz, scaling, rotation, cov3d, opacity, bool(opacity_is_logit), feat0, cam16)
    return out


def bin_tiles(proj, W, H):
    lib = load()
    n = proj["depths"].shape[0]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    sorted_ids = np.empty(max(n, 1), np.int32)
    ns = c_int64(0)
    D = int(proj["tiles_touched"].astype(np.int64).sum())
    entry_ids = np.empty(max(D, 1), np.int32)
    ranges = np.zeros((tiles, 2), np.int32)
    d2 = lib.oracle_bin(c_int64(n), _p(proj["depths"]), _p(proj["tiles_touched"]), _p(proj["rect"]), c_int(W), c_int(H),
                        _p(sorted_ids), ctypes.byref(ns), _p(entry_ids), _p(ranges))
    assert d2 == D
    return sorted_ids[:ns.value], entry_ids[:D], ranges


def raster_fwd(proj, entry_ids, ranges, bg, W, H, tile_first=0, tile_count=None, any_visible=True):
    lib = load()
    tiles = ranges.shape[0]
    tile_count = tiles - tile_first if tile_count is None else tile_count
    bg = _f32(bg)
    out = dict(image=np.zeros((3, H, W), np.float32), alpha=np.zeros((1, H, W), np.float32), depth=np.zeros((1, H, W), np.float32),
               pix_state=np.zeros((H * W, 4), np.float32), n_consumed=np.zeros((H, W), np.int32))
    lib.oracle_raster_fwd(c_int(W), c_int(H), _p(entry_ids), _p(ranges), _p(proj["means2D"]), _p(proj["conics"]),
                          _p(proj["depths"]), _p(proj["colors"]), _p(proj["opacities"]), _p(bg), c_int(int(any_visible)),
                          c_int(tile_first), c_int(tile_count), _p(out["image"]), _p(out["alpha"]), _p(out["depth"]),
                          _p(out["pix_state"]), _p(out["n_consumed"]))
    return out


def raster_bwd(proj, entry_ids, ranges, bg, W, H, g_image, g_alpha, g_depth, tile_first=0, tile_count=None):
    lib = load()
    tiles = ranges.shape[0]
    tile_count = tiles - tile_first if tile_count is None else tile_count
    n = proj["depths"].shape[0]
    g = dict(means2D=np.zeros((n, 2)), conics=np.zeros((n, 2, 2)), depths=np.zeros(n), colors=np.zeros((n, 3)), opacities=np.zeros(n))
    lib.oracle_raster_bwd(c_int(W), c_int(H), _p(entry_ids), _p(ranges), _p(proj["means2D"]), _p(proj["conics"]), _p(proj["depths"]),
                          _p(proj["colors"]), _p(proj["opacities"]), _p(_f32(bg)), c_int(tile_first), c_int(tile_count),
                          _p(_f32(g_image)), _p(_f32(g_alpha)), _p(_f32(g_depth)),
                          _p(g["means2D"]), _p(g["conics"]), _p(g["depths"]), _p(g["colors"]), _p(g["opacities"]))
    return g


def project_bwd(proj, g):
    lib = load()
    xyz, scaling, rotation, cov3d, opacity, is_logit, feat0, cam16 = proj["_inputs"]
    n = xyz.shape[0]
    out = dict(xyz=np.zeros((n, 3)), scaling=np.zeros((n, 3)), rotation=np.zeros((n, 4)),
               cov3d=np.zeros((n, 3, 3)) if cov3d is not None else None, opacity=np.zeros(n), feat0=np.zeros((n, 3)))
    lib.oracle_project_bwd(c_int64(n), _p(xyz), _p(scaling), _p(rotation), _p(cov3d), _p(opacity), c_int(int(is_logit)), _p(feat0),
                           c_int64(feat0.shape[1]), _p(cam16), _p(g["means2D"]), _p(g["conics"]), _p(g["depths"]), _p(g["colors"]),
                           _p(g["opacities"]), _p(out["xyz"]), _p(out["scaling"]), _p(out["rotation"]), _p(out["cov3d"]),
                           _p(out["opacity"]), _p(out["feat0"]))
    return out


def render_fwd_bwd(cam16, W, H, params, bg, weights, tile_first=0, tile_count=None, backward=True):
    """Whole path on the CPU port: project -> bin -> raster (-> raster_bwd -> project_bwd) with the
    SURVEY 8d fixed-weight loss  sum(wi*image) + sum(wa*alpha) + 0.1 sum(wd*depth)."""
    proj = project(cam16, W, H, params["xyz"], params["scaling"], params["rotation"], None, params["opacity"], True,
                   np.asarray(params["features_dc"]).reshape(-1, 3))
    sorted_ids, entry_ids, ranges = bin_tiles(proj, W, H)
    fwd = raster_fwd(proj, entry_ids, ranges, bg, W, H, tile_first, tile_count, any_visible=bool(proj["vis"].any()))
    res = {"proj": proj, "sorted_ids": sorted_ids, "entry_ids": entry_ids, "ranges": ranges, **fwd}
    if backward:
        wi, wa, wd = (np.asarray(w, dtype=np.float32) for w in weights)
        g = raster_bwd(proj, entry_ids, ranges, bg, W, H, wi, wa, 0.1 * wd, tile_first, tile_count)
        res["g_raster"] = g
        res["grads"] = project_bwd(proj, g)
    return res
