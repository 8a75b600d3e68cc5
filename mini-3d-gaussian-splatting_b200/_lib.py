"""ctypes binding of libgsplat_b200.so (include/gsplat_b200.h).

There is no fallback: if the library is missing or was built for another ABI the import of the
renderer fails with an explicit error.  Build it with `make` or `__graft_entry__.build()`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p

ABI_VERSION = 18
LIB_NAME = "libgsplat_b200.so"
# GSPLAT_B200_LIB points at an alternative build of the same ABI (kernel-variant experiments, tools/)
LIB_PATH = os.environ.get("GSPLAT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", LIB_NAME)

# symbol -> (restype, argtypes); mirrors include/gsplat_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "gs_abi_version": (c_int32, []),
    "gs_last_error_string": (c_char_p, []),
    "gs_built_for_sm": (c_int32, []),
    "gs_kernel_launch_count": (c_int64, []),
    "gs_project_fwd": (c_int32, [c_int64, _P, _P, _P, _P, _P, c_int32, _P, c_int64, _P, c_int64, c_int32, _P,
                                  c_int32, c_int32, c_int32, c_float, c_float,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gs_project_bwd": (c_int32, [c_int64, _P, _P, _P, _P, _P, c_int32, _P, c_int64, _P, c_int64, c_int32, _P,
                                  _P, _P, _P, _P, _P,
                                  _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int32, _P, _P, _P, _P, _P, _P]),
    "gs_densify_workspace_bytes": (c_int64, [c_int64]),
    "gs_densify_plan": (c_int32, [c_int64, _P, _P, _P, c_float, c_float, c_float, c_float, _P, c_int64, _P, _P]),
    "gs_densify_apply": (c_int32, [c_int64, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P, _P, _P, _P]),
    "gs_peer_allreduce": (c_int32, [_P, ctypes.c_uint64, c_int32, c_int32, c_int64, c_int64, c_int64, c_int64, c_int32, _P]),
    "gs_bin_workspace_bytes": (c_int64, [c_int64, c_int64, c_int32]),
    "gs_bin_prepare": (c_int32, [c_int64, _P, _P, _P, c_int64, _P, _P, _P, _P]),
    "gs_bin_sort": (c_int32, [c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int32, c_int32, c_int32,
                               _P, c_int64, _P, _P, _P, _P, c_int32, _P, _P, _P]),
    "gs_bin_complete": (c_int32, [c_int64, c_int64, c_int64, _P, _P, _P, c_int32, c_int32, _P, c_int64,
                                   c_int32, _P, _P, _P, _P, _P]),
    "gs_raster_fwd": (c_int32, [c_int32, c_int32, c_int32, _P, _P, _P, _P, c_int32, _P, _P,
                                 c_int32, _P, _P, c_int32,
                                 _P, _P, _P, _P, _P, _P, _P]),
    "gs_raster_bwd": (c_int32, [c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, c_int32,
                                 _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gs_tile_order": (c_int32, [c_int32, _P, _P, _P, _P]),
    "gs_loss_workspace_bytes": (c_int64, []),
    "gs_weighted_sum": (c_int32, [c_int32, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "gs_l1_loss": (c_int32, [_P, _P, c_int64, c_float, _P, _P, _P, c_int64, _P]),
}


class GsplatLibraryError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library once; raise loudly if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsplatLibraryError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run `make` at the repo root "
            f"(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover - build/ABI mismatch
            raise GsplatLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    got = lib.gs_abi_version()
    if got != ABI_VERSION:
        raise GsplatLibraryError(f"{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}; rebuild with `make`")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().gs_last_error_string()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def ptr(t) -> c_void_p:
    """Device pointer of a tensor (None -> NULL)."""
    return c_void_p(0 if t is None else t.data_ptr())
