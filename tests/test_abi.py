"""CPU: the C-ABI library loads and exports every symbol include/gsplat_b200.h declares, and the
host-side mirror refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "gsplat_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    import gsplat_b200  # noqa: F401  (loads the library or raises)
    from importlib import import_module
    _lib = import_module("mini-3d-gaussian-splatting_b200._lib")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_abi_version_and_build_target():
    from importlib import import_module
    _lib = import_module("mini-3d-gaussian-splatting_b200._lib")
    lib = _lib.load()
    assert lib.gs_abi_version() == _lib.ABI_VERSION
    assert lib.gs_built_for_sm() == 100
    m = re.search(r"#define GS_ABI_VERSION (\d+)", open(HEADER).read())
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_argument_validation_without_gpu():
    """Entry points validate before touching the device: no compute happens here."""
    from importlib import import_module
    _lib = import_module("mini-3d-gaussian-splatting_b200._lib")
    lib = _lib.load()
    cam = (ctypes.c_float * 20)()
    rc = lib.gs_project_fwd(8, None, None, None, None, None, 0, None, 3, None, 0, 0, cam, 64, 64, 0, 0.01, 50.0,
                            None, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"tile_size" in lib.gs_last_error_string()       # GS_ERR_INVALID_ARGUMENT (any size >= 1 is accepted)
    rc = lib.gs_project_fwd(8, None, None, None, None, None, 0, None, 3, None, 0, 0, cam, 64, 64, 16, 0.01, 50.0,
                            None, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1                                                       # GS_ERR_INVALID_ARGUMENT
    assert lib.gs_bin_workspace_bytes(1000, 50000, 64) > 0
    assert lib.gs_bin_workspace_bytes(-1, 0, 64) < 0


def test_renderer_refuses_cpu_tensors():
    import gsplat_b200 as gb
    m = gb.GaussianModel(device="cpu")
    m.create_from_random(16, seed=0)
    rd = gb.GaussianRenderer()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rd.render(gb.Camera.look_at_origin_c0(32, 32), m, gb.RenderSettings(32, 32, torch.zeros(3)))


def _call(lib, sigs, name, **by_index):
    """Call `name` with zeros / NULLs everywhere except the positions given as a{index}=value."""
    args = []
    for i, t in enumerate(sigs[name][1]):
        v = by_index.get(f"a{i}")
        if v is None:
            v = None if t is ctypes.c_void_p else 0
        args.append(v)
    rc = getattr(lib, name)(*args)
    return rc, (lib.gs_last_error_string() or b"").decode()


def test_every_entry_point_rejects_bad_arguments_before_touching_the_device():
    """GS_ERR_INVALID_ARGUMENT (-1) with a message naming the problem, for every stage of the ABI: sizes, NULL arrays,
    the tile grid, the truncated-list contract, the exchange's extents.  Nothing here reaches a CUDA call (this box has
    no device), which is the point: validation comes first."""
    from importlib import import_module
    _lib = import_module("mini-3d-gaussian-splatting_b200._lib")
    lib, sigs = _lib.load(), _lib.SIGNATURES
    cases = [
        ("gs_raster_fwd", dict(a0=64, a1=64, a2=16), "NULL array"),
        ("gs_raster_bwd", dict(a0=64, a1=64, a2=16), ""),
        ("gs_tile_order", dict(a0=0), "bad arguments"),
        ("gs_tile_order", dict(a0=16), "bad arguments"),                              # neither estimate given
        ("gs_bin_prepare", dict(a0=-1), "n < 0"),
        ("gs_bin_prepare", dict(a0=5), "counters is NULL"),
        ("gs_bin_sort", dict(a0=-1), "bad sizes"),
        ("gs_bin_sort", dict(a0=5, a1=6), "bad sizes"),                               # more sorted splats than splats
        ("gs_bin_sort", dict(a0=5, a1=5, a2=10, a7=4, a8=0), "bad tile grid"),
        ("gs_bin_sort", dict(a0=5, a1=5, a2=10, a7=4, a8=16), "tile_ranges is NULL"),
        ("gs_bin_complete", dict(a0=5, a1=5, a2=10, a6=4, a7=16, a10=0), "gs_bin_complete needs"),
        ("gs_weighted_sum", dict(a0=0), "1..4 terms"),
        ("gs_weighted_sum", dict(a0=5), "1..4 terms"),
        ("gs_weighted_sum", dict(a0=2), "NULL argument"),
        ("gs_l1_loss", dict(a2=100), "bad arguments"),
        ("gs_densify_plan", dict(a0=-1), "n < 0"),
        ("gs_densify_plan", dict(a0=10), "counts is NULL"),
        ("gs_densify_apply", dict(a0=10, a2=-1), "negative size"),
        ("gs_peer_allreduce", dict(a2=2, a3=0), "peer_ptrs_host is NULL"),
    ]
    for name, over, needle in cases:
        rc, msg = _call(lib, sigs, name, **over)
        assert rc == -1, (name, over, rc, msg)
        assert needle in msg and name in msg, (name, over, msg)
    # the exchange: world / rank / extents are checked before any pointer is read
    ptrs = (ctypes.c_uint64 * 2)(4096, 8192)
    P = ctypes.cast(ptrs, ctypes.c_void_p)
    for over, needle in ((dict(a0=P, a2=17, a3=0), "world must be 1..16"), (dict(a0=P, a2=2, a3=2), "rank out of range"),
                         (dict(a0=P, a2=2, a3=0, a5=-4), "negative extent"), (dict(a0=P, a2=2, a3=0, a5=6), "multiples of 4")):
        rc, msg = _call(lib, sigs, "gs_peer_allreduce", **over)
        assert rc == -1 and needle in msg, (over, rc, msg)
    assert lib.gs_densify_workspace_bytes(1000) > 0 and lib.gs_loss_workspace_bytes() > 0
