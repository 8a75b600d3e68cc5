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
