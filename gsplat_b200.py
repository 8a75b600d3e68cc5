"""Importable alias of the `mini-3d-gaussian-splatting_b200` package (whose directory name, fixed
by the build contract, contains hyphens)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("mini-3d-gaussian-splatting_b200")
sys.modules[__name__] = _pkg
