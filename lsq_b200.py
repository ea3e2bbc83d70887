"""Importable alias of the package directory `local-search-quantization_b200` (its name has a '-')."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("local-search-quantization_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
api = importlib.import_module("local-search-quantization_b200.api")
device = importlib.import_module("local-search-quantization_b200.device")
build = importlib.import_module("local-search-quantization_b200.build")
parallel = importlib.import_module("local-search-quantization_b200.parallel")
io = importlib.import_module("local-search-quantization_b200.io")
__all__ = list(_pkg.__all__) + ["api", "device", "build", "parallel", "io"]
