"""Importable alias of the hyphen-named package ``medical-image-denoising-using-diffusion_b200``."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _il.import_module("medical-image-denoising-using-diffusion_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
models = _il.import_module("medical-image-denoising-using-diffusion_b200.models")
_lib = _il.import_module("medical-image-denoising-using-diffusion_b200._lib")
__all__ = list(_pkg.__all__) + ["models"]
