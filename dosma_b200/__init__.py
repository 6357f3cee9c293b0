"""dosma_b200 -- Blackwell-native per-voxel curve fitting, a drop-in for the hot path of ad12/DOSMA.

    from dosma_b200 import CurveFitter, MonoExponentialFit, curve_fit, monoexponential, biexponential

mirror `dosma.core.fitting` (same signatures / return types); the arithmetic runs in hand-written
sm_100a CUDA kernels behind the C-ABI of `include/dfit.h` (libdfit.so, loaded with ctypes).
Importing this package does not create a CUDA context and does not load the library; the first
fit does, and raises if the library or a GPU is missing (no CPU fallback).
"""
from .fitting import CurveFitter, MonoExponentialFit, curve_fit, set_default_compute_dtype  # noqa: F401
from .med_volume import MedicalVolume  # noqa: F401
from .models import biexponential, linear, monoexponential  # noqa: F401

__version__ = "0.1.0"
__all__ = ["CurveFitter", "MonoExponentialFit", "curve_fit", "monoexponential", "biexponential", "linear",
           "MedicalVolume", "set_default_compute_dtype", "patch_dosma"]


def patch_dosma():
    """Monkey-patch an importable DOSMA so its scan pipelines use this engine."""
    import importlib

    from . import fitting as F

    targets = ["dosma.core.fitting", "dosma.core", "dosma", "dosma.scan_sequences.mri.cube_quant",
               "dosma.scan_sequences.mri.mapss", "dosma.scan_sequences.mri.cones"]
    patched = []
    for name in targets:
        try:
            mod = importlib.import_module(name)
        except Exception:
            continue
        for attr in ("CurveFitter", "MonoExponentialFit", "curve_fit"):
            if hasattr(mod, attr):
                setattr(mod, attr, getattr(F, attr))
                patched.append(f"{name}.{attr}")
    return patched
