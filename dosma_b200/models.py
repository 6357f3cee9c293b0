"""Closed registry of model functions the CUDA engine implements, and how user callables map onto it.

The reference API takes an arbitrary Python callable `func(x, *params)` (dosma/core/fitting.py:304,
755) and evaluates it ~54 times per voxel through SciPy.  A GPU engine with no CPU fallback can
only run models it has kernels for, so callables are *recognised*: by identity, then by a numeric
fingerprint (the callable is evaluated on a small probe and compared with each built-in that has
the same number of parameters).  This makes user re-definitions such as the README's own
`monoexponential` (README.md:113-114) or the tests' `_linear` (tests/core/test_fitting.py:52-53)
work unchanged.  Anything else raises NotImplementedError with the list of supported models.
"""
import inspect

import numpy as np

from . import _cabi

__all__ = ["monoexponential", "biexponential", "linear", "resolve_model", "resolve_ufunc", "param_names"]


def monoexponential(x, a, b):
    """:math:`f(x) = a e^{b x}` -- same definition as dosma/core/fitting.py:1016-1018."""
    return a * np.exp(b * x)


def biexponential(x, a1, b1, a2, b2):
    """:math:`f(x) = a_1 e^{b_1 x} + a_2 e^{b_2 x}` -- dosma/core/fitting.py:1021-1023."""
    return a1 * np.exp(b1 * x) + a2 * np.exp(b2 * x)


def linear(x, a):
    """:math:`f(x) = a x`."""
    return a * x


_BUILTINS = (
    (_cabi.MODEL_MONOEXP, monoexponential, 2),
    (_cabi.MODEL_BIEXP, biexponential, 4),
    (_cabi.MODEL_LINEAR, linear, 1),
)

_PROBE_X = np.array([0.25, 0.9, 1.7, 3.1])
_PROBE_P = np.array([[0.7, -0.4, 1.3, -0.9], [-1.2, 0.3, 0.6, 0.15], [2.5, -1.1, -0.8, -0.05]])


def param_names(func):
    """Parameter names by signature introspection, as the reference does (fitting.py:818-821)."""
    names = list(inspect.signature(func).parameters)
    return names[2:] if "self" in names else names[1:]


def resolve_model(func):
    """Return (model_id, nparams) for a user callable or raise NotImplementedError."""
    for mid, f, n in _BUILTINS:
        if func is f:
            return mid, n
    try:
        nparams = len(param_names(func))
    except (TypeError, ValueError) as e:
        raise ValueError(f"Unable to determine number of fit parameters of {func}") from e
    for mid, f, n in _BUILTINS:
        if n != nparams:
            continue
        try:
            with np.errstate(all="ignore"):
                same = all(
                    np.allclose(np.asarray(func(_PROBE_X, *p[:n]), dtype=np.float64), f(_PROBE_X, *p[:n]),
                                rtol=1e-12, atol=0)
                    for p in _PROBE_P
                )
        except Exception:
            same = False
        if same:
            return mid, n
    name = getattr(func, "__name__", type(func).__name__)
    raise NotImplementedError(
        f"Model function '{name}' ({nparams} parameters) is not one the CUDA engine implements. "
        "Supported: monoexponential a*exp(b*x), biexponential a1*exp(b1*x)+a2*exp(b2*x), linear a*x. "
        "dosma_b200 has no CPU fallback by design."
    )


_UFUNC_PROBE = np.array([-2.5, -0.5, 0.25, 3.0])
_UFUNCS = (
    (_cabi.UFUNC_NONE, lambda v: v),
    (_cabi.UFUNC_INV_ABS, lambda v: 1 / np.abs(v)),
    (_cabi.UFUNC_NEG_INV, lambda v: -1 / v),
    (_cabi.UFUNC_ABS, lambda v: np.abs(v)),
    (_cabi.UFUNC_INV, lambda v: 1 / v),
)


def resolve_ufunc(fn):
    """Map a post-processing callable (fitting.py:123-128) onto an epilogue id, or None if it has
    to be applied on the host after the kernel."""
    if fn is None:
        return _cabi.UFUNC_NONE
    try:
        out = np.asarray(fn(_UFUNC_PROBE.copy()), dtype=np.float64)
    except Exception:
        return None
    if out.shape != _UFUNC_PROBE.shape:
        return None
    for uid, ref in _UFUNCS:
        if np.allclose(out, ref(_UFUNC_PROBE), rtol=1e-14, atol=0):
            return uid
    return None
