"""Closed registry of model functions the CUDA engine implements, and how user callables map onto it.

The reference API takes an arbitrary Python callable `func(x, *params)` (dosma/core/fitting.py:304,
755) and evaluates it ~54 times per voxel through SciPy.  A GPU engine with no CPU fallback can
only run models it has kernels for, so callables are *recognised*:

  1. by identity (the package's own functions);
  2. by code: a function whose bytecode, constants and global names equal a built-in's and whose `np` is numpy
     computes the same expression -- this covers re-definitions such as the reference's own
     `dosma.core.fitting.monoexponential`, the README's (README.md:113-114) and the tests' `_linear`
     (tests/core/test_fitting.py:52-53);
  3. models only: by an EXACT numeric fingerprint on an adversarial probe (six decades of |x|, both signs of
     every parameter, arguments that overflow and underflow `exp`), so that a look-alike such as a clipped
     exponent is not mistaken for the built-in.

Anything else raises NotImplementedError (models) or is applied on the host after the kernel (post-processing
ufuncs, which have a host path: `fitting._process_params_host`).
"""
import inspect

import numpy as np

from . import _cabi

__all__ = ["monoexponential", "biexponential", "linear", "resolve_model", "resolve_ufunc", "param_names", "inv_abs"]


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

def _probe():
    """Adversarial probe: x over six decades, parameters of both signs and magnitudes from 1e-3 to 1e3, so that
    b * x spans exp's underflow / overflow range (|b x| up to ~1e6) as well as its ordinary one."""
    rng = np.random.default_rng(20240607)
    x = np.concatenate([[0.0, 1.0, -1.0], 10.0 ** rng.uniform(-3, 3, 29) * rng.choice([-1.0, 1.0], 29)])
    params = 10.0 ** rng.uniform(-3, 3, (24, 4)) * rng.choice([-1.0, 1.0], (24, 4))
    params[:4] = [[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0], [2.5, -0.04, 0.5, -0.3], [-3.0, 0.7, 2.0, -1e-3]]
    return x, params


_PROBE_X, _PROBE_P = _probe()


def param_names(func):
    """Parameter names by signature introspection, as the reference does (fitting.py:818-821)."""
    names = list(inspect.signature(func).parameters)
    return names[2:] if "self" in names else names[1:]


def _same_code(func, builtin):
    """True when `func` is a plain Python function computing `builtin`'s expression: identical bytecode,
    constants and global names, no closure, and the global `np` (if used) is numpy itself."""
    fc, bc = getattr(func, "__code__", None), builtin.__code__
    if fc is None or getattr(func, "__closure__", None):
        return False
    if (fc.co_code, fc.co_consts, fc.co_names, fc.co_argcount) != (bc.co_code, bc.co_consts, bc.co_names, bc.co_argcount):
        return False
    if getattr(func, "__defaults__", None) or getattr(func, "__kwdefaults__", None):
        return False
    return "np" not in fc.co_names or getattr(func, "__globals__", {}).get("np") is np


def resolve_model(func):
    """Return (model_id, nparams) for a user callable or raise NotImplementedError."""
    for mid, f, n in _BUILTINS:
        if func is f:
            return mid, n
    for mid, f, n in _BUILTINS:
        if _same_code(func, f):
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
                    np.array_equal(np.asarray(func(_PROBE_X, *p[:n]), dtype=np.float64), f(_PROBE_X, *p[:n]), equal_nan=True)
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


# Post-processing ufuncs the fused epilogue implements (fitting.py:123-128).  `inv_abs` is the function object
# MonoExponentialFit hands to its CurveFitter (fitting.py:725).
def identity(v):
    return v


def inv_abs(v):
    return 1 / np.abs(v)


def neg_inv(v):
    return -1 / v


def absolute(v):
    return np.abs(v)


def inv(v):
    return 1 / v


_UFUNCS = (
    (_cabi.UFUNC_NONE, identity),
    (_cabi.UFUNC_INV_ABS, inv_abs),
    (_cabi.UFUNC_NEG_INV, neg_inv),
    (_cabi.UFUNC_ABS, absolute),
    (_cabi.UFUNC_INV, inv),
)


def resolve_ufunc(fn):
    """Map a post-processing callable (fitting.py:123-128) onto an epilogue id -- by identity or by code (see the
    module docstring; e.g. a user's own `lambda x: 1 / np.abs(x)`), `np.abs` / `np.absolute` by identity -- or
    return None: the callable is then applied on the host after the kernel (`fitting._process_params_host`).
    Nothing is guessed from sample evaluations: a callable that merely agrees with a built-in on some inputs
    (a clamped reciprocal ...) keeps its own semantics."""
    if fn is None:
        return _cabi.UFUNC_NONE
    if fn is np.abs or fn is np.absolute or fn is np.fabs or fn is abs:
        return _cabi.UFUNC_ABS
    for uid, ref in _UFUNCS:
        if fn is ref or _same_code(fn, ref):
            return uid
    return None
