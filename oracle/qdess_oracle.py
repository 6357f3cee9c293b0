"""CPU ORACLE for the qDESS analytic T2 map -- test infrastructure, NOT product code.

numpy restatement of dosma/scan_sequences/mri/qdess.py:193-255 (the arithmetic of
`QDess.generate_t2_map`).  Parity status: PINNED -- tests/test_qdess_oracle.py checks it, bit for bit, against
tests/golden/qdess_*.npz, the outputs of the real reference method (tests/golden/make_golden_next.py loads qdess.py
verbatim through tests/golden/ref_loader.py::load_reference_qdess), and in the build container against the live
reference on fresh seeds.
"""
import math

import numpy as np


def constants(tr, te, tg, gl_area, alpha, t1, diffusivity=1.25e-9):
    """qdess.py:193-223."""
    TR, TE, Tg, T1 = tr * 1e-3, te * 1e-3, tg * 1e-6, t1 * 1e-3  # :196-199
    alpha = math.radians(alpha)  # :203
    Gl = gl_area / (Tg * 1e6) * 100  # :211
    gamma = 4258 * 2 * math.pi  # :212
    dkL = gamma * Gl * Tg  # :213
    e = np.exp(-TR / T1 - TR * np.power(dkL, 2) * diffusivity)
    k = np.power(np.sin(alpha / 2), 2) * (1 + e) / (1 - np.cos(alpha) * e)  # :216-220
    c1 = (TR - Tg / 3) * np.power(dkL, 2) * diffusivity  # :222
    return float(k), float(c1), TR, TE


def t2_map(echo_1, echo_2, tr, te, tg, gl_area, alpha, t1, diffusivity=1.25e-9, nan_bounds=(0, 100), nan_to_num=0.0,
           decimals=1, suppress_fat=False, suppress_fluid=False, beta=1.2):
    """qdess.py:225-255."""
    k, c1, TR, TE = constants(tr, te, tg, gl_area, alpha, t1, diffusivity)
    with np.errstate(all="ignore"):
        mask = np.ones(echo_1.shape)
        ratio = mask * echo_2 / echo_1  # :228
        ratio = np.nan_to_num(ratio)  # :229
        t2map = -2000 * (TR - TE) / (np.log(abs(ratio) / k) + c1)  # :232
        t2map = np.nan_to_num(t2map)  # :234
        if nan_bounds is not None:
            lower, upper = nan_bounds
            t2map[(t2map < lower) | (t2map > upper)] = np.nan  # :237-239
        if nan_to_num is not None:
            t2map = np.nan_to_num(t2map) if isinstance(nan_to_num, bool) else np.nan_to_num(t2map, nan=nan_to_num)
        if decimals is not None:
            t2map = np.around(t2map, decimals)  # :247-248
        if suppress_fat:
            t2map = t2map * (echo_1 > 0.15 * np.max(echo_1))  # :250-251
        if suppress_fluid:
            vol_null_fluid = echo_1 - beta * echo_2
            t2map = t2map * (vol_null_fluid > 0.1 * np.max(vol_null_fluid))  # :253-255
    return t2map
