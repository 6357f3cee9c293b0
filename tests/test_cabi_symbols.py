"""CPU checks of the C-ABI: the library loads, exports every symbol include/dfit.h declares, and
refuses to compute without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dosma_b200 import _cabi, build

    build.build()
    return _cabi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dfit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfit_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from dosma_b200 import _cabi

    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dfit.h but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == syms


def test_opts_struct_matches_header(lib):
    from dosma_b200 import _cabi

    o = _cabi.default_opts(_cabi.MODEL_MONOEXP)
    assert o.struct_size == ctypes.sizeof(_cabi.DfitOpts)
    assert (o.maxfev, o.ftol, o.r2_eps) == (100, 1e-5, 1e-8)  # fitting.py:761-763
    assert list(o.p0) == [1.0] * 4 and o.y_lo == float("-inf") and o.y_hi == float("inf")
    assert lib.dfit_model_nparams(_cabi.MODEL_MONOEXP) == 2
    assert lib.dfit_model_nparams(_cabi.MODEL_BIEXP) == 4
    assert lib.dfit_model_nparams(_cabi.MODEL_LINEAR) == 1
    assert lib.dfit_version() == 300 and o.out_param == -1
    bad = _cabi.DfitOpts()
    assert lib.dfit_default_opts(ctypes.byref(bad), 99) != 0


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute path must fail loudly."""
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dosma_b200 import _cabi
    import dosma_b200 as D

    assert lib.dfit_device_count() == 0
    h = ctypes.c_void_p()
    assert lib.dfit_create(0, ctypes.byref(h)) == -2  # DFIT_ERR_NO_DEVICE
    with pytest.raises(_cabi.DfitError):
        D.curve_fit(D.monoexponential, [1.0, 2, 3, 4], np.ones((4, 8)))


def test_product_does_not_touch_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/ (or the hostsim)."""
    pkg = os.path.join(ROOT, "dosma_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M), f
                assert "liboracle" not in src and "hostsim" not in src.replace("tests/hostsim", ""), f


def test_ctypes_structs_have_the_header_layout(tmp_path):
    """The ctypes mirrors in dosma_b200/_cabi.py against include/dfit.h as gcc lays it out: sizes and the offsets of the
    fields most recently added (an ABI guard for the binding a DOSMA maintainer would copy from INTEGRATION.md)."""
    import subprocess
    import textwrap

    from dosma_b200 import _cabi

    src = tmp_path / "layout.c"
    src.write_text(textwrap.dedent('''
        #include <stddef.h>
        #include <stdio.h>
        #include "dfit.h"
        int main(void) {
          printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dfit_opts), sizeof(dfit_stats), sizeof(dfit_gather_desc),
                 sizeof(dfit_qdess_opts), offsetof(dfit_opts, out_param), offsetof(dfit_opts, decimals),
                 offsetof(dfit_stats, n_deferred), offsetof(dfit_gather_desc, y_voxel0));
          return 0;
        }'''))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_cabi.DfitOpts), ctypes.sizeof(_cabi.DfitStats), ctypes.sizeof(_cabi.DfitGatherDesc),
            ctypes.sizeof(_cabi.DfitQdessOpts), _cabi.DfitOpts.out_param.offset, _cabi.DfitOpts.decimals.offset,
            _cabi.DfitStats.n_deferred.offset, _cabi.DfitGatherDesc.y_voxel0.offset]
    assert got == want, (got, want)
