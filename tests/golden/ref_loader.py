"""Load the *reference* ``dosma.core.fitting`` module verbatim, in THIS container only.

TEST INFRASTRUCTURE -- not product code.  Used by ``make_golden.py`` (and by the
``needs_reference`` tests) to obtain real reference outputs.  ``/root/reference`` does not
exist on the GPU box, so nothing in the ``-m gpu`` tests / ``smoke()`` / ``bench.py`` imports
this file.

Why this dance:
  * ``import dosma`` fails here (termcolor, nibabel, pydicom, matplotlib, ... are not installed
    and there is no network), but ``dosma/core/fitting.py`` itself only needs numpy, scipy, tqdm
    and three sibling modules.
  * ``dosma/defaults.py:30-31`` *writes* ``resources/preferences.yml`` next to itself on import,
    so the package is first copied to a scratch directory; ``/root/reference`` is never written.
  * The heavy package ``__init__`` files are skipped by pre-registering namespace modules, and
    the missing third-party imports are replaced by tiny stand-ins that only provide the
    attributes the fitting path touches (``nib.aff2axcodes`` and nibabel's spatial slicer).
"""
import importlib
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DOSMA_REFERENCE_ROOT", "/root/reference")

_CACHE = {}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dosma", "core"))


def _aff2axcodes(aff, labels=(("L", "R"), ("P", "A"), ("I", "S"))):
    """Stand-in for ``nibabel.aff2axcodes`` (closest world axis per voxel axis)."""
    rzs = np.asarray(aff, dtype=float)[:3, :3]
    zooms = np.sqrt((rzs * rzs).sum(axis=0))
    zooms[zooms == 0] = 1.0
    rs = rzs / zooms
    # Closest pure rotation (polar decomposition) so oblique scans resolve like nibabel does.
    u, s, vt = np.linalg.svd(rs)
    tol = s.max() * 3 * np.finfo(s.dtype).eps
    keep = s > tol
    r = u[:, keep] @ vt[keep, :]
    codes = [None, None, None]
    r = np.array(r)
    for in_ax in range(3):
        col = r[:, in_ax]
        if np.allclose(col, 0):
            continue
        out_ax = int(np.argmax(np.abs(col)))
        codes[in_ax] = labels[out_ax][1] if col[out_ax] > 0 else labels[out_ax][0]
        r[out_ax, :] = 0
    return tuple(codes)


class _SpatialFirstSlicer:
    """Stand-in for ``nibabel.spatialimages.SpatialFirstSlicer`` (only what ``__getitem__`` uses)."""

    def __init__(self, img):
        self.img = img

    def check_slicing(self, slicer, return_spatial=False):
        if not isinstance(slicer, tuple):
            slicer = (slicer,)
        ndim = self.img.volume.ndim if hasattr(self.img, "volume") else self.img.ndim
        # expand Ellipsis
        if any(s is Ellipsis for s in slicer):
            i = [k for k, s in enumerate(slicer) if s is Ellipsis][0]
            n_real = len([s for s in slicer if s is not Ellipsis and s is not None])
            slicer = slicer[:i] + (slice(None),) * (ndim - n_real) + slicer[i + 1 :]
        slicer = slicer + (slice(None),) * (ndim - len([s for s in slicer if s is not None]))
        spatial = slicer[:3]
        for s in spatial:
            if s is None:
                raise ValueError("Cannot add axes in spatial dimensions")
            if isinstance(s, (int, np.integer)):
                raise ValueError("Cannot drop spatial dimensions")
        return spatial if return_spatial else slicer

    def slice_affine(self, slicer):
        slicer = self.check_slicing(slicer, return_spatial=True)
        shape = self.img.shape[:3]
        transform = np.eye(4)
        for i, s in enumerate(slicer):
            if isinstance(s, slice):
                start, _, step = s.indices(shape[i])
                transform[i, i] = step
                transform[i, 3] = start
            else:  # boolean / fancy index: keep the affine (reference behaviour is undefined)
                pass
        return np.asarray(self.img.affine).dot(transform)


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def load_reference_fitting():
    """Return ``(fitting_module, MedicalVolume_class)`` of the real reference code."""
    if "F" in _CACHE:
        return _CACHE["F"], _CACHE["MV"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")

    scratch = tempfile.mkdtemp(prefix="dosma_ref_")
    pkg = os.path.join(scratch, "dosma")
    shutil.copytree(os.path.join(REFERENCE_ROOT, "dosma"), pkg)
    for root, dirs, files in os.walk(pkg):  # reference tree is read-only; the copy must not be
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)

    if not hasattr(np, "round_"):
        np.round_ = np.round  # removed in numpy 2; dosma/core/numpy_routines.py:164
    if not hasattr(np, "bool"):
        np.bool = bool

    def _nested_lookup(key, d):
        out = []

        def rec(x):
            if isinstance(x, dict):
                for k, v in x.items():
                    if k == key:
                        out.append(v)
                    rec(v)
            elif isinstance(x, (list, tuple)):
                for v in x:
                    rec(v)

        rec(d)
        return out

    stubs = {
        "matplotlib": _stub("matplotlib", rcParams={}),
        "nested_lookup": _stub(
            "nested_lookup",
            nested_lookup=_nested_lookup,
            get_occurrence_of_key=lambda d, key: len(_nested_lookup(key, d)),
            nested_update=lambda *a, **k: None,
        ),
        "nibabel": _stub("nibabel", aff2axcodes=_aff2axcodes),
        "nibabel.spatialimages": _stub("nibabel.spatialimages", SpatialFirstSlicer=_SpatialFirstSlicer),
        "nibabel.orientations": _stub("nibabel.orientations"),
        "pydicom": _stub("pydicom"),
        "termcolor": _stub("termcolor", colored=lambda s, *a, **k: s),
    }
    stubs["nibabel"].spatialimages = stubs["nibabel.spatialimages"]
    stubs["nibabel"].orientations = stubs["nibabel.orientations"]
    for name, mod in stubs.items():
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = mod

    for name, sub in (
        ("dosma", ""),
        ("dosma.core", "core"),
        ("dosma.core.io", os.path.join("core", "io")),
        ("dosma.utils", "utils"),
    ):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(pkg, sub) if sub else pkg]
        sys.modules[name] = mod

    sys.path.insert(0, scratch)
    try:
        F = importlib.import_module("dosma.core.fitting")
        importlib.import_module("dosma.core.numpy_routines")
        MV = importlib.import_module("dosma.core.med_volume").MedicalVolume
    finally:
        sys.path.remove(scratch)
    _CACHE["F"], _CACHE["MV"], _CACHE["scratch"] = F, MV, scratch
    return F, MV


def load_reference_quant_vals():
    """The real ``dosma/core/quant_vals.py`` (``QuantitativeValue.to_metrics``, :145-229).  Its two I/O imports
    (``dosma.core.io.format_io[_utils]``, which pull in natsort / nibabel / pydicom / h5py for load / save) are
    replaced by empty stand-ins; the module itself runs verbatim on the real ``MedicalVolume``."""
    if "Q" in _CACHE:
        return _CACHE["Q"]
    load_reference_fitting()
    import enum

    class ImageDataFormat(enum.Enum):
        nifti = 1
        dicom = 2

    for name, attrs in (("dosma.core.io.format_io", {"ImageDataFormat": ImageDataFormat}),
                        ("dosma.core.io.format_io_utils", {})):
        if name not in sys.modules:
            sys.modules[name] = _stub(name, **attrs)
    sys.modules["dosma.core.io"].format_io_utils = sys.modules["dosma.core.io.format_io_utils"]
    sys.path.insert(0, _CACHE["scratch"])
    try:
        Q = importlib.import_module("dosma.core.quant_vals")
    finally:
        sys.path.remove(_CACHE["scratch"])
    _CACHE["Q"] = Q
    return Q


def load_reference_qdess():
    """The real ``dosma/scan_sequences/mri/qdess.py`` (``QDess.generate_t2_map``, :100-258).  What the module imports
    besides numpy -- pydicom, the Keras segmentation models, the scan base class, the tissue classes, the CLI
    helpers -- is replaced by minimal stand-ins that provide exactly the members ``generate_t2_map`` touches
    (``ScanSequence.__init__ / get_metadata``, ``pydicom.Dataset``, ``Tag``); the method itself runs verbatim on
    the real ``MedicalVolume`` and wraps its result in the real ``quant_vals.T2``."""
    if "QD" in _CACHE:
        return _CACHE["QD"]
    load_reference_quant_vals()

    class Dataset(dict):
        pass

    class ScanSequence:
        def __init__(self, volumes):
            self.volumes = volumes
            self.ref_dicom = None
            self._metadata = {}

        def get_metadata(self, key, default=None):  # scans.py:88-116
            metadata = self._metadata.get(key, None)
            if metadata is None and self.ref_dicom is not None:
                metadata = self.ref_dicom[key].value if key in self.ref_dicom else None
            if metadata is None and default is False:
                raise KeyError(f"Metadata '{key}' not found")
            return default if metadata is None else metadata

    pyd = sys.modules.get("pydicom")
    if not hasattr(pyd, "Dataset"):
        pyd.Dataset = Dataset
    if "pydicom.tag" not in sys.modules:
        sys.modules["pydicom.tag"] = _stub("pydicom.tag", Tag=lambda v: v)
        pyd.tag = sys.modules["pydicom.tag"]
    for name, attrs in (
        ("dosma.models", {}),
        ("dosma.models.seg_model", {"SegModel": type("SegModel", (), {})}),
        ("dosma.scan_sequences", {}),
        ("dosma.scan_sequences.scans", {"ScanSequence": ScanSequence}),
        ("dosma.scan_sequences.mri", {}),
        ("dosma.tissues", {}),
        ("dosma.tissues.tissue", {"Tissue": type("Tissue", (), {})}),
        ("dosma.utils.cmd_line_utils", {"ActionWrapper": type("ActionWrapper", (), {"__init__": lambda self, *a, **k: None})}),
    ):
        if name not in sys.modules:
            mod = _stub(name, **attrs)
            sys.modules[name] = mod
    scratch = _CACHE["scratch"]
    sys.modules["dosma.scan_sequences.mri"].__path__ = [os.path.join(scratch, "dosma", "scan_sequences", "mri")]
    sys.path.insert(0, scratch)
    try:
        QD = importlib.import_module("dosma.scan_sequences.mri.qdess")
    finally:
        sys.path.remove(scratch)
    _CACHE["QD"] = QD
    return QD
