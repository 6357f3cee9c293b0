"""Minimal spatial-volume container used at the drop-in boundary.

The fitters accept and return *the caller's* volume type by duck typing: with a real
`dosma.MedicalVolume` they call the same members the reference calls on it
(`dosma/core/fitting.py:157-235`: `.volume`, `.affine`, `.orientation`, `.headers()`,
`.reformat()`, `.reformat_as()`, `.is_same_dimensions()`, `._partial_clone()`) and hand back
objects made by `y[0]._partial_clone(...)`, so DOSMA pipelines keep their own class.

This module provides an independent, much smaller container with that member surface for use
when DOSMA itself is not importable (e.g. on the GPU box: nibabel/pydicom are not installed).
It is NOT a rebuild of `dosma/core/med_volume.py` (I/O, numpy protocol, SimpleITK/torch interop
are out of scope) -- only what the curve-fit path touches.

Conventions follow DOSMA's documented ones (`dosma/core/orientation.py:1-58`): RAS+ affine,
orientation strings such as ("SI", "AP", "LR") name the direction in which each array axis runs.
"""
from copy import deepcopy

import numpy as np

__all__ = ["MedicalVolume", "is_volume"]

_AXIS_CODES = (("RL", "LR"), ("AP", "PA"), ("SI", "IS"))  # world x, y, z: (decreasing, increasing)
_AXIS_OF = {"LR": 0, "RL": 0, "PA": 1, "AP": 1, "IS": 2, "SI": 2}
_ORIGIN_DECIMALS = 4


def is_volume(obj):
    """Duck-type test used instead of isinstance so that real DOSMA volumes are accepted too."""
    return all(hasattr(obj, a) for a in ("volume", "affine", "orientation", "reformat", "_partial_clone"))


def _orientation_from_affine(affine):
    """Closest world axis (and sign) for each of the three spatial array axes, assigned greedily by
    decreasing direction cosine so that oblique scans still get three distinct axes."""
    rot = np.asarray(affine, dtype=np.float64)[:3, :3]
    norms = np.sqrt((rot * rot).sum(axis=0))
    norms[norms == 0] = 1.0
    cos = rot / norms
    remaining = np.abs(cos)
    out = [None, None, None]
    for _ in range(3):
        w, i = np.unravel_index(np.argmax(remaining), remaining.shape)
        out[i] = _AXIS_CODES[w][1 if cos[w, i] > 0 else 0]
        remaining[w, :] = -1
        remaining[:, i] = -1
    return tuple(out)


class MedicalVolume:
    """ndarray + 4x4 RAS+ affine (+ optional broadcastable header array)."""

    def __init__(self, volume, affine, headers=None):
        self._volume = np.asarray(volume)
        self._affine = np.array(affine, dtype=np.float64)
        if self._affine.shape != (4, 4):
            raise ValueError("`affine` must be 4x4")
        self._headers = self._format_headers(headers) if headers is not None else None

    # ----------------------------------------------------------------- basic properties
    def _format_headers(self, headers):
        headers = np.asarray(headers)
        if headers.ndim > self._volume.ndim:
            raise ValueError("`headers` has too many dimensions")
        headers = headers.reshape((1,) * (self._volume.ndim - headers.ndim) + headers.shape)
        for d in range(self._volume.ndim):
            if headers.shape[d] not in (1, self._volume.shape[d]):
                raise ValueError("`headers` must broadcast against the volume")
        return headers

    @property
    def volume(self):
        return self._volume

    @volume.setter
    def volume(self, value):
        value = np.asarray(value)
        if value.ndim != self._volume.ndim:
            raise ValueError("New volume must have the same number of dimensions")
        self._volume = value

    @property
    def A(self):
        return self._volume

    @property
    def affine(self):
        return self._affine

    @property
    def shape(self):
        return tuple(self._volume.shape)

    @property
    def ndim(self):
        return self._volume.ndim

    @property
    def dtype(self):
        return self._volume.dtype

    @property
    def device(self):
        return "cpu"

    @property
    def orientation(self):
        return _orientation_from_affine(self._affine)

    @property
    def pixel_spacing(self):
        return tuple(np.sqrt((self._affine[:3, :3] ** 2).sum(axis=0)))

    @property
    def scanner_origin(self):
        return tuple(self._affine[:3, 3])

    def headers(self, flatten=False):
        if flatten and self._headers is not None:
            return self._headers.flatten()
        return self._headers

    # ----------------------------------------------------------------- cloning / casting
    def _partial_clone(self, **kwargs):
        """Constructor arguments default to copies of this volume's (same contract as the reference's
        private helper that `_Fitter.fit` relies on, fitting.py:232-235): `headers=True` deep-copies,
        `headers=None` drops, `volume=False` shares the array."""
        if kwargs.get("volume", None) is False:
            kwargs["volume"] = self._volume
        for k in ("volume", "affine"):
            if k not in kwargs or kwargs[k] is True:
                kwargs[k] = getattr(self, "_" + k).copy()
        if "headers" not in kwargs:
            kwargs["headers"] = self._headers
        elif kwargs["headers"] is True:
            kwargs["headers"] = deepcopy(self._headers)
        return type(self)(**kwargs)

    def clone(self, headers=True):
        return self._partial_clone(headers=headers)

    def astype(self, dtype, **kwargs):
        return self._partial_clone(volume=self._volume.astype(dtype, **kwargs))

    # ----------------------------------------------------------------- orientation
    def reformat(self, new_orientation, inplace=False):
        """Transpose/flip the spatial axes into `new_orientation`; the affine follows."""
        new_orientation = tuple(new_orientation)
        cur = self.orientation
        if new_orientation == cur:
            return self if inplace else self._partial_clone(volume=self._volume)
        if sorted(_AXIS_OF[o] for o in new_orientation) != [0, 1, 2]:
            raise ValueError(f"Invalid orientation {new_orientation}")
        cur_axes = [_AXIS_OF[o] for o in cur]
        perm = tuple(cur_axes.index(_AXIS_OF[o]) for o in new_orientation)
        full_perm = perm + tuple(range(3, self._volume.ndim))
        vol = np.transpose(self._volume, full_perm)
        hdr = np.transpose(self._headers, full_perm) if self._headers is not None else None
        aff = self._affine.copy()
        aff[:, :3] = self._affine[:, list(perm)]
        flips = [i for i in range(3) if cur[perm[i]] != new_orientation[i]]
        if flips:
            vol = np.flip(vol, axis=tuple(flips))
            if hdr is not None:
                hdr = np.flip(hdr, axis=tuple(flips))
            origin = aff[:3, 3].copy()
            for i in flips:
                origin = origin + aff[:3, i] * (vol.shape[i] - 1)
                aff[:3, i] = -aff[:3, i]
            aff[:3, 3] = np.round(origin, _ORIGIN_DECIMALS)
        aff[aff == 0] = 0  # no negative zeros
        if inplace:
            self._volume, self._affine, self._headers = vol, aff, hdr
            return self
        return self._partial_clone(volume=vol, affine=aff, headers=hdr)

    def reformat_as(self, other, inplace=False):
        return self.reformat(other.orientation, inplace=inplace)

    def is_same_dimensions(self, mv, precision=None, err=False):
        if not is_volume(mv):
            raise TypeError("`mv` must be a MedicalVolume.")
        if precision is not None:
            tol = 10 ** (-precision)
            close = np.allclose(mv.affine[:3, :3], self.affine[:3, :3], atol=tol) and np.allclose(
                mv.affine[:3, 3], self.affine[:3, 3], rtol=tol
            )
        else:
            close = bool((np.asarray(mv.affine) == self.affine).all())
        same_o = tuple(mv.orientation) == self.orientation
        same_s = tuple(mv.volume.shape) == self.shape
        out = close and same_o and same_s
        if err and not out:
            raise ValueError(
                "Volumes differ: affine close=%s, orientation equal=%s, shape equal=%s" % (close, same_o, same_s)
            )
        return out

    def is_identical(self, mv):
        return self.is_same_dimensions(mv) and bool((np.asarray(mv.volume) == self._volume).all())

    # ----------------------------------------------------------------- slicing
    def __getitem__(self, key):
        """Numpy-style slicing; spatial axes may be sliced (not dropped) and the affine follows.
        Headers are indexed with a *tuple* (the reference's list indexing breaks on numpy >= 1.23,
        SURVEY.md Appendix D)."""
        if is_volume(key):
            key = key.reformat_as(self).volume
        if not isinstance(key, tuple):
            key = (key,)
        nd = self._volume.ndim
        if any(k is Ellipsis for k in key):
            i = [j for j, k in enumerate(key) if k is Ellipsis][0]
            key = key[:i] + (slice(None),) * (nd - (len(key) - 1)) + key[i + 1:]
        key = key + (slice(None),) * (nd - len(key))
        if len(key) > nd:
            raise IndexError("too many indices for volume")
        for k in key[:3]:
            if isinstance(k, (int, np.integer)) or k is None:
                raise IndexError("Cannot drop or add spatial dimensions")
        vol = self._volume[key]
        if any(d == 0 for d in vol.shape):
            raise IndexError("Empty slice requested")
        hdr = self._headers
        if hdr is not None:
            hkey = tuple(
                (0 if isinstance(k, (int, np.integer)) else slice(None)) if hdr.shape[i] == 1 else k
                for i, k in enumerate(key)
            )
            hdr = hdr[hkey]
        aff = self._affine.copy()
        for i, k in enumerate(key[:3]):
            if isinstance(k, slice):
                start, _, step = k.indices(self._volume.shape[i])
                aff[:3, 3] = aff[:3, 3] + aff[:3, i] * start
                aff[:3, i] = aff[:3, i] * step
        return self._partial_clone(volume=vol, affine=aff, headers=hdr)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._volume, dtype=dtype)

    def __gt__(self, other):
        return self._partial_clone(volume=self._volume > other, headers=None)

    def __bool__(self):
        return True

    def __repr__(self):
        return (f"MedicalVolume(shape={self.shape}, dtype={self.dtype}, orientation={self.orientation}, "
                f"spacing={tuple(round(float(s), 4) for s in self.pixel_spacing)})")
