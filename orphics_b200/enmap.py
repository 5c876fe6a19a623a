"""Host-side geometry for the flat-sky path: a minimal stand-in for the bits of
pixell.enmap that orphics.maps touches at set-up time (shape/wcs bookkeeping, the
Fourier axes and the patch area).  Everything per-pixel is computed on the device
through the Geometry handle (liborphx ox_geometry_*).

Reference call sites: enmap.geometry maps.py:1490; enmap.area maps.py:1567,1605;
enmap.lmap/modlmap/laxes maps.py:1374,1607,1938-1939.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib, check, ptr

degree = np.pi / 180.0
arcmin = degree / 60.0

# pixell's extent model for separable (CAR) projections; "intermediate" = plain |cdelt|*shape
DEFAULT_EXTENT = "cylindrical"


class FlatWCS:
    """cdelt/crval/crpix (degrees, [x(ra), y(dec)] order, 1-based crpix) of a CAR patch.
    A real astropy/pixell WCS is accepted anywhere a FlatWCS is, by duck typing on
    ``wcs.wcs.cdelt/crval/crpix``."""

    def __init__(self, cdelt, crval, crpix, ctype=("RA---CAR", "DEC--CAR")):
        self.cdelt = np.array(cdelt, dtype=np.float64)
        self.crval = np.array(crval, dtype=np.float64)
        self.crpix = np.array(crpix, dtype=np.float64)
        self.ctype = tuple(ctype)

    @property
    def wcs(self):
        return self

    def key(self):
        return tuple(self.cdelt) + tuple(self.crval) + tuple(self.crpix)

    def __repr__(self):
        return "car:{cdelt:[%.4g,%.4g],crval:[%.4g,%.4g],crpix:[%.2f,%.2f]}" % (
            self.cdelt[0], self.cdelt[1], self.crval[0], self.crval[1], self.crpix[0], self.crpix[1])


def _w(wcs):
    w = getattr(wcs, "wcs", wcs)
    return np.asarray(w.cdelt, dtype=np.float64), np.asarray(w.crval, dtype=np.float64), np.asarray(w.crpix, dtype=np.float64)


class ndmap(np.ndarray):
    """numpy array that carries a wcs (what the reference's methods return, maps.py:1677)."""

    def __new__(cls, arr, wcs=None):
        obj = np.asarray(arr).view(cls)
        obj.wcs = wcs
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.wcs = getattr(obj, "wcs", None)


def enmap(arr, wcs=None):
    return ndmap(arr, wcs)


def samewcs(arr, *maps):
    for m in maps:
        if getattr(m, "wcs", None) is not None:
            return ndmap(arr, m.wcs)
    return arr


def geometry(pos, res, proj="car"):
    """Plate-carree geometry of the box pos=[[dec0,ra0],[dec1,ra1]] (radians) at
    resolution res (radians): crval=[mid ra, 0]; pos[0] is the outer corner of
    pixel (0,0); shape = round(|pixel of pos[1]|)."""
    if proj != "car":
        raise NotImplementedError("only proj='car' is supported")
    box = np.asarray(pos, dtype=np.float64)[:, ::-1] / degree
    step = (np.zeros(2) + np.asarray(res, dtype=np.float64)) / degree
    crval = np.array([box[:, 0].mean(), 0.0])
    cdelt = np.where(box[1] < box[0], -step, step)
    crpix = 0.5 - (box[0] - crval) / cdelt
    far = (box[1] - crval) / cdelt + crpix - 1
    shape = tuple(int(v) for v in np.round(np.abs(far[::-1])))
    return shape, FlatWCS(cdelt, crval, crpix)


def extent(shape, wcs, signed=False, method=None):
    method = method or DEFAULT_EXTENT
    cdelt, crval, crpix = _w(wcs)
    ny, nx = shape[-2:]
    ext = np.array([ny * cdelt[1], nx * cdelt[0]]) * degree
    if method == "cylindrical":
        dec = (crval[1] + (np.array([-0.5, ny - 0.5]) + 1 - crpix[1]) * cdelt[1]) * degree
        ext[1] *= (np.sin(dec[1]) - np.sin(dec[0])) / (dec[1] - dec[0])
    elif method != "intermediate":
        raise ValueError(f"unknown extent method {method!r}")
    return ext if signed else np.abs(ext)


def area(shape, wcs, method=None):
    return float(np.prod(extent(shape, wcs, method=method)))


def pixsize(shape, wcs, method=None):
    return area(shape, wcs, method) / (shape[-2] * shape[-1])


def laxes(shape, wcs, method=None):
    step = extent(shape, wcs, signed=True, method=method) / np.array(shape[-2:], dtype=np.float64)
    ly = np.fft.fftfreq(shape[-2], step[0]) * 2 * np.pi
    lx = np.fft.fftfreq(shape[-1], step[1]) * 2 * np.pi
    return ly, lx


class Geometry:
    """Device-side geometry handle (ox_geometry): the Fourier axes live in HBM and
    per-pixel quantities (modlmap, rotation matrix, l-masks, interpolated spectra)
    are produced by kernels."""

    _cache = {}

    def __init__(self, shape, wcs, method=None):
        _capi.require_device()
        self.shape = tuple(int(s) for s in shape[-2:])
        self.wcs = wcs
        self.method = method or DEFAULT_EXTENT
        self.ly, self.lx = laxes(self.shape, wcs, self.method)
        self.ly = np.ascontiguousarray(self.ly)
        self.lx = np.ascontiguousarray(self.lx)
        self.area = area(self.shape, wcs, self.method)
        h = C.c_void_p()
        check(lib.ox_geometry_create(self.shape[0], self.shape[1], ptr(self.ly), ptr(self.lx), self.area, C.byref(h)))
        self.handle = h
        self._modlmap = None

    @classmethod
    def get(cls, shape, wcs, method=None):
        cdelt, crval, crpix = _w(wcs)
        key = (tuple(int(s) for s in shape[-2:]), tuple(cdelt), tuple(crval), tuple(crpix), method or DEFAULT_EXTENT)
        g = cls._cache.get(key)
        if g is None:
            g = cls._cache[key] = cls(shape, wcs, method)
        return g

    @property
    def npix(self):
        return self.shape[0] * self.shape[1]

    def modlmap(self):
        if self._modlmap is None:
            out = np.empty(self.shape, dtype=np.float64)
            check(lib.ox_geometry_modlmap(self.handle, ptr(out), _capi.OX_HOST))
            self._modlmap = out
        return self._modlmap

    def rotmat(self, iau=False):
        out = np.empty((2, 2) + self.shape, dtype=np.float64)
        check(lib.ox_geometry_rotmat(self.handle, _capi.FLAG_IAU if iau else 0, ptr(out), _capi.OX_HOST))
        return out

    def mask_kspace(self, lxcut=None, lycut=None, lmin=None, lmax=None):
        nan = float("nan")
        f = lambda v: nan if v is None else float(v)
        out = np.empty(self.shape, dtype=np.int32)
        check(lib.ox_geometry_mask_kspace(self.handle, f(lxcut), f(lycut), f(lmin), f(lmax), ptr(out), _capi.OX_HOST))
        return out

    def interp_spec(self, spec):
        spec = np.ascontiguousarray(spec, dtype=np.float64)
        lead = spec.shape[:-1]
        flat = spec.reshape(-1, spec.shape[-1])
        out = np.empty((flat.shape[0],) + self.shape, dtype=np.float64)
        check(lib.ox_geometry_interp_spec(self.handle, ptr(flat), flat.shape[0], flat.shape[1], ptr(out), _capi.OX_HOST))
        return out.reshape(lead + self.shape)

    def __del__(self):
        try:
            lib.ox_geometry_destroy(self.handle)
        except Exception:
            pass


def lmap(shape, wcs, method=None):
    ly, lx = laxes(shape, wcs, method)
    out = np.empty((2,) + tuple(shape[-2:]))
    out[0] = ly[:, None]
    out[1] = lx[None, :]
    return ndmap(out, wcs)


def modlmap(shape, wcs, method=None):
    """|l| on the Fourier grid, computed on the device (bit-identical to numpy's
    sum(lmap**2,0)**0.5)."""
    return ndmap(Geometry.get(shape, wcs, method).modlmap().copy(), wcs)
