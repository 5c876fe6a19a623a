"""Host-side geometry for the flat-sky path: a minimal stand-in for the bits of
pixell.enmap that orphics.maps touches at set-up time (shape/wcs bookkeeping, the
Fourier axes and the patch area).  Everything per-pixel is computed on the device
through the Geometry handle (liborphx ox_geometry_*).

Reference call sites: enmap.geometry maps.py:1490; enmap.area maps.py:1567,1605;
enmap.lmap/modlmap/laxes maps.py:1374,1607,1938-1939.
"""
import ctypes as C
import os

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


#: True: the per-pixel results of the reference-signature calls (MapGen.get_map, FourierCalc.iqu2teb / fft / ifft /
#: f2power / power2d, filter_map, get_taper, qest.kappa_from_map) come back as ``devmap``: arrays that live in HBM,
#: are accepted as such by every one of those calls and by bin2D.bin, and turn into the numpy ``ndmap`` the
#: reference returns on first host access.  ORPHX_DEVICE_MAPS=0 restores plain host ndmaps.
DEVICE_RESIDENT = os.environ.get("ORPHX_DEVICE_MAPS", "1") != "0"


class _PoolBuf:
    """Owner of one block of the library's stream-ordered device pool (ox_malloc_pooled)."""
    __slots__ = ("ptr", "nbytes")

    def __init__(self, nbytes):
        _capi.require_device()
        p = C.c_void_p()
        check(lib.ox_malloc_pooled(C.byref(p), int(nbytes)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                lib.ox_free_pooled(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


_OPS = {"mul": 0, "add": 1, "sub": 2, "div": 3, "rsub": 4, "rdiv": 5}
_UFUNC_OPS = {np.multiply: "mul", np.add: "add", np.subtract: "sub", np.true_divide: "div"}
_REFLECT = {"mul": "mul", "add": "add", "sub": "rsub", "div": "rdiv"}


class devmap(object):
    """A map that lives in device memory and stands in for the ``ndmap`` the reference returns
    (maps.py:1585-1587, 1617, 1677): shape / dtype / wcs / indexing / arithmetic as numpy has them.

    * Handed to another call of this package (power2d, f2power, ifft, filter_map, bin2D.bin, kappa_from_map, ...)
      it is consumed where it is: no PCIe transfer.
    * ``emap * taper``, ``a + b``, ``p2d / w2`` between devmaps (or with scalars / host arrays) run on the device
      (ox_map_op) and return devmaps.
    * Anything else -- np.asarray, slicing, reductions, attribute access -- materialises the host copy once
      (``__array__``) and behaves as the numpy ndmap; results of such operations are host ndmaps.
    devmaps are immutable on the device: item assignment edits the host copy and re-uploads on next device use."""

    __array_priority__ = 1000.0

    def __init__(self, owner, ptr_, shape, dtype, wcs=None, host=None):
        self._owner = owner          # _PoolBuf (shared by views) or None when only the host copy exists
        self._ptr = ptr_
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        self.wcs = wcs
        self._host = host            # ndmap or None
        self._dev_valid = ptr_ is not None

    # ---- construction
    @classmethod
    def empty(cls, shape, dtype, wcs=None):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        buf = _PoolBuf(n)
        return cls(buf, buf.ptr, shape, dtype, wcs)

    @classmethod
    def from_host(cls, arr, wcs=None, copy=True):
        """Both copies valid: the array is uploaded now and a private copy (copy=False: the array itself, for arrays
        the caller hands over) is kept as the host view -- later edits of the caller's array must not reach one copy only."""
        wcs = getattr(arr, "wcs", None) if wcs is None else wcs
        a = np.array(arr, copy=True, order="C") if copy else np.ascontiguousarray(arr)
        out = cls.empty(a.shape, a.dtype, wcs)
        check(lib.ox_memcpy_h2d(C.c_void_p(out._ptr), ptr(a), a.nbytes))
        out._host = ndmap(a, wcs)
        return out

    # ---- numpy-like surface
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def ptr(self):
        """Device address (uploads the host copy first if that is the newer one)."""
        if not self._dev_valid:
            a = np.ascontiguousarray(self._host)
            if self._owner is None or self._owner.nbytes < a.nbytes:
                self._owner = _PoolBuf(a.nbytes)
                self._ptr = self._owner.ptr
            check(lib.ox_memcpy_h2d(C.c_void_p(self._ptr), ptr(a), a.nbytes))
            self._dev_valid = True
        return self._ptr

    def host(self):
        """The numpy ndmap (device -> host once, then cached)."""
        if self._host is None:
            out = np.empty(self.shape, dtype=self.dtype)
            check(lib.ox_memcpy_d2h(ptr(out), C.c_void_p(self._ptr), out.nbytes))
            self._host = ndmap(out, self.wcs)
        # read-only: writing into what np.asarray() hands out would change the host copy behind the device copy's back;
        # item assignment on the devmap itself (which re-uploads) and np.array(m) (a copy) are the ways to edit
        self._host.flags.writeable = False
        return self._host

    def __array__(self, dtype=None, copy=None):
        h = self.host()
        if dtype is not None and np.dtype(dtype) != h.dtype:
            return np.asarray(h).astype(dtype)
        return np.array(h, copy=True) if copy else np.asarray(h)

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __getitem__(self, idx):
        # a leading-axis integer (kmapTEB[0], maps[i]) is a view of the same device block; anything else is numpy's
        if isinstance(idx, (int, np.integer)) and self.ndim >= 1 and self._dev_valid and self.ndim > 2:
            i = int(idx)
            if i < 0:
                i += self.shape[0]
            if not 0 <= i < self.shape[0]:
                raise IndexError(f"index {idx} is out of bounds for axis 0 with size {self.shape[0]}")
            sub = self.shape[1:]
            step = int(np.prod(sub, dtype=np.int64)) * self.dtype.itemsize
            h = None if self._host is None else self._host[i]
            return devmap(self._owner, self._ptr + i * step, sub, self.dtype, self.wcs, host=h)
        return self.host()[idx]

    def __setitem__(self, idx, value):
        h = self.host()
        h.flags.writeable = True
        try:
            h[idx] = value
        finally:
            h.flags.writeable = False
        self._dev_valid = False      # the host copy is now the newer one
        if self._owner is not None and self._ptr != self._owner.ptr:
            self._owner = None       # (a view: never write through into the parent's block)

    def copy(self):
        out = devmap.empty(self.shape, self.dtype, self.wcs)
        check(lib.ox_memcpy_d2d(C.c_void_p(out._ptr), C.c_void_p(self.ptr), self.nbytes))
        return out

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)) else shape
        shape = tuple(int(v) for v in shape)
        if -1 in shape:
            known = -int(np.prod(shape, dtype=np.int64))
            shape = tuple(self.size // known if v == -1 else v for v in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {shape}")
        if not self._dev_valid:
            return self.host().reshape(shape)
        return devmap(self._owner, self._ptr, shape, self.dtype, self.wcs,
                      host=None if self._host is None else self._host.reshape(shape))

    def __getattr__(self, name):
        # everything numpy offers that is not provided above (mean, sum, real, T, astype, ...) -> the host ndmap
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.host(), name)

    def __repr__(self):
        return "devmap(" + repr(self.host()) + ")"

    # ---- arithmetic on the device
    def _kind(self):
        return {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.complex128): 2, np.dtype(np.complex64): 3}.get(self.dtype)

    def _device_op(self, op, other):
        """self (op) other on the device, or NotImplemented when numpy's own rules are needed (type
        promotion, general broadcasting, integer / complex operands)."""
        kind = self._kind()
        if kind is None:
            return NotImplemented
        real = np.dtype(np.float32) if kind in (1, 3) else np.dtype(np.float64)
        if kind >= 2 and op == "rdiv":
            return NotImplemented
        if isinstance(other, (bool, np.bool_)):
            return NotImplemented
        if isinstance(other, (int, float, np.integer, np.floating)):
            if isinstance(other, np.floating) and other.dtype.itemsize > real.itemsize:
                return NotImplemented            # float32 map (op) np.float64 scalar promotes in numpy 2
            out = devmap.empty(self.shape, self.dtype, self.wcs)
            check(lib.ox_map_op(_OPS[op], C.c_void_p(self.ptr), None, C.c_double(float(other)), self.size, 0, kind, C.c_void_p(out._ptr)))
            return out
        if isinstance(other, devmap):
            b = other
        elif isinstance(other, np.ndarray):
            if other.dtype != real or other.ndim == 0:
                return NotImplemented
            b = None
        else:
            return NotImplemented
        oshape = tuple(other.shape)
        if np.dtype(other.dtype) != real:
            return NotImplemented
        # the operand must be the trailing axes of self (equal shapes, or a (Ny,Nx) taper against (ncomp,Ny,Nx))
        if len(oshape) > self.ndim or oshape != self.shape[self.ndim - len(oshape):] or not oshape:
            return NotImplemented
        if b is None:
            b = devmap.from_host(other, copy=False)      # (a temporary for this one operation)
        out = devmap.empty(self.shape, self.dtype, self.wcs if self.wcs is not None else getattr(other, "wcs", None))
        check(lib.ox_map_op(_OPS[op], C.c_void_p(self.ptr), C.c_void_p(b.ptr), C.c_double(0.0), self.size, b.size, kind,
                            C.c_void_p(out._ptr)))
        return out

    def _binop(self, op, other, ufunc, reflected):
        r = self._device_op(_REFLECT[op] if reflected else op, other)
        if r is not NotImplemented:
            return r
        o = other.host() if isinstance(other, devmap) else other
        return ufunc(o, self.host()) if reflected else ufunc(self.host(), o)

    def __mul__(self, o): return self._binop("mul", o, np.multiply, False)
    def __rmul__(self, o): return self._binop("mul", o, np.multiply, True)
    def __add__(self, o): return self._binop("add", o, np.add, False)
    def __radd__(self, o): return self._binop("add", o, np.add, True)
    def __sub__(self, o): return self._binop("sub", o, np.subtract, False)
    def __rsub__(self, o): return self._binop("sub", o, np.subtract, True)
    def __truediv__(self, o): return self._binop("div", o, np.true_divide, False)
    def __rtruediv__(self, o): return self._binop("div", o, np.true_divide, True)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # host_array * devmap etc. arrive here; the four arithmetic ufuncs stay on the device, the rest is numpy's
        if method == "__call__" and not kwargs and len(inputs) == 2 and ufunc in _UFUNC_OPS:
            a, b = inputs
            op = _UFUNC_OPS[ufunc]
            if isinstance(a, devmap):
                r = a._device_op(op, b)
            else:
                r = b._device_op(_REFLECT[op], a)
            if r is not NotImplemented:
                return r
        outs = kwargs.get("out", ())
        if outs:
            hosts = []
            for o in outs:
                if isinstance(o, devmap):
                    h = o.host()
                    h.flags.writeable = True
                    hosts.append(h)
                else:
                    hosts.append(o)
            kwargs["out"] = tuple(hosts)
        args = [x.host() if isinstance(x, devmap) else x for x in inputs]
        try:
            res = getattr(ufunc, method)(*args, **kwargs)
        finally:
            for o in outs:
                if isinstance(o, devmap):     # written on the host: the device copy is stale now
                    o._host.flags.writeable = False
                    o._dev_valid = False
                    if o._owner is not None and o._ptr != o._owner.ptr:
                        o._owner = None
        return res


def _host_fallback(name):
    def f(self, *args):
        args = [a.host() if isinstance(a, devmap) else a for a in args]
        return getattr(self.host(), name)(*args)
    f.__name__ = name
    return f


for _n in ("__neg__", "__pos__", "__abs__", "__pow__", "__rpow__", "__floordiv__", "__rfloordiv__", "__mod__", "__rmod__",
           "__matmul__", "__rmatmul__", "__lt__", "__le__", "__gt__", "__ge__", "__eq__", "__ne__", "__and__", "__or__",
           "__xor__", "__invert__", "__bool__", "__float__", "__int__", "__complex__", "__contains__"):
    setattr(devmap, _n, _host_fallback(_n))
devmap.__hash__ = None


def to_device(x, dtype=None, wcs=None):
    """devmap of x: x itself when it already is one of that dtype, else an upload (both copies kept)."""
    if isinstance(x, devmap) and (dtype is None or x.dtype == np.dtype(dtype)):
        return x
    a = np.asarray(x) if dtype is None else np.asarray(x, dtype=dtype)
    aliases = isinstance(x, np.ndarray) and np.shares_memory(a, x)      # (a fresh conversion needs no second copy)
    return devmap.from_host(a, getattr(x, "wcs", None) if wcs is None else wcs, copy=aliases)


def result_map(shape, dtype, wcs):
    """(array-or-devmap, void* argument, OX_HOST | OX_DEVICE) for an output of a reference-signature call."""
    if DEVICE_RESIDENT:
        d = devmap.empty(shape, dtype, wcs)
        return d, C.c_void_p(d._ptr), _capi.OX_DEVICE
    a = np.empty(shape, dtype=dtype)
    return ndmap(a, wcs), ptr(a), _capi.OX_HOST


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
