"""orphics.maps hot-path mirror on B200: MapGen (maps.py:1553-1587), FourierCalc
(maps.py:1594-1677) and the helpers around them (rect_geometry :1472, binned_power
:1350, get_taper/cosine_window :1873-1920, filter_map :1922, gauss_beam :1925,
mask_kspace :1936), with the reference's names, arguments and return conventions.
All per-pixel arithmetic runs in liborphx.so (CUDA, sm_100a); there is no CPU
fallback.  Batched entry points (get_maps, power2d_batch, binned_power_batch,
SimPipeline) are additions: the reference treats a leading axis as polarisation
components, never as a batch (maps.py:1648,1661).
"""
import ctypes as C

import numpy as np

from . import _capi, enmap
from ._capi import lib, check, ptr, OX_HOST, OX_DEVICE
from .enmap import Geometry, ndmap, devmap, result_map


def _dev_in(x, dtype):
    """(void* argument, OX_HOST | OX_DEVICE, keep-alive) of an input array: a devmap of the plan's dtype is consumed
    in HBM, anything else goes through a C-contiguous host array of that dtype."""
    if isinstance(x, devmap) and x.dtype == np.dtype(dtype):
        return C.c_void_p(x.ptr), OX_DEVICE, x
    a = np.ascontiguousarray(np.asarray(x), dtype=dtype)
    return ptr(a), OX_HOST, a


# --------------------------------------------------------------------------- geometry
def rect_geometry(width_arcmin=None, width_deg=None, px_res_arcmin=0.5, proj="car", pol=False,
                  height_deg=None, height_arcmin=None, xoffset_degree=0., yoffset_degree=0., extra=False, **kwargs):
    """Shape and wcs of a rectangular patch centred on the given offsets (maps.py:1472-1498)."""
    if width_deg is not None:
        width_arcmin = 60. * width_deg
    if height_deg is not None:
        height_arcmin = 60. * height_deg
    hwidth = width_arcmin / 2.
    vwidth = hwidth if height_arcmin is None else height_arcmin / 2.
    arcmin, degree = enmap.arcmin, enmap.degree
    lo = [-vwidth * arcmin + yoffset_degree * degree, -hwidth * arcmin + xoffset_degree * degree]
    hi = [vwidth * arcmin + yoffset_degree * degree, hwidth * arcmin + xoffset_degree * degree]
    shape, wcs = enmap.geometry(pos=[lo, hi], res=px_res_arcmin * arcmin, proj=proj, **kwargs)
    if pol:
        shape = (3,) + shape
    if extra:
        modlmap = enmap.modlmap(shape, wcs)
        ells = np.arange(0, modlmap.max(), 1.)
        return shape, wcs, modlmap, ells
    return shape, wcs


def _ncomp(shape):
    return int(shape[-3]) if len(shape) > 2 else 1


# --------------------------------------------------------------------------- spec2flat
def _eigpow(A, e):
    """Symmetric matrix power along axes [0,1] of (n,n,...) (pixell utils.eigpow as used by
    enmap.multi_pow, maps.py:1571): negative / relatively tiny eigenvalues -> 0 for
    non-integer or negative exponents."""
    A = np.asarray(A, dtype=np.float64)
    if A.shape[0] == 1:
        a = A[0, 0]
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(a > 0, np.abs(a) ** e, 0.0)[None, None]
    n = A.shape[0]
    offdiag = ~np.eye(n, dtype=bool)
    if not A[offdiag].any():
        # diagonal matrices (e.g. white-noise spectra, lensing.py:482-487): eigenvalues are the diagonal
        D = np.stack([A[i, i] for i in range(n)])
        emax = np.max(np.abs(D), 0, keepdims=True)
        bad = (D < emax * np.finfo(np.float64).resolution * 100) | (D < np.finfo(np.float64).tiny * 1e4)
        with np.errstate(invalid="ignore", divide="ignore"):
            De = np.where(bad, 0.0, np.abs(D) ** e)
        out = np.zeros_like(A)
        for i in range(n):
            out[i, i] = De[i]
        return out
    M = np.moveaxis(A, (0, 1), (-2, -1))
    E, V = np.linalg.eigh(M)
    emax = np.max(np.abs(E), -1, keepdims=True)
    bad = (E < emax * np.finfo(np.float64).resolution * 100) | (E < np.finfo(np.float64).tiny * 1e4)
    with np.errstate(invalid="ignore", divide="ignore"):
        Ee = np.where(bad, 0.0, np.abs(E) ** e)
    return np.moveaxis(np.einsum("...ik,...k,...jk->...ij", V, Ee, V), (-2, -1), (0, 1))


def multi_pow(mat, exp):
    """enmap.multi_pow(mat, exp) on axes [0,1] of a (ncomp, ncomp, Ny, Nx) stack of symmetric matrices
    (maps.py:1571): per-pixel eigen-decomposition and power on the device (ox_multi_pow)."""
    A = np.ascontiguousarray(mat, dtype=np.float64)
    if A.ndim < 3 or A.shape[0] != A.shape[1]:
        raise ValueError("multi_pow expects (ncomp, ncomp, ...) matrices")
    if A.shape[0] > 4:
        raise NotImplementedError("multi_pow on the device supports up to 4 components")
    out = np.empty_like(A)
    check(lib.ox_multi_pow(ptr(A), A.shape[0], int(np.prod(A.shape[2:])), C.c_double(float(exp)), OX_HOST, ptr(out), OX_HOST))
    return out


try:                                   # 1-D host set-up only (the per-pixel work is on the device)
    import scipy.fft as _fft1d
except ImportError:                    # pragma: no cover
    _fft1d = np.fft


def _sym_convolve(a, b):
    # C_l l^2 spans ~10 decades, so two FFT libraries agree on this convolution only to ~1e-9 of the peak;
    # scipy.fft (pocketfft, what the oracle uses) keeps the set-up reproducible to rounding
    sa = np.concatenate([a, a[:, -2:0:-1]], -1)
    sb = np.concatenate([b, b[:, -2:0:-1]], -1)
    out = _fft1d.irfft(_fft1d.rfft(sa, axis=-1) * _fft1d.rfft(sb, axis=-1), n=sa.shape[-1], axis=-1)
    return out[:, :a.shape[-1]]


def smooth_spectrum(ps, width):
    """Mode-weighted Gaussian smoothing along l (enmap.smooth_spectrum, kernel='gauss', weight='mode')."""
    ps = np.asarray(ps, dtype=np.float64)
    flat = ps.reshape(-1, ps.shape[-1])
    l = np.arange(flat.shape[-1], dtype=np.float64)
    K = np.tile(np.exp(-0.5 * (l / width) ** 2), (flat.shape[0], 1))
    W = np.tile(l ** 2, (flat.shape[0], 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        res = _sym_convolve(flat * W, K) / _sym_convolve(W, K)
    return res.reshape(ps.shape)


def spec2flat(shape, wcs, cov, exp=1.0, mode="constant", smooth="auto", method=None):
    """enmap.spec2flat as MapGen calls it (maps.py:1573): (ncomp,ncomp,nl) -> (ncomp,ncomp,Ny,Nx).
    The short 1-D treatment (smoothing, x Npix/area, matrix power) is host set-up; the
    per-pixel order-1 interpolation at |l| runs on the device."""
    if mode != "constant":
        raise NotImplementedError("only mode='constant'")
    g = Geometry.get(shape, wcs, method)
    cov = np.array(cov, dtype=np.float64)
    if cov.ndim == 1:
        cov = cov[None, None]
    if smooth == "auto":
        smooth = 0.5 * (abs(g.ly[1]) + abs(g.lx[1])) / 3.41
    if smooth and smooth > 0:
        cov = smooth_spectrum(cov, smooth)
    cov = cov * g.npix / g.area
    if exp != 1.0:
        cov = _eigpow(cov, exp)
    cov[~np.isfinite(cov)] = 0
    return ndmap(g.interp_spec(cov), wcs)


def spec1d_to_2d(shape, wcs, ps):
    """maps.py:1591-1592."""
    return spec2flat(shape, wcs, ps) / (np.prod(shape[-2:]) / enmap.area(shape, wcs))


# --------------------------------------------------------------------------- MapGen
class MapGen(object):
    """Pre-computes covsqrt for a geometry, then draws Gaussian random fields
    (maps.py:1553-1587).

    noise: "numpy" (default) draws the white noise with numpy's global legacy RNG exactly as
    the reference does (np.random.seed(seed); rand_gauss_harm) and uploads it, so maps agree
    with the reference for identical seeds; "philox" / "philox_hermitian" draw on the
    device (throughput modes, see include/orphx.h).
    dtype: np.float64 (default) or np.float32 arithmetic."""

    def __init__(self, shape, wcs, cov=None, covsqrt=None, pixel_units=False, smooth="auto", ndown=None, order=1,
                 noise="numpy", dtype=np.float64, max_batch=1, method=None):
        self.shape = tuple(int(s) for s in shape)
        self.wcs = wcs
        self.geometry = Geometry.get(shape, wcs, method)
        if covsqrt is not None:
            self.covsqrt = covsqrt
        else:
            assert cov.ndim >= 3, "Power spectra have to be of shape (ncomp,ncomp,lmax) or (ncomp,ncomp,Ny,Nx)."
            if cov.ndim == 4:
                if not (pixel_units):
                    cov = cov * np.prod(shape[-2:]) / self.geometry.area
                if ndown:
                    raise NotImplementedError("ndown (downsample_power) is outside the accelerated path")
                self.covsqrt = ndmap(multi_pow(cov, 0.5), wcs)
            else:
                self.covsqrt = spec2flat(shape, wcs, cov, 0.5, mode="constant", smooth=smooth, method=method)
        cs = np.asarray(self.covsqrt, dtype=np.float64)
        if cs.ndim == 2:
            cs = cs[None, None]
        self.ncomp = cs.shape[0]
        if cs.shape[:2] != (self.ncomp, self.ncomp) or cs.shape[-2:] != self.geometry.shape:
            raise ValueError(f"covsqrt shape {cs.shape} does not match geometry {self.shape}")
        if _ncomp(self.shape) != self.ncomp:
            raise ValueError(f"shape {self.shape} has {_ncomp(self.shape)} components but covsqrt has {self.ncomp}")
        self.noise = noise
        self.dtype = _capi.ox_dtype(dtype)
        self.max_batch = int(max_batch)
        h = C.c_void_p()
        cs = np.ascontiguousarray(cs)
        check(lib.ox_simplan_create(self.geometry.handle, self.ncomp, ptr(cs), OX_HOST, self.dtype, self.max_batch, C.byref(h)))
        self.handle = h

    # -- noise exactly as the reference draws it (maps.py:1577-1578)
    def _numpy_noise(self, seed, out=None):
        if seed is not None:
            np.random.seed(seed)
        shp = self.shape if len(self.shape) > 2 else self.shape[-2:]
        if out is None:
            out = np.empty((2,) + (self.ncomp,) + self.geometry.shape, dtype=np.float64)
        out[0] = np.random.standard_normal(shp).reshape((self.ncomp,) + self.geometry.shape)
        out[1] = np.random.standard_normal(shp).reshape((self.ncomp,) + self.geometry.shape)
        return out

    def _generate(self, seeds, flags, noise_mode=None, noise_arrays=None):
        mode = _capi.NOISE_MODES[noise_mode or self.noise]
        nsim = len(seeds)
        if nsim > self.max_batch:
            raise ValueError(f"{nsim} sims requested but max_batch={self.max_batch}")
        noise = None
        seeds64 = np.ascontiguousarray([0 if s is None else s for s in seeds], dtype=np.int64)
        if mode == _capi.NOISE_HOST:
            if noise_arrays is not None:
                noise = np.ascontiguousarray(noise_arrays, dtype=np.float64)
            else:
                noise = np.empty((nsim, 2, self.ncomp) + self.geometry.shape, dtype=np.float64)
                for i, s in enumerate(seeds):
                    self._numpy_noise(s, noise[i])
        rdt, cdt = _capi.np_dtype(self.dtype), _capi.np_cdtype(self.dtype)
        out, optr, oloc = result_map((nsim, self.ncomp) + self.geometry.shape, cdt if flags & _capi.FLAG_HARM else rdt, self.wcs)
        check(lib.ox_sim_generate(self.handle, ptr(seeds64), nsim, mode, ptr(noise), OX_HOST, flags, optr, oloc))
        return out

    def _flags(self, scalar, iau, harm):
        flags = 0
        if harm:
            flags |= _capi.FLAG_HARM
        elif not scalar and self.ncomp > 1:
            if self.ncomp != 3:
                raise NotImplementedError("EB->QU rotation (scalar=False) needs ncomp == 3; pass scalar=True")
            flags |= _capi.FLAG_ROT
        if iau:
            flags |= _capi.FLAG_IAU
        return flags

    def get_map(self, seed=None, scalar=False, iau=False, real=False, harm=False):
        """maps.py:1576-1587."""
        if real:
            # maps.py:1578: rand = enmap.fft(enmap.rand_gauss(shape, wcs)) -- one real white-noise draw, unitary transform
            # (on the device), handed to the generator as its harmonic noise
            if seed is not None:
                np.random.seed(seed)
            shp = self.shape if len(self.shape) > 2 else self.shape[-2:]
            white = np.random.standard_normal(shp).reshape((self.ncomp,) + self.geometry.shape)
            if getattr(self, "_noise_fc", None) is None:
                self._noise_fc = FourierCalc((self.ncomp,) + self.geometry.shape if self.ncomp > 1 else self.geometry.shape, self.wcs)
            k = np.asarray(self._noise_fc.iqu2teb(white if self.ncomp > 1 else white[0], normalize=True, rot=False))
            k = k.reshape((self.ncomp,) + self.geometry.shape)
            out = self._generate([None], self._flags(scalar, iau, harm), noise_mode="numpy", noise_arrays=np.stack([k.real, k.imag])[None])[0]
        else:
            out = self._generate([seed], self._flags(scalar, iau, harm))[0]
        if len(self.shape) == 2:
            out = out[0]
        return out if isinstance(out, devmap) else ndmap(out, self.wcs)

    def get_maps(self, seeds, scalar=False, iau=False, harm=False, noise=None):
        """Batched get_map: (nsim,[ncomp,]Ny,Nx)."""
        out = self._generate(list(seeds), self._flags(scalar, iau, harm), noise_mode=noise)
        if len(self.shape) == 2:
            out = out.reshape((out.shape[0],) + out.shape[2:])
        return out if isinstance(out, devmap) else ndmap(out, self.wcs)

    def __del__(self):
        try:
            lib.ox_simplan_destroy(self.handle)
        except Exception:
            pass


# --------------------------------------------------------------------------- FourierCalc
class FourierCalc(object):
    """Pre-computes what Fourier transforms and power spectra of a geometry need
    (maps.py:1594-1677)."""

    def __init__(self, shape, wcs, iau=False, dtype=np.float64, max_batch=1, method=None):
        self.shape = tuple(int(s) for s in shape)
        self.wcs = wcs
        self.iau = iau
        self.geometry = Geometry.get(shape, wcs, method)
        self.normfact = self.geometry.area / (np.prod(self.shape[-2:]) ** 2.)
        self.ncomp = _ncomp(self.shape)
        if self.ncomp > 3:
            raise NotImplementedError("FourierCalc supports up to 3 components")
        self.dtype = _capi.ox_dtype(dtype)
        self.max_batch = int(max_batch)
        self._rot = None
        self._plans = {}
        self._retired = []

    @property
    def rot(self):
        """queb_rotmat(lmap) (maps.py:1607), built on the device on first use."""
        if self._rot is None and len(self.shape) > 2 and self.shape[-3] > 1:
            self._rot = ndmap(self.geometry.rotmat(self.iau), self.wcs)
        return self._rot

    def _plan(self, ncomp, nbatch=1):
        nb = max(self.max_batch, nbatch)
        key = ncomp
        if key in self._plans and self._plans[key][1] >= nb:
            return self._plans[key][0]
        if key in self._plans:
            self._retired.append(self._plans[key][0])  # may still be referenced by a SimPipeline
        h = C.c_void_p()
        check(lib.ox_powerplan_create(self.geometry.handle, ncomp, self.dtype, nb, C.byref(h)))
        self._plans[key] = (h, nb)
        return h

    def _as_stack(self, emap):
        """(void* argument, location, ncomp, keep-alive) of a real map (ncomp,Ny,Nx) or (Ny,Nx)."""
        shp, dt = tuple(np.shape(emap)), getattr(emap, "dtype", None)
        if dt is None:
            emap = np.asarray(emap)
            shp, dt = emap.shape, emap.dtype
        if np.issubdtype(dt, np.complexfloating):
            raise NotImplementedError("FourierCalc transforms real maps; complex input is outside the accelerated path")
        if shp[-2:] != self.geometry.shape:
            raise ValueError(f"map shape {shp} does not match geometry {self.geometry.shape}")
        nc = shp[-3] if len(shp) > 2 else 1
        if int(np.prod(shp, dtype=np.int64)) != nc * self.geometry.npix:
            raise ValueError(f"map shape {shp}: expected ([ncomp,] Ny, Nx)")
        a, loc, keep = _dev_in(emap, _capi.np_dtype(self.dtype))
        return a, loc, nc, keep

    def _flags(self, rot=True, pixel_units=False, skip_cross=False, normalize=False):
        f = 0
        if rot:
            f |= _capi.FLAG_ROT
        if pixel_units:
            f |= _capi.FLAG_PIXEL_UNITS
        if skip_cross:
            f |= _capi.FLAG_SKIP_CROSS
        if self.iau:
            f |= _capi.FLAG_IAU
        if normalize:
            f |= _capi.FLAG_UNITARY
        return f

    def iqu2teb(self, emap, nthread=0, normalize=True, rot=True):
        """2-D FFT of the map(s) with the QU->EB rotation (maps.py:1609-1617). nthread is ignored."""
        phys = isinstance(normalize, str) and normalize in ("phy", "phys", "physical")     # enmap.fft's spellings
        if not phys and normalize not in (True, False):
            raise ValueError(f"normalize={normalize!r}")
        a, loc, nc, _keep = self._as_stack(emap)
        out, optr, oloc = result_map(np.shape(emap), _capi.np_cdtype(self.dtype), self.wcs)
        flags = self._flags(rot=rot and nc >= 2, normalize=bool(normalize))     # (the last two components, maps.py:1615)
        check(lib.ox_power_fft(self._plan(nc), a, loc, 1, flags, optr, oloc))
        if phys:
            # enmap.fft(normalize="phys") = the unitary transform x pixsize^1/2 (lensing.py:403, 653)
            out = out * float(np.sqrt(self.geometry.area / self.geometry.npix))
        return out

    def f2power(self, kmap1, kmap2, pixel_units=False):
        """Re(conj(k1) k2) * normfact for already transformed maps (maps.py:1620-1624)."""
        cdt = _capi.np_cdtype(self.dtype)
        if np.shape(kmap1) != np.shape(kmap2):
            raise ValueError("kmap shapes differ")
        dev = isinstance(kmap1, devmap) and isinstance(kmap2, devmap) and kmap1.dtype == kmap2.dtype == np.dtype(cdt)
        if dev:
            a1, a2, loc = C.c_void_p(kmap1.ptr), C.c_void_p(kmap2.ptr), OX_DEVICE
        else:      # (the two k-maps share one location argument: mixed inputs go through the host)
            k1 = np.ascontiguousarray(np.asarray(kmap1), dtype=cdt)
            k2 = k1 if kmap2 is kmap1 else np.ascontiguousarray(np.asarray(kmap2), dtype=cdt)
            a1, a2, loc = ptr(k1), ptr(k2), OX_HOST
        shp = np.shape(kmap1)
        out, optr, oloc = result_map(shp, _capi.np_dtype(self.dtype), getattr(kmap1, "wcs", None))
        check(lib.ox_power_f2power(self._plan(1), a1, a2, loc, int(np.prod(shp, dtype=np.int64)),
                                   self._flags(rot=False, pixel_units=pixel_units), optr, oloc))
        return out if isinstance(out, devmap) else np.asarray(out)

    def f1power(self, map1, kmap2, pixel_units=False, nthread=0):
        """maps.py:1626-1630."""
        kmap1 = self.iqu2teb(map1, nthread, normalize=False)
        return self.f2power(kmap1, kmap2, pixel_units), kmap1

    def ifft(self, kmap):
        """Backward c2c / Npix (maps.py:1632-1633)."""
        cdt = _capi.np_cdtype(self.dtype)
        shp = np.shape(kmap)
        nc = shp[-3] if len(shp) > 2 else 1
        a, loc, _keep = _dev_in(kmap, cdt)
        out, optr, oloc = result_map(shp, cdt, self.wcs)
        check(lib.ox_power_ifft(self._plan(nc), a, loc, 1, optr, oloc))
        return out

    def fft(self, emap):
        """Raw forward FFT, no rotation (maps.py:1635-1636)."""
        return self.iqu2teb(emap, normalize=False, rot=False)

    def power2d(self, emap=None, emap2=None, nthread=0, pixel_units=False, skip_cross=False, rot=True, kmap=None,
                kmap2=None, dtype=None):
        """Power spectrum of emap crossed with emap2 (= emap if None); radians^2 unless
        pixel_units (maps.py:1639-1677)."""
        wcs = getattr(emap, "wcs", None) if emap is not None else getattr(kmap, "wcs", None)
        wcs = self.wcs if wcs is None else wcs
        if kmap is not None or kmap2 is not None:
            # already-transformed inputs: spectra from f2power, as the reference does
            lteb1 = kmap if kmap is not None else self.iqu2teb(emap, nthread, normalize=False, rot=rot)
            lteb2 = kmap2 if kmap2 is not None else (self.iqu2teb(emap2, nthread, normalize=False, rot=rot) if emap2 is not None else lteb1)
            assert np.shape(lteb1) == np.shape(lteb2)
            ndim = np.ndim(lteb1)
            ncomp = np.shape(lteb1)[-3] if ndim > 2 else 1
            if ndim > 2 and ncomp > 1:
                retpow = np.zeros((ncomp, ncomp) + np.shape(lteb1)[-2:], dtype=dtype)
                for i in range(ncomp):
                    retpow[i, i] = self.f2power(lteb1[i], lteb2[i], pixel_units)
                if not (skip_cross):
                    for i in range(ncomp):
                        for j in range(i + 1, ncomp):
                            retpow[i, j] = self.f2power(lteb1[i], lteb2[j], pixel_units)
                            retpow[j, i] = retpow[i, j]
                return retpow, lteb1, lteb2
            if ndim > 2:
                lteb1, lteb2 = lteb1[0], lteb2[0]
            keep = lambda m: m if isinstance(m, devmap) else ndmap(m, wcs)
            return keep(self.f2power(lteb1, lteb2, pixel_units)), keep(lteb1), keep(lteb2)
        a1, loc1, nc, _k1 = self._as_stack(emap)
        a2 = None
        if emap2 is not None:
            a2, loc2, nc2, _k2 = self._as_stack(emap2)
            assert np.shape(emap) == np.shape(emap2)
            if loc2 != loc1:       # (the two maps share one location argument)
                dt = _capi.np_dtype(self.dtype)
                _k1, _k2 = np.ascontiguousarray(np.asarray(emap), dtype=dt), np.ascontiguousarray(np.asarray(emap2), dtype=dt)
                a1, a2, loc1 = ptr(_k1), ptr(_k2), OX_HOST
        rdt, cdt = _capi.np_dtype(self.dtype), _capi.np_cdtype(self.dtype)
        multi = np.ndim(emap) > 2 and nc > 1
        kshape = (nc,) + self.geometry.shape if np.ndim(emap) > 2 else self.geometry.shape
        p2d, pptr, oloc = result_map((nc, nc) + self.geometry.shape if multi else self.geometry.shape, rdt, None if multi else wcs)
        k1, k1ptr, _ = result_map(kshape, cdt, wcs)
        k2, k2ptr = k1, None
        if a2 is not None:
            k2, k2ptr, _ = result_map(kshape, cdt, wcs)
        flags = self._flags(rot=rot and nc >= 2, pixel_units=pixel_units, skip_cross=skip_cross)
        check(lib.ox_power2d(self._plan(nc), a1, a2, loc1, 1, flags, pptr, k1ptr, k2ptr, oloc))
        if multi:
            # the reference returns a plain (ncomp,ncomp,Ny,Nx) array here (np.zeros, maps.py:1662)
            if dtype is not None:
                p2d = np.asarray(p2d).astype(dtype)
            elif not isinstance(p2d, devmap):
                p2d = np.asarray(p2d)
            return p2d, k1, k2
        if np.ndim(emap) > 2:
            k1, k2 = k1[0], k2[0]
        return p2d, k1, k2

    # ---- batched / fused additions
    def binned_power_batch(self, binner, maps, maps2=None, window=None, pixel_units=False, skip_cross=False, rot=True):
        """power2d + bin2D.bin for a stack (nbatch,[ncomp,]Ny,Nx) without materialising p2d;
        returns bandpowers (nbatch, nspec, nbins), nspec ordered (0,0),(0,1)..,(1,1),.. .
        ``binner`` must be a stats.bin2D built with geometry=."""
        nc = self.ncomp
        rdt = _capi.np_dtype(self.dtype)
        size = int(np.prod(np.shape(maps), dtype=np.int64))
        nb = size // (nc * self.geometry.npix)
        if nb * nc * self.geometry.npix != size or np.shape(maps)[-2:] != self.geometry.shape:
            raise ValueError(f"maps of shape {np.shape(maps)} are not a stack of ({nc},) + {self.geometry.shape} maps")
        a1, loc, _k1 = _dev_in(maps, rdt)
        a2 = None
        if maps2 is not None:
            if np.shape(maps2) != np.shape(maps):
                raise ValueError("maps2 must have the shape of maps")
            a2, loc2, _k2 = _dev_in(maps2, rdt)
            if loc2 != loc:
                _k1, _k2 = np.ascontiguousarray(np.asarray(maps), dtype=rdt), np.ascontiguousarray(np.asarray(maps2), dtype=rdt)
                a1, a2, loc = ptr(_k1), ptr(_k2), OX_HOST
        wp, wloc, _kw = (None, OX_HOST, None) if window is None else _dev_in(window, rdt)
        ns = nc if (skip_cross and nc > 1) else nc * (nc + 1) // 2
        out = np.empty((nb, ns, binner.centers.size), dtype=np.float64)
        if rot and nc == 2:
            raise NotImplementedError("binned_power_batch: the fused power+bin kernel rotates (I,Q,U) stacks only; use power2d + bin for (Q,U)")
        flags = self._flags(rot=rot and nc == 3, pixel_units=pixel_units, skip_cross=skip_cross)
        check(lib.ox_power_bin(self._plan(nc, nb), binner.handle, a1, a2, loc, nb, flags, wp, wloc, ptr(out), OX_HOST))
        # bin2D's short-bincount quirk (stats.py:796-797): same length as binner.bin() returns
        return out[..., :binner.trimmed_nbins]

    def __del__(self):
        try:
            for h, _ in self._plans.values():
                lib.ox_powerplan_destroy(h)
            for h in self._retired:
                lib.ox_powerplan_destroy(h)
        except Exception:
            pass


# --------------------------------------------------------------------------- helpers
def binned_power(imap, bin_edges=None, binner=None, fc=None, modlmap=None, imap2=None, mask=1):
    """Binned power spectrum of a map in one line (maps.py:1350-1361)."""
    from . import stats
    shape, wcs = imap.shape, imap.wcs
    fc = FourierCalc(shape, wcs) if fc is None else fc
    if binner is None:
        if modlmap is None:
            binner = stats.bin2D(fc.geometry.modlmap(), bin_edges, geometry=fc.geometry)
        else:
            binner = stats.bin2D(modlmap, bin_edges)
    if getattr(binner, "geometry", None) is fc.geometry and np.ndim(imap) == 2:
        w = None if np.isscalar(mask) and mask == 1 else np.broadcast_to(np.asarray(mask, dtype=np.float64), fc.geometry.shape)
        p1d = fc.binned_power_batch(binner, imap, imap2, window=w)[0, 0]
        return binner.centers, p1d / np.mean(np.asarray(mask, dtype=np.float64) ** 2.)
    p2d, _, _ = fc.power2d(imap * mask, imap2 * mask if imap2 is not None else None)
    cents, p1d = binner.bin(p2d)
    return cents, p1d / np.mean(mask ** 2.)


def cosine_window(Ny, Nx, lenApodY=30, lenApodX=30, padY=0, padX=0):
    """Separable raised-cosine apodisation with zero padding (maps.py:1893-1920)."""
    def edge_profile(n, lap, pad):
        prof = np.ones(n)
        if lap > 0:
            idx = np.arange(n)
            lo = idx <= (lap + pad)
            prof[lo] = 1. / 2 * (1 - np.cos(-np.pi * (idx[lo].astype(float) - pad) / lap))
            hi = idx >= ((n - 1) - lap - pad)
            prof[hi] = 1. / 2 * (1 - np.cos(-np.pi * ((n - 1) - idx[hi] - pad).astype(float) / lap))
        return prof
    win = np.ones((Ny, Nx)) * edge_profile(Nx, lenApodX, padX)[None, :]
    win = win * edge_profile(Ny, lenApodY, padY)[:, None]
    win[0:padY, :] = 0
    win[:, 0:padX] = 0
    win[Ny - padY:, :] = 0
    win[:, Nx - padX:] = 0
    return win


def get_taper(shape, wcs, taper_percent=12.0, pad_percent=3.0, weight=None):
    """maps.py:1873-1879."""
    Ny, Nx = shape[-2:]
    if weight is None:
        weight = np.ones(shape[-2:])
    apod = int(taper_percent * min(Ny, Nx) / 100.)
    pad = int(pad_percent * min(Ny, Nx) / 100.)
    taper = cosine_window(Ny, Nx, lenApodY=apod, lenApodX=apod, padY=pad, padX=pad) * weight
    w2 = np.mean(taper ** 2.)
    if enmap.DEVICE_RESIDENT:
        # host and device copies: `imap * taper` with a device-resident imap stays on the device
        return devmap.from_host(taper, wcs, copy=False), w2
    return ndmap(taper, wcs), w2


def gauss_beam(ell, fwhm):
    """maps.py:1925-1927 (fwhm in arcmin)."""
    tht_fwhm = np.deg2rad(fwhm / 60.)
    return np.exp(-(tht_fwhm ** 2.) * (ell ** 2.) / (16. * np.log(2.)))


def mask_kspace(shape, wcs, lxcut=None, lycut=None, lmin=None, lmax=None, method=None):
    """Integer l-space mask, computed on the device (maps.py:1936-1948)."""
    g = Geometry.get(shape, wcs, method)
    return ndmap(g.mask_kspace(lxcut, lycut, lmin, lmax).astype(int), wcs)


def split_calc(isplits, jsplits, icoadd, jcoadd, fourier_calc=None, alt=True):
    """Best estimate of the signal (mean of crosses) and of the noise (total - crosses) power from the
    Fourier transforms (nsplits, Ny, Nx) of windowed splits and of their coadds (maps.py:2295-2332).
    One device pass over the pixels instead of the reference's O(nsplits^2) f2power calls."""
    a = np.asarray(isplits)
    assert a.ndim == 3
    wcs = getattr(isplits, "wcs", None)
    fc = fourier_calc if fourier_calc is not None else FourierCalc(a.shape[-2:], wcs)
    b = np.asarray(jsplits)
    if alt:
        assert a.shape[0] == b.shape[0]
    cdt = np.complex128
    a, b = np.ascontiguousarray(a, dtype=cdt), np.ascontiguousarray(b, dtype=cdt)
    ic, jc = np.ascontiguousarray(icoadd, dtype=cdt), np.ascontiguousarray(jcoadd, dtype=cdt)
    if a.shape[1:] != b.shape[1:] or ic.shape != a.shape[1:] or jc.shape != a.shape[1:]:
        raise ValueError("split_calc: splits and coadds must share one (Ny, Nx) shape")
    npix = ic.size
    out = np.empty((3,) + ic.shape, dtype=np.float64)
    check(lib.ox_split_calc(ptr(a), ptr(b), ptr(ic), ptr(jc), OX_HOST, a.shape[0], b.shape[0], npix, C.c_double(fc.normfact),
                            int(bool(alt)), ptr(out[0]), ptr(out[1]), ptr(out[2]), OX_HOST))
    w = fc.wcs if wcs is None else wcs
    return ndmap(out[0], w), ndmap(out[1], w), ndmap(out[2], w)


def noise_from_splits(splits, fourier_calc=None, nthread=0, do_cross=True):
    """Noise power (auto - cross of splits)/nsplits of the I,Q,U components and the mean cross spectrum of
    the split pairs, from (nsplits, ncomp, Ny, Nx) or (nsplits, Ny, Nx) maps (maps.py:2337-2411).  As in
    the reference the splits are cast to float32 first (maps.py:2355) and its "T,E,B" cross spectrum is in
    fact the unrotated I,Q,U one (the rotation branch of maps.py:2379 can never be taken).  All
    nsplits*ncomp transforms and the O(nsplits^2) pair sums are one batched FFT and one kernel."""
    wcs = getattr(splits, "wcs", None)
    if wcs is None:
        wcs = getattr(splits[0], "wcs", None)
    arr = np.asarray(splits).astype(np.float32)
    assert arr.ndim == 3 or arr.ndim == 4
    if arr.ndim == 3:
        arr = arr[:, None, :, :]
    nsplits, ncomp = arr.shape[:2]
    if fourier_calc is None:
        fourier_calc = FourierCalc(arr.shape[-3:] if do_cross else arr.shape[-2:], wcs)
    fc = fourier_calc
    if do_cross:
        assert ncomp == 3 or ncomp == 1
    if arr.shape[-2:] != fc.geometry.shape:
        raise ValueError(f"splits of shape {arr.shape} do not match geometry {fc.geometry.shape}")
    stack = np.ascontiguousarray(arr, dtype=_capi.np_dtype(fc.dtype))
    noise = np.empty((ncomp, ncomp) + fc.geometry.shape, dtype=np.float64)
    cross = np.empty_like(noise) if do_cross else None
    check(lib.ox_noise_from_splits(fc._plan(min(ncomp, 3)), ptr(stack), OX_HOST, nsplits, ncomp, int(bool(do_cross)), ptr(noise),
                                   ptr(cross), OX_HOST))
    if ncomp == 1:
        return ndmap(noise[0, 0], wcs), (ndmap(cross[0, 0], wcs) if do_cross else None)
    return noise, cross


# ---- Fourier-space ILC (maps.py:1952-2050): per-pixel small linear algebra, one device pass each
def ilc_def_response(response, cinv):
    """Default CMB response: a vector of ones (maps.py:2007-2013)."""
    if response is None:
        response = np.ones((np.shape(cinv)[0],))
    return response


def _ilc(mode, kmaps, cinv, response_a, response_b):
    cinv = np.ascontiguousarray(cinv, dtype=np.float64)
    if cinv.ndim not in (3, 4) or cinv.shape[0] != cinv.shape[1]:
        raise ValueError("cinv must be (nfreq, nfreq, Ny, Nx) or (nfreq, nfreq, nbins)")     # ilc_index, maps.py:2015-2023
    nfreq, trail = cinv.shape[0], cinv.shape[2:]
    npix = int(np.prod(trail))
    ra = None if response_a is None else np.ascontiguousarray(response_a, dtype=np.float64)
    rb = None if response_b is None else np.ascontiguousarray(response_b, dtype=np.float64)
    for r in (ra, rb):
        if r is not None and r.shape != (nfreq,):
            raise ValueError("response vectors must have one entry per frequency")
    km = None
    if mode <= 1:
        km = np.ascontiguousarray(kmaps, dtype=np.complex128)
        if km.shape != (nfreq,) + trail:
            raise ValueError(f"kmaps of shape {km.shape} do not match cinv {cinv.shape}")
    out = np.empty(trail, dtype=np.complex128 if mode <= 1 else np.float64)
    check(lib.ox_ilc(ptr(km), ptr(cinv), ptr(ra), ptr(rb), nfreq, npix, OX_HOST, mode, ptr(out), OX_HOST))
    return out


def silc(kmaps, cinv, response=None):
    """Standard ILC of Fourier maps (nfreq, Ny, Nx) given the inverse covariance (maps.py:1952-1974)."""
    return _ilc(0, kmaps, cinv, response, None)


def cilc(kmaps, cinv, response_a, response_b):
    """Constrained ILC: component a with component b projected out (maps.py:1976-2005)."""
    return _ilc(1, kmaps, cinv, response_a, response_b)


def silc_noise(cinv, response=None):
    """maps.py:2025-2028."""
    return _ilc(2, None, cinv, response, None)


def cilc_noise(cinv, response_a, response_b):
    """maps.py:2030-2041."""
    return _ilc(3, None, cinv, response_a, response_b)


def filter_map(imap, kfilter, fc=None):
    """Re(ifft(fft(imap) * kfilter)) / Npix (maps.py:1922-1923)."""
    fc = FourierCalc(imap.shape, imap.wcs) if fc is None else fc
    a, loc, nc, _keep = fc._as_stack(imap)
    if np.issubdtype(getattr(kfilter, "dtype", np.asarray(kfilter).dtype), np.complexfloating):
        # Re(ifft(fft(m) f)) keeps k(p) 1/2 [f(p) + conj f(p')]: the same fused pass with a complex multiplier
        if isinstance(kfilter, devmap) and kfilter.dtype == np.complex128 and kfilter.shape == fc.geometry.shape:
            kp, kloc, _kk = C.c_void_p(kfilter.ptr), OX_DEVICE, kfilter
        else:
            _kk = np.ascontiguousarray(np.broadcast_to(np.asarray(kfilter, dtype=np.complex128), fc.geometry.shape))
            kp, kloc = ptr(_kk), OX_HOST
        out, optr, oloc = result_map(np.shape(imap), _capi.np_dtype(fc.dtype), getattr(imap, "wcs", fc.wcs))
        check(lib.ox_power_filter_complex(fc._plan(nc), a, loc, 1, kp, kloc, optr, oloc))
        return out
    if isinstance(kfilter, devmap) and kfilter.dtype == np.float64 and kfilter.shape == fc.geometry.shape:
        kp, kloc, _kk = C.c_void_p(kfilter.ptr), OX_DEVICE, kfilter
    else:
        _kk = np.ascontiguousarray(np.broadcast_to(np.asarray(kfilter, dtype=np.float64), fc.geometry.shape))
        kp, kloc = ptr(_kk), OX_HOST
    out, optr, oloc = result_map(np.shape(imap), _capi.np_dtype(fc.dtype), getattr(imap, "wcs", fc.wcs))
    # one device pass: r2c -> x 1/2[f(l)+f(-l)]/Npix -> c2r (ox_power_filter)
    check(lib.ox_power_filter(fc._plan(nc), a, loc, 1, kp, kloc, optr, oloc))
    return out


# --------------------------------------------------------------------------- fused pipeline
class SimPipeline(object):
    """sim -> FFT -> power2d -> bin2D for batches of seeds in one device-resident pass
    (ox_pipeline_*): MapGen.get_map, an optional real-space taper, FourierCalc.power2d and
    bin2D.bin, plus the Statistics triple of the bandpower vectors accumulated on the
    device.  Bandpowers come back as (nsim, nspec, nbins)."""

    def __init__(self, mapgen, fc, binner, window=None):
        if binner.geometry is None:
            raise ValueError("SimPipeline needs a bin2D built with geometry=")
        self.mapgen, self.fc, self.binner = mapgen, fc, binner
        self.window = None if window is None else np.ascontiguousarray(window, dtype=np.float64)
        self._pplan = fc._plan(mapgen.ncomp, mapgen.max_batch)
        h = C.c_void_p()
        check(lib.ox_pipeline_create(mapgen.handle, self._pplan, binner.handle, ptr(self.window), OX_HOST, C.byref(h)))
        self.handle = h
        self.nspec = mapgen.ncomp * (mapgen.ncomp + 1) // 2
        self.nbins = binner.centers.size
        self.dim = self.nspec * self.nbins
        path = C.c_int(0)
        check(lib.ox_pipeline_path(self.handle, C.byref(path)))
        #: "fused" = hand-written FFT kernels (power-of-two maps), "cufft" = cuFFT passes
        self.path = {1: "cufft", 2: "fused"}[path.value]

    def last_maps(self, nsim):
        """Real-space maps (before the taper) of the last run of nsim sims (needs
        keep_maps=True on the fused path)."""
        p = C.c_void_p()
        check(lib.ox_pipeline_maps(self.handle, C.byref(p)))
        if not p.value:
            raise RuntimeError("last_maps: the last run did not keep its real-space maps (fused path: pass keep_maps=True)")
        mg = self.mapgen
        out = np.empty((nsim, mg.ncomp) + mg.geometry.shape, dtype=_capi.np_dtype(mg.dtype))
        check(lib.ox_memcpy_d2h(ptr(out), p, out.nbytes))
        return out

    def _flags(self, scalar, iau, keep_maps=False):
        f = _capi.FLAG_KEEP_MAPS if keep_maps else 0
        if not scalar and self.mapgen.ncomp == 3:
            f |= _capi.FLAG_ROT
        if iau:
            f |= _capi.FLAG_IAU
        return f

    def run(self, seeds, scalar=False, iau=False, noise=None, out=None, fetch=True, keep_maps=False):
        """Bandpowers of len(seeds) sims.  noise: None -> the MapGen's mode."""
        mg = self.mapgen
        mode = _capi.NOISE_MODES[noise or mg.noise]
        seeds = list(seeds)
        nsim = len(seeds)
        res = np.empty((nsim, self.nspec, self.nbins), dtype=np.float64) if out is None else out
        flags = self._flags(scalar, iau, keep_maps)
        done = 0
        while done < nsim:
            n = min(mg.max_batch, nsim - done)
            chunk = seeds[done:done + n]
            s64 = np.ascontiguousarray(chunk, dtype=np.int64)
            nz = None
            if mode == _capi.NOISE_HOST:
                nz = np.empty((n, 2, mg.ncomp) + mg.geometry.shape, dtype=np.float64)
                for i, s in enumerate(chunk):
                    mg._numpy_noise(s, nz[i])
            dst = res[done:done + n]
            check(lib.ox_pipeline_run(self.handle, ptr(s64), n, mode, ptr(nz), OX_HOST, flags,
                                      ptr(dst) if fetch else None, OX_HOST))
            done += n
        # bin2D's short-bincount quirk (stats.py:796-797): the bandpowers handed back have the length
        # binner.bin() returns; the device Statistics triple keeps the full nbins-long vectors
        return res[..., :self.binner.trimmed_nbins] if fetch else None

    def run_raw(self, seeds_i64, nsim, mode, flags, out_ptr):
        """Thin call for benchmarks: seeds_i64 is a contiguous int64 numpy array (e.g. pinned)."""
        check(lib.ox_pipeline_run(self.handle, ptr(seeds_i64), nsim, mode, None, OX_HOST, flags, out_ptr, OX_HOST))

    STAGES = ("sim_fill", "cufft_inverse", "window", "cufft_forward", "power_bin", "statistics")

    def profile(self, seeds_i64, nsim, mode, flags):
        """Per-stage device times (ms) of one run, from CUDA events between the stages."""
        ms = (C.c_float * 6)()
        check(lib.ox_pipeline_profile(self.handle, ptr(seeds_i64), nsim, mode, flags, ms))
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def stats_pointer(self):
        """Device address of the packed float64 [N | SUM[dim] | CROSS[dim][dim]] accumulator and dim."""
        p = C.c_void_p()
        d = C.c_int()
        check(lib.ox_pipeline_stats(self.handle, C.byref(p), C.byref(d)))
        return p.value, d.value

    def stats(self):
        """(N, SUM, CROSS) accumulated on the device so far."""
        p, d = self.stats_pointer()
        packed = np.empty(1 + d + d * d, dtype=np.float64)
        check(lib.ox_memcpy_d2h(ptr(packed), C.c_void_p(p), packed.nbytes))
        return int(round(packed[0])), packed[1:1 + d].copy(), packed[1 + d:].reshape(d, d).copy()

    def allreduce(self, comm):
        """Sum the Statistics triple over the ranks of an mpi.Comm: ONE ncclAllReduce (stats.py:1215-1217)."""
        check(lib.ox_pipeline_allreduce(comm.handle, self.handle))

    def reset_stats(self):
        check(lib.ox_pipeline_stats_reset(self.handle))

    def __del__(self):
        try:
            lib.ox_pipeline_destroy(self.handle)
        except Exception:
            pass
