"""orphics.lensing hot-path mirror: the Hu-Okamoto flat-sky quadratic estimator ``qest``.

The class is absent from the reference snapshot; its interface is taken from the surviving
call sites -- constructor tutorials/tt_verification.ipynb:81, ``kappa_from_map``
tutorials/tt_verification.ipynb:608,610 and lensing.py:973-976, ``.N.Nlkk[XY]`` -- and the
arithmetic from the historical estimator.  ``reconstruct`` is the alias
BASELINE.json names.  Filters and the normalisation A_L are built once per geometry (host
arithmetic on 2-D arrays + device FFTs); every per-realisation step runs in liborphx.so
(ox_qe_reconstruct): fused leg/product/divergence kernels around cuFFT transforms.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib, check, ptr, OX_HOST
from . import enmap as _enmap
from .enmap import Geometry, ndmap, devmap

#: bound on max|kappa - kappa_ref| / max|kappa_ref| of the float32 estimator against the float64 reference chain
#: on the same (float32-representable) inputs: north_star's 1e-5.  Measured on B200: 7e-7 (TT) / 8e-7 (EB) at 512^2
#: on both the hand-written and the cuFFT chains (tools/diag_fp32_qe.py, profiles/r02_fp32_qe.txt)
QE_FP32_BOUND = 1e-5


def _fmask(arr, mask):
    arr = arr.copy()
    if mask is not None:
        arr[np.asarray(mask) < 1.e-3] = 0.
    return arr


def _dense_theory_tables(theory, L_max):
    """1-D tables t[which+key][l], l = 0..lmax, equal to theory.lCl / uCl on a dense integer grid, when the theory is
    piecewise linear between integer knots (the CAMB tables, cosmology.py:888-943): linear interpolation of the dense
    tables at |l| is then the theory's own interpolation.  None when that cannot be guaranteed."""
    from .cosmology import TheorySpectra
    if not isinstance(theory, TheorySpectra):
        return None
    out, lo = {}, None
    for which in ("l", "u"):
        for key in ("TT", "EE", "BB", "TE"):
            if key not in theory._tables[which]:
                return None
            x, y = theory._tables[which][key]
            if x.size < 2 or np.any(x != np.round(x)) or np.any(np.diff(x) <= 0):
                return None
            hi = int(min(x[-1], np.ceil(L_max) + 1))
            ell = np.arange(0, hi + 1, dtype=np.float64)
            out[which + key] = np.interp(ell, x, y, left=0., right=0.)
            lo = x[0] if lo is None else max(lo, x[0])
    return out, lo


class QuadNorm(object):
    """Filters W_XY, W_Y and the normalisation on the 2-D Fourier grid (historical QuadNorm).

    Device set-up (default): the 2-D spectra are interpolated on the device, the filters are one kernel each
    (ox_qe_filter) and A_L is 13 / 24 cuFFT transforms with the products between them in HBM (ox_qe_norm); the
    tables are device-resident maps that turn into numpy arrays on host access.  ORPHX_QE_SETUP=host (or a theory
    object that is not the package's piecewise-linear TheorySpectra) keeps the numpy arithmetic of round 1."""

    def __init__(self, shape, wcs, theory, noise2d, noise2d_P, noise2d_B, beam2d, kmask, kmask_P, kmask_K, grad_cut,
                 unlensed_equals_lensed, bigell, geometry, fft_plan):
        import os
        g = self.geometry = geometry
        self._plan = fft_plan
        self.shape, self.wcs = tuple(shape[-2:]), wcs
        ext = _enmap.extent(g.shape, wcs, method=g.method)
        self.pixScaleY, self.pixScaleX = ext[0] / g.shape[0], ext[1] / g.shape[1]
        self.bigell = bigell
        self.gradCut = bigell if grad_cut is None else grad_cut
        self.Nlkk, self.AL = {}, {}
        self._lazy = {}
        dense = None
        if _enmap.DEVICE_RESIDENT and os.environ.get("ORPHX_QE_SETUP", "device") != "host":
            lmax = float(np.hypot(np.abs(g.ly).max(), np.abs(g.lx).max()))
            dense = _dense_theory_tables(theory, lmax)
            if dense is not None:
                # no pixel may fall strictly between 0 and the first knot, where the dense table ramps and the theory is 0
                nz = np.concatenate([np.abs(g.ly)[np.abs(g.ly) > 0], np.abs(g.lx)[np.abs(g.lx) > 0]])
                if nz.size and nz.min() < dense[1]:
                    dense = None
        self.device = dense is not None
        if self.device:
            tabs = dense[0]
            keys = ("TT", "EE", "BB", "TE")
            n = max(t.size for t in tabs.values())
            stack = np.zeros((8, n))
            for i, k in enumerate(keys):
                stack[i, :tabs["l" + k].size] = tabs["l" + k]
                stack[4 + i, :tabs["u" + k].size] = tabs["u" + k]
            planes = devmap.empty((8,) + g.shape, np.float64, wcs)
            check(lib.ox_geometry_interp_spec(g.handle, ptr(stack), 8, n, C.c_void_p(planes.ptr), _capi.OX_DEVICE))
            self.lClFid2d = {k: planes[i] for i, k in enumerate(keys)}
            self.uClFid2d = dict(self.lClFid2d) if unlensed_equals_lensed else {k: planes[4 + i] for i, k in enumerate(keys)}
            up = lambda a: _enmap.to_device(np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), g.shape)), np.float64, wcs)
            self.noise = {"TT": None if noise2d is None else up(noise2d)}
            if noise2d_P is not None:
                self.noise["EE"] = up(noise2d_P)
            else:
                self.noise["EE"] = None if self.noise["TT"] is None else self.noise["TT"] * 2.
            self.noise["BB"] = self.noise["EE"] if noise2d_B is None else up(noise2d_B)
            self.kBeam = None if beam2d is None else up(beam2d)
            self.fMask = {"TT": None if kmask is None else up(kmask), "EE": None if kmask_P is None else up(kmask_P)}
            self.fMask["BB"] = self.fMask["EE"]
            self.fmaskK = None if kmask_K is None else up(kmask_K)
            return
        L = self.modLMap
        self.uClFid2d = {k: (theory.lCl(k, L) if unlensed_equals_lensed else theory.uCl(k, L)) for k in ("TT", "EE", "BB", "TE")}
        self.lClFid2d = {k: theory.lCl(k, L) for k in ("TT", "EE", "BB", "TE")}
        z = np.zeros(self.shape)
        self.noise = {"TT": z + (0. if noise2d is None else noise2d)}
        self.noise["EE"] = 2. * self.noise["TT"] if noise2d_P is None else z + noise2d_P
        self.noise["BB"] = self.noise["EE"] if noise2d_B is None else z + noise2d_B
        self.kBeam = z + (1. if beam2d is None else beam2d)
        self.fMask = {"TT": kmask, "EE": kmask_P, "BB": kmask_P}
        self.fmaskK = kmask_K

    # 2-D coordinate arrays of the host route, built on first use (the device route never needs them)
    def _coord(self, name):
        if name not in self._lazy:
            g = self.geometry
            lyMap, lxMap = np.meshgrid(g.ly, g.lx, indexing="ij")
            modL = g.modlmap()
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = np.nan_to_num(1. / modL, posinf=0., neginf=0.)
            self._lazy.update(lyMap=lyMap, lxMap=lxMap, modLMap=modL, lxHatMap=lxMap * inv, lyHatMap=lyMap * inv)
        return self._lazy[name]

    lyMap = property(lambda self: self._coord("lyMap"))
    lxMap = property(lambda self: self._coord("lxMap"))
    modLMap = property(lambda self: self._coord("modLMap"))
    lxHatMap = property(lambda self: self._coord("lxHatMap"))
    lyHatMap = property(lambda self: self._coord("lyHatMap"))

    # device transforms of full-plane complex arrays: raw forward, backward / Npix (lensing.py:20)
    def _fft(self, a, inverse=False):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        out = np.empty_like(a)
        nplanes = a.size // (self.shape[0] * self.shape[1])
        scale = 1.0 / (self.shape[0] * self.shape[1]) if inverse else 1.0
        check(lib.ox_fft_c2c(self._plan, ptr(a), OX_HOST, nplanes, 1 if inverse else -1, scale, ptr(out), OX_HOST))
        return out

    def _filter_dev(self, num, XX, cut_gt):
        p = lambda m: None if m is None else C.c_void_p(m.ptr)
        out = devmap.empty(self.shape, np.float64, self.wcs)
        check(lib.ox_qe_filter(self.geometry.handle, p(num), p(self.lClFid2d[XX]), p(self.noise[XX]), p(self.kBeam), p(self.fMask[XX]),
                               C.c_double(cut_gt), C.c_double(self.bigell), C.c_void_p(out.ptr)))
        return out

    def WXY(self, XY):
        X, Y = XY
        if Y == 'B':
            Y = 'E'
        if self.device:
            return self._filter_dev(self.uClFid2d[X + Y], X + X, float(self.gradCut))
        with np.errstate(divide="ignore", invalid="ignore"):
            tot = self.lClFid2d[X + X] * self.kBeam ** 2. + self.noise[X + X]
            W = _fmask(np.nan_to_num(self.uClFid2d[X + Y] / tot, posinf=0., neginf=0.) * self.kBeam, self.fMask[X + X])
        W[self.modLMap > self.gradCut] = 0.
        W[self.modLMap >= self.bigell] = 0.
        return W

    def WY(self, YY):
        if self.device:
            return self._filter_dev(None, YY, float("inf"))
        with np.errstate(divide="ignore", invalid="ignore"):
            tot = self.lClFid2d[YY] * self.kBeam ** 2. + self.noise[YY]
            W = _fmask(np.nan_to_num(1. / tot, posinf=0., neginf=0.) * self.kBeam, self.fMask[YY])
        W[self.modLMap >= self.bigell] = 0.
        return W

    def getNlkk2d(self, XY):
        if XY not in ('TT', 'EB'):
            raise NotImplementedError(f"estimator {XY!r} (TT and EB are on the accelerated path)")
        if self.device:
            W1, W2 = self.WXY(XY), self.WY('TT' if XY == 'TT' else 'BB')
            if self.kBeam is not None:
                W1, W2 = W1 * self.kBeam, W2 * self.kBeam
            Cl = self.uClFid2d['TT' if XY == 'TT' else 'EE']
            nl = devmap.empty(self.shape, np.float64, self.wcs)
            al = devmap.empty(self.shape, np.float64, self.wcs)
            check(lib.ox_qe_norm(self.geometry.handle, _capi.QE_TT if XY == 'TT' else _capi.QE_EB, C.c_void_p(Cl.ptr), C.c_void_p(W1.ptr),
                                 C.c_void_p(W2.ptr), None if self.fmaskK is None else C.c_void_p(self.fmaskK.ptr),
                                 C.c_double(self.bigell), C.c_double(self.pixScaleX * self.pixScaleY), C.c_void_p(nl.ptr), C.c_void_p(al.ptr)))
            self.Nlkk[XY], self.AL[XY] = nl, al
            return al
        lx, ly, L = self.lxMap, self.lyMap, self.modLMap
        ifft = lambda a: self._fft(a, inverse=True)
        if XY == 'TT':
            Cl = self.uClFid2d['TT']
            WXY, WY = self.WXY('TT') * self.kBeam, self.WY('TT') * self.kBeam
            rfact = 2. ** 0.25
            g0 = ifft(WY)
            acc = 0.
            for ell1, ell2 in ((lx, lx), (ly, ly), (rfact * lx, rfact * ly)):
                f = ifft(np.stack([ell1 * ell2 * Cl * WXY, ell1 * WXY, ell2 * Cl * WY]))
                acc = acc + ell1 * ell2 * self._fft(f[0] * g0 + f[1] * f[2])
        else:
            Cl = self.uClFid2d['EE']
            s2, c2 = 2. * self.lxHatMap * self.lyHatMap, self.lyHatMap ** 2 - self.lxHatMap ** 2
            fF = (s2 ** 2., c2 ** 2., 1.j * np.sqrt(2.) * s2 * c2)
            fG = (c2 ** 2., s2 ** 2., 1.j * np.sqrt(2.) * s2 * c2)
            WXY, WY = self.WXY('EB') * self.kBeam, self.WY('BB') * self.kBeam
            gs = ifft(np.stack([WY * b for b in fG]))
            acc = 0.
            for ellsq in (lx * lx, ly * ly, np.sqrt(2.) * lx * ly):
                fs = ifft(np.stack([ellsq * Cl * WXY * a for a in fF]))
                acc = acc + ellsq * self._fft(fs[0] * gs[0] + fs[1] * gs[1] + fs[2] * gs[2])
        with np.errstate(divide="ignore", invalid="ignore"):
            alval = np.nan_to_num(1. / np.real(acc), posinf=0., neginf=0.)
        alval = _fmask(alval, self.fmaskK)
        NL = (L ** 2.) * ((L + 1.) ** 2.) * alval / 4.
        NL[(L >= self.bigell) | (L < 2.)] = 0.
        retval = np.nan_to_num(NL.real * self.pixScaleX * self.pixScaleY)
        self.Nlkk[XY] = retval.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            self.AL[XY] = retval * 2. * np.nan_to_num(1. / L / (L + 1.), posinf=0., neginf=0.)
        return self.AL[XY]


def _symmetric(a, rtol=0.):
    """a(l) == a(-l) on the FFT grid (to rtol of the peak: A_L comes out of FFTs, symmetric to rounding)."""
    ny, nx = a.shape
    iy, ix = (-np.arange(ny)) % ny, (-np.arange(nx)) % nx
    return bool(np.max(np.abs(a - a[iy][:, ix])) <= rtol * np.max(np.abs(a)))


class qest(object):
    def __init__(self, shape, wcs, theory, noise2d=None, beam2d=None, kmask=None, noise2d_P=None, kmask_P=None,
                 kmask_K=None, pol=False, grad_cut=None, unlensed_equals_lensed=False, bigell=9000, noise2d_B=None,
                 noise_keys2d=None, dtype=np.float64, max_batch=1, method=None, quadnorm=None):
        _capi.require_device()
        self.shape, self.wcs = tuple(int(s) for s in shape), wcs
        self.geometry = Geometry.get(shape, wcs, method)
        self.dtype = _capi.ox_dtype(dtype)
        self.max_batch = int(max_batch)
        self.pol = pol
        self._fftplan = C.c_void_p()   # float64 single-component plan for the set-up transforms
        check(lib.ox_powerplan_create(self.geometry.handle, 1, _capi.OX_F64, 4, C.byref(self._fftplan)))
        # quadnorm: filters and A_L of another qest on the same geometry (e.g. the float64 estimator's, reused
        # by a float32 one) instead of building them again
        self.N = quadnorm if quadnorm is not None else QuadNorm(
            shape, wcs, theory, noise2d, noise2d_P, noise2d_B, beam2d, kmask, kmask_P, kmask_K, grad_cut,
            unlensed_equals_lensed, bigell, self.geometry, self._fftplan)
        self._plans = {}
        for XY in (('TT', 'EB') if pol else ('TT',)):
            self._make_plan(XY)

    def _make_plan(self, XY):
        N = self.N
        AL = N.AL[XY] if XY in N.AL else N.getNlkk2d(XY)
        ny, nx = self.geometry.shape
        h = C.c_void_p()
        est = _capi.QE_TT if XY == 'TT' else _capi.QE_EB
        if getattr(N, "device", False) and isinstance(AL, devmap):
            # tables stay in HBM: filters, the masked normalisation, the symmetry test that selects the half-plane paths
            wxy, wy = N.WXY(XY), N.WY(XY[1] + XY[1])
            norm = AL                      # (ox_qe_norm has applied kmask_K and nan_to_num already)

            def sym(a, rtol):
                r = (C.c_double * 3)()
                check(lib.ox_plane_symmetry(self.geometry.handle, C.c_void_p(a.ptr), r))
                return r[0] <= rtol * r[1], r[2] != 0.0
            (s1, n1), (s2, n2), (s3, _) = sym(wxy, 0.), sym(wy, 0.), sym(norm, 1e-12)
            real_path = s1 and s2 and s3 and not ((ny % 2 == 0 or nx % 2 == 0) and (n1 or n2))
            check(lib.ox_qeplan_create(self.geometry.handle, est, C.c_void_p(wxy.ptr), C.c_void_p(wy.ptr), C.c_void_p(norm.ptr),
                                       _capi.OX_DEVICE, self.dtype, self.max_batch, int(real_path), C.byref(h)))
            self._plans[XY] = (h, real_path)
            return
        wxy = np.ascontiguousarray(N.WXY(XY), dtype=np.float64)
        wy = np.ascontiguousarray(N.WY(XY[1] + XY[1]), dtype=np.float64)
        norm = np.ascontiguousarray(_fmask(np.nan_to_num(np.asarray(AL)), N.fmaskK), dtype=np.float64)
        # symmetric filters that vanish on the Nyquist row/column: Hermitian inputs (transforms of real maps)
        # give real fields, so TT -- and EB, split into real and imaginary parts -- run on half planes
        real_path = (_symmetric(wxy) and _symmetric(wy) and _symmetric(norm, 1e-12)
                     and (ny % 2 or not (wxy[ny // 2].any() or wy[ny // 2].any()))
                     and (nx % 2 or not (wxy[:, nx // 2].any() or wy[:, nx // 2].any())))
        check(lib.ox_qeplan_create(self.geometry.handle, est, ptr(wxy), ptr(wy), ptr(norm), OX_HOST, self.dtype,
                                   self.max_batch, int(real_path), C.byref(h)))
        self._plans[XY] = (h, real_path)

    def path(self, XY):
        """Implementation behind estimator XY for Hermitian inputs: 'c2c' (full-plane chain on cuFFT), 'half' (TT
        on half planes, cuFFT), 'fused' (TT) / 'fused_eb' (EB) on half planes with the hand-written FFT passes
        (include/orphx.h ox_qe_path).  k-maps that are not Hermitian always take the c2c chain."""
        if XY not in self._plans:
            self._make_plan(XY)
        return {0: 'c2c', 1: 'half', 2: 'fused', 3: 'fused_eb'}[lib.ox_qe_path(self._plans[XY][0])]

    def _run(self, XY, X, Y, alreadyFTed, returnFt, accumulate, out=None):
        if XY not in self._plans:
            if XY in ('TT', 'EB'):
                self._make_plan(XY)
            else:
                raise NotImplementedError(f"estimator {XY!r} (TT and EB are on the accelerated path)")
        h, _ = self._plans[XY]
        g = self.geometry.shape
        rdt, cdt = _capi.np_dtype(self.dtype), _capi.np_cdtype(self.dtype)
        idt = np.dtype(cdt if alreadyFTed else rdt)
        npix = g[0] * g[1]
        size = int(np.prod(np.shape(X), dtype=np.int64))
        nb = size // npix
        if nb * npix != size or tuple(np.shape(X)[-2:]) != tuple(g):
            raise ValueError(f"input of shape {np.shape(X)} is not a stack of {g} maps")
        if Y is not None and np.shape(Y) != np.shape(X):
            raise ValueError("the X and Y legs must have one shape")
        if nb > self.max_batch:
            raise ValueError(f"{nb} realisations but max_batch={self.max_batch}")
        dev = isinstance(X, devmap) and X.dtype == idt and (Y is None or (isinstance(Y, devmap) and Y.dtype == idt))
        if dev:     # device-resident inputs (FourierCalc.power2d's k-maps, MapGen's maps) are consumed in HBM
            xp, yp, loc = C.c_void_p(X.ptr), (None if Y is None else C.c_void_p(Y.ptr)), _capi.OX_DEVICE
        else:
            x = np.ascontiguousarray(np.asarray(X), dtype=idt)
            y = None if Y is None else np.ascontiguousarray(np.asarray(Y), dtype=idt)
            xp, yp, loc = ptr(x), ptr(y), OX_HOST
        odt = cdt if returnFt else rdt
        if out is None:
            out, optr, oloc = _enmap.result_map((nb,) + g, odt, self.wcs)
        else:
            if out.dtype != odt or out.size != nb * npix or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous array of the result's dtype and size")
            optr, oloc = ptr(out), OX_HOST
        check(lib.ox_qe_reconstruct(h, xp, yp, loc, nb, int(bool(alreadyFTed)), int(bool(returnFt)),
                                    int(bool(accumulate)), optr, oloc))
        return out

    def kappa_from_map(self, XY, T2DData, E2DData=None, B2DData=None, T2DDataY=None, E2DDataY=None, B2DDataY=None,
                       alreadyFTed=False, returnFt=False, accumulate_meanfield=False):
        """tutorials/tt_verification.ipynb:608,610; lensing.py:973-976."""
        fields = {'T': T2DData, 'E': E2DData, 'B': B2DData}
        fieldsY = {'T': T2DDataY, 'E': E2DDataY, 'B': B2DDataY}
        Xl, Yl = XY
        X = fields[Xl]
        Y = fieldsY[Yl] if fieldsY[Yl] is not None else fields[Yl]
        if X is None or Y is None:
            raise ValueError(f"{XY} needs the {Xl} and {Yl} maps")
        out = self._run(XY, X, None if (Y is X) else Y, alreadyFTed, returnFt, accumulate_meanfield)[0]
        return out if isinstance(out, devmap) else ndmap(out, self.wcs)

    reconstruct = kappa_from_map

    def kappa_from_maps(self, XY, X, Y=None, alreadyFTed=False, returnFt=False, accumulate_meanfield=False, out=None):
        """Batched: X (and Y) are stacks (nbatch, Ny, Nx); returns (nbatch, Ny, Nx).  out: optional result array
        (e.g. pinned host memory, _capi.PinnedArray) to fill instead of a fresh pageable one."""
        return self._run(XY, X, Y, alreadyFTed, returnFt, accumulate_meanfield, out=out)

    def meanfield(self, XY):
        """(sum of kappa_hat(l) on the half plane, count) accumulated on the device."""
        p, nel = self.meanfield_pointer(XY)
        ny, nx = self.geometry.shape
        packed = np.empty(2 * nel + 1, dtype=np.float64)
        check(lib.ox_memcpy_d2h(ptr(packed), C.c_void_p(p), packed.nbytes))
        acc = packed[:2 * nel].view(np.complex128).reshape(ny, nx // 2 + 1).copy()
        return acc, int(round(packed[2 * nel]))

    def meanfield_count(self, XY):
        """Number of realisations in the device mean-field stack (8-byte read)."""
        p, nel = self.meanfield_pointer(XY)
        c = np.empty(1, dtype=np.float64)
        check(lib.ox_memcpy_d2h(ptr(c), C.c_void_p(p + 16 * nel), 8))
        return int(round(c[0]))

    def meanfield_pointer(self, XY):
        """Device address of the packed float64 [stack (complex128 half plane) | count] and the stack's
        number of complex elements."""
        h, _ = self._plans[XY]
        p, n = C.c_void_p(), C.c_longlong()
        check(lib.ox_qe_meanfield(h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def allreduce_meanfield(self, XY, comm):
        """Sum the mean-field stack and its count over the ranks of an mpi.Comm: ONE ncclAllReduce
        (Statistics.add_stack / allreduce, stats.py:1227-1228)."""
        check(lib.ox_qe_meanfield_allreduce(comm.handle, self._plans[XY][0]))

    def reset_meanfield(self, XY):
        check(lib.ox_qe_meanfield_reset(self._plans[XY][0]))

    def __del__(self):
        try:
            for h, _ in self._plans.values():
                lib.ox_qeplan_destroy(h)
            lib.ox_powerplan_destroy(self._fftplan)
        except Exception:
            pass


class SplitLensing(object):
    """Split-based lensing power estimator (lensing.py:959-1003).  qfrag(a, b) is kappa_hat(l) of the
    quadratic estimator with Fourier-space X leg a and Y leg b; cross_estimator combines the estimators of
    all pairs of splits.  Here the 1 + 3n + n(n-1) estimators run as batches on the device (splits uploaded
    once, kappa_hat(l) kept in HBM) and the per-pixel combination is one kernel (ox_split_lensing_combine)."""

    def __init__(self, shape, wcs, qest, XY="TT"):
        from . import maps
        self.fc = maps.FourierCalc(shape, wcs)
        self.qest = qest
        self.est = XY
        if XY != "TT":
            raise NotImplementedError("SplitLensing: only the TT estimator (the reference marks its EE branch 'wrong!')")

    def qpower(self, k1, k2):
        return self.fc.f2power(k1, k2)

    def qfrag(self, a, b):
        return self.qest.kappa_from_map(self.est, T2DData=np.array(a), T2DDataY=np.array(b), alreadyFTed=True, returnFt=True)

    def cross_estimator(self, ksplits):
        q = self.qest
        splits = np.asanyarray(ksplits)
        n = int(splits.shape[0])
        if n < 4:
            raise ValueError("cross_estimator needs at least four splits (its normalisation divides by n-3)")
        g = q.geometry.shape
        if tuple(splits.shape[1:]) != tuple(g):
            raise ValueError(f"ksplits of shape {splits.shape} do not match geometry {g}")
        if self.est not in q._plans:
            q._make_plan(self.est)
        h = q._plans[self.est][0]
        cdt = _capi.np_cdtype(q.dtype)
        npix = g[0] * g[1]
        plane = npix * np.dtype(cdt).itemsize
        maps_h = np.empty((n + 1,) + tuple(g), dtype=cdt)          # m_0 .. m_{n-1}, then the mean s
        maps_h[:n] = splits
        maps_h[n] = np.mean(splits, axis=0)
        M = _capi.DeviceBuffer(maps_h.nbytes).upload(maps_h)
        S = n
        pairs = [(S, S)]
        for i in range(n):
            pairs += [(i, S), (S, i), (i, i)]
        for i in range(n):
            for j in range(i + 1, n):
                pairs += [(i, j), (j, i)]
        nb = q.max_batch
        K = _capi.DeviceBuffer(len(pairs) * plane)
        X, Y = _capi.DeviceBuffer(nb * plane), _capi.DeviceBuffer(nb * plane)
        try:
            for c0 in range(0, len(pairs), nb):
                chunk = pairs[c0:c0 + nb]
                for t, (ia, ib) in enumerate(chunk):
                    check(lib.ox_memcpy_d2d(C.c_void_p(X.ptr + t * plane), C.c_void_p(M.ptr + ia * plane), plane))
                    check(lib.ox_memcpy_d2d(C.c_void_p(Y.ptr + t * plane), C.c_void_p(M.ptr + ib * plane), plane))
                check(lib.ox_qe_reconstruct(h, C.c_void_p(X.ptr), C.c_void_p(Y.ptr), _capi.OX_DEVICE, len(chunk), 1, 1, 0,
                                            C.c_void_p(K.ptr + c0 * plane), _capi.OX_DEVICE))
            out = np.empty(g, dtype=np.float64)
            check(lib.ox_split_lensing_combine(C.c_void_p(K.ptr), _capi.OX_DEVICE, q.dtype, n, npix, C.c_double(self.fc.normfact),
                                               ptr(out), OX_HOST))
        finally:
            for b in (M, K, X, Y):
                b.free()
        return out


# ---------------------------------------------------------------------------------------------
# callers of the hot path: kappa <-> phi (lensing.py:651-665), Taylens (lensing.py:395-440),
# FlatLensingSims (lensing.py:458-521)
def fkappa_to_fphi(fkappa, modlmap):
    """phi(l) = 2 kappa(l) / (L (L+1)), zero for L < 2 (lensing.py:662-665)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        kmap = np.nan_to_num(2. * fkappa / modlmap / (modlmap + 1.))
    kmap[modlmap < 2.] = 0.
    return kmap


class _C2C(object):
    """Batched device c2c transforms of full-plane arrays (ox_fft_c2c) for the callers below."""

    def __init__(self, geometry, nplanes=8):
        self.geometry = geometry
        self.npix = geometry.npix
        self.plan = C.c_void_p()
        check(lib.ox_powerplan_create(geometry.handle, 1, _capi.OX_F64, nplanes, C.byref(self.plan)))

    def __call__(self, a, inverse=False, scale=None):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        out = np.empty_like(a)
        if scale is None:
            scale = 1.0 / self.npix if inverse else 1.0
        check(lib.ox_fft_c2c(self.plan, ptr(a), OX_HOST, a.size // self.npix, 1 if inverse else -1, float(scale), ptr(out), OX_HOST))
        return out

    def __del__(self):
        try:
            lib.ox_powerplan_destroy(self.plan)
        except Exception:
            pass


class _LensPlan(object):
    """Device state of the lensing callers for one geometry (ox_lensplan_*): transforms, the deflection field of the
    current potential and its Taylens split."""

    _cache = {}

    def __init__(self, geometry, wcs, max_planes=8):
        self.geometry = geometry
        ext = _enmap.extent(geometry.shape, wcs, method=geometry.method)
        self.py, self.px = ext[0] / geometry.shape[0], ext[1] / geometry.shape[1]     # enmap.pixshape (lensing.py:420)
        self.max_planes = int(max_planes)
        self.handle = C.c_void_p()
        check(lib.ox_lensplan_create(geometry.handle, C.c_double(self.py), C.c_double(self.px), self.max_planes, C.byref(self.handle)))

    @classmethod
    def get(cls, shape, wcs, max_planes=8):
        g = Geometry.get(shape, wcs)
        key = (id(g), int(max_planes))
        if key not in cls._cache:
            cls._cache[key] = cls(g, wcs, max_planes)
        return cls._cache[key]

    def __del__(self):
        try:
            lib.ox_lensplan_destroy(self.handle)
        except Exception:
            pass


def _f64_in(x):
    """(void* argument, location, keep-alive) of a float64 map for the lensing calls."""
    if isinstance(x, devmap) and x.dtype == np.float64:
        return C.c_void_p(x.ptr), _capi.OX_DEVICE, x
    a = np.ascontiguousarray(np.asarray(x), dtype=np.float64)
    return ptr(a), OX_HOST, a


def kappa_to_phi(kappa, modlmap, return_fphi=False, _fft=None):
    """lensing.py:651-657: phi = Re ifft(2 fft(kappa) / (L (L+1))), on the device (ox_lens_kappa_to_phi).  The
    reference transforms with enmap.fft/ifft(normalize='phys'): forward = raw FFT x (pixsize/Npix)^1/2, backward =
    raw backward FFT / (pixsize Npix)^1/2.  The factors cancel in phi; the returned fphi carries the forward one.
    ``modlmap`` must be the geometry's (as the reference passes kappa.modlmap()); any other array takes the generic
    route through full-plane transforms."""
    wcs = kappa.wcs
    g = Geometry.get(kappa.shape, wcs)
    own = modlmap is None or (np.shape(modlmap) == g.shape and np.array_equal(np.asarray(modlmap), g.modlmap()))
    if own and not return_fphi:
        lp = _LensPlan.get(kappa.shape, wcs)
        a, loc, _keep = _f64_in(kappa)
        out, optr, oloc = _enmap.result_map(g.shape, np.float64, wcs)
        check(lib.ox_lens_kappa_to_phi(lp.handle, a, loc, optr, oloc))
        return out
    f = _fft or _C2C(g)
    pix = _enmap.pixsize(g.shape, wcs, method=g.method)
    fphi = fkappa_to_fphi(f(np.asarray(kappa), scale=float(np.sqrt(pix / g.npix))), np.asarray(modlmap))
    phi = ndmap(f(fphi, inverse=True, scale=float(1.0 / np.sqrt(pix * g.npix))).real, wcs)
    return (phi, ndmap(fphi, wcs)) if return_fphi else phi


def alpha_from_kappa(kappa=None, posmap=None, phi=None, grad=True):
    """Deflection field grad(phi) as (2, Ny, Nx) = (d/dy, d/dx) in radians (lensing.py:443-454, grad=True branch), by
    FFT on the device.  (The reference differentiates with enmap.grad, also a Fourier derivative.)"""
    if not grad:
        raise NotImplementedError("alpha_from_kappa(grad=False) (pixel positions through pixell's sky2pix) is outside the accelerated path")
    if phi is None:
        phi = kappa_to_phi(kappa, None)
    wcs = phi.wcs
    lp = _LensPlan.get(phi.shape, wcs)
    a, loc, _keep = _f64_in(phi)
    check(lib.ox_lens_set_phi(lp.handle, a, loc))
    out, optr, oloc = _enmap.result_map((2,) + lp.geometry.shape, np.float64, wcs)
    check(lib.ox_lens_alpha(lp.handle, optr, oloc))
    return out


def flat_taylens(phi, imap, taylor_order=5, _fft=None):
    """Lens imap by the potential phi with the Taylens algorithm (lensing.py:395-440): nearest-pixel remap plus a
    Taylor series in the sub-pixel deflection.  Everything runs on the device (ox_lens_set_phi / ox_lens_taylens):
    the deflection and every derivative are half-plane inverse transforms, the remap and the series one gather
    kernel per order."""
    wcs = phi.wcs
    g = Geometry.get(phi.shape, wcs)
    if tuple(np.shape(imap)[-2:]) != g.shape:
        raise ValueError(f"imap of shape {np.shape(imap)} does not match phi {g.shape}")
    lp = _LensPlan.get(phi.shape, wcs, max_planes=max(8, int(taylor_order)))
    a, loc, _kp = _f64_in(phi)
    check(lib.ox_lens_set_phi(lp.handle, a, loc))
    return _taylens_current(lp, imap, taylor_order, getattr(imap, "wcs", wcs))


def _taylens_current(lp, imap, taylor_order, wcs):
    nmaps = int(np.prod(np.shape(imap)[:-2], dtype=np.int64)) if np.ndim(imap) > 2 else 1
    a, loc, _ki = _f64_in(imap)
    out, optr, oloc = _enmap.result_map(np.shape(imap), np.float64, wcs)
    check(lib.ox_lens_taylens(lp.handle, a, loc, nmaps, int(taylor_order), optr, oloc))
    return out


def displace_map(imap, phi=None, order=3, _plan=None):
    """Stand-in for pixell.lensing.displace_map(imap, alpha_pix, order) as FlatLensingSims calls it (lensing.py:512):
    the map interpolated at the positions displaced by grad(phi), periodic boundaries, bicubic (Keys a = -1/2)
    convolution on the device (ox_lens_displace).  pixell interpolates with prefiltered splines of order
    ``order``; only order=3 (cubic) has a device kernel here and the two cubic schemes differ at the 1e-3 level of the
    map's small-scale power (not a parity path: pixell is absent from the reference snapshot)."""
    if order != 3:
        raise NotImplementedError("displace_map: the device kernel is bicubic (order=3); use flat_taylens for higher orders")
    wcs = getattr(imap, "wcs", None) if phi is None else phi.wcs
    lp = _plan or _LensPlan.get(np.shape(imap), wcs)
    if phi is not None:
        a, loc, _kp = _f64_in(phi)
        check(lib.ox_lens_set_phi(lp.handle, a, loc))
    nmaps = int(np.prod(np.shape(imap)[:-2], dtype=np.int64)) if np.ndim(imap) > 2 else 1
    a, loc, _ki = _f64_in(imap)
    out, optr, oloc = _enmap.result_map(np.shape(imap), np.float64, getattr(imap, "wcs", wcs))
    check(lib.ox_lens_displace(lp.handle, a, loc, nmaps, optr, oloc))
    return out


def get_central(img, fracy, fracx=None):
    """maps.get_central (maps.py:1322-1336)."""
    if fracy is None and fracx is None:
        return img
    fracx = fracy if fracx is None else fracx
    Ny, Nx = img.shape[-2:]
    cropy, cropx = int(fracy * Ny), int(fracx * Nx)
    if (cropy % 2 == 0 and Ny % 2 == 1) or (cropy % 2 == 1 and Ny % 2 == 0):
        cropy -= 1
    if (cropx % 2 == 0 and Nx % 2 == 1) or (cropx % 2 == 1 and Nx % 2 == 0):
        cropx -= 1
    sy, sx = (Ny - cropy) // 2, (Nx - cropx) // 2
    return img[..., sy:sy + cropy, sx:sx + cropx]


class FlatLensingSims(object):
    """lensing.py:458-521: unlensed GRF -> lensing -> beam -> + noise GRF.  Every step between the seeds and
    ``observed`` runs on the device and the maps handed from step to step stay in HBM (enmap.devmap): MapGen.get_map
    x3, kappa_to_phi, the lensing operation, filter_map (one fused r2c -> beam -> c2r pass) and the final sum.
    The reference remaps with pixell.lensing.displace_map (spline interpolation of order lens_order; third party,
    absent from the snapshot).  lensing="taylens" (default) uses the reference's own FFT-based flat_taylens
    (lensing.py:395-440) with taylor_order = lens_order; lensing="bicubic" uses the bicubic displacement kernel
    (displace_map above).  noise: the MapGens' noise source ("numpy" = the reference's seeds, "philox" = device)."""

    def __init__(self, shape, wcs, theory, beam_arcmin, noise_uk_arcmin, noise_e_uk_arcmin=None, noise_b_uk_arcmin=None,
                 pol=False, fixed_lens_kappa=None, lensing="taylens", noise="numpy"):
        from . import maps, cosmology
        if len(shape) < 3 and pol:
            shape = (3,) + tuple(shape)
        if noise_e_uk_arcmin is None:
            noise_e_uk_arcmin = np.sqrt(2.) * noise_uk_arcmin
        if noise_b_uk_arcmin is None:
            noise_b_uk_arcmin = noise_e_uk_arcmin
        if lensing not in ("taylens", "bicubic"):
            raise ValueError("lensing must be 'taylens' or 'bicubic'")
        self.lensing = lensing
        self.shape, self.wcs = tuple(shape), wcs
        self.geometry = Geometry.get(shape, wcs)
        self.modlmap = ndmap(self.geometry.modlmap(), wcs)
        Ny, Nx = shape[-2:]
        ells = np.arange(0, self.modlmap.max(), 1)
        ps_cmb = cosmology.power_from_theory(ells, theory, lensed=False, pol=pol)
        self.mgen = maps.MapGen(shape, wcs, ps_cmb, noise=noise)
        self._lp = _LensPlan.get(shape, wcs)
        self._fc = maps.FourierCalc(shape, wcs)
        if fixed_lens_kappa is not None:
            self._fixed = True
            self.update_kappa(fixed_lens_kappa)
        else:
            self._fixed = False
            ps_kk = theory.gCl('kk', self.modlmap).reshape((1, 1, Ny, Nx))
            self.kgen = maps.MapGen(shape[-2:], wcs, ps_kk, noise=noise)
            self.ps_kk = ps_kk
        self.kbeam = maps.gauss_beam(self.modlmap, beam_arcmin)
        self._kbeam_dev = _enmap.to_device(np.asarray(self.kbeam), np.float64, wcs) if _enmap.DEVICE_RESIDENT else self.kbeam
        ncomp = 3 if pol else 1
        ps_noise = np.zeros((ncomp, ncomp, Ny, Nx))
        ps_noise[0, 0] = (noise_uk_arcmin * np.pi / 180. / 60.) ** 2.
        if pol:
            ps_noise[1, 1] = (noise_e_uk_arcmin * np.pi / 180. / 60.) ** 2.
            ps_noise[2, 2] = (noise_b_uk_arcmin * np.pi / 180. / 60.) ** 2.
        self.ngen = maps.MapGen(shape, wcs, ps_noise, noise=noise)
        self.ps_noise = ps_noise

    def update_kappa(self, kappa):
        self.kappa = kappa
        k = kappa if isinstance(kappa, devmap) else ndmap(np.asarray(kappa), self.wcs)
        self.phi = kappa_to_phi(k, None)
        a, loc, _keep = _f64_in(self.phi)
        check(lib.ox_lens_set_phi(self._lp.handle, a, loc))     # the deflection and its Taylens split stay in the plan

    def get_unlensed(self, seed=None):
        return self.mgen.get_map(seed=seed)

    def get_kappa(self, seed=None):
        return self.kgen.get_map(seed=seed)

    def get_sim(self, seed_cmb=None, seed_kappa=None, seed_noise=None, lens_order=5, return_intermediate=False,
                skip_lensing=False, cfrac=None):
        from . import maps
        unlensed = self.get_unlensed(seed_cmb)
        if skip_lensing:
            lensed = unlensed
            kappa = ndmap(np.zeros(self.geometry.shape), self.wcs)
        else:
            if not (self._fixed):
                kappa = self.get_kappa(seed_kappa)
                self.update_kappa(kappa)
            else:
                kappa = None
                assert seed_kappa is None
            if self.lensing == "taylens":
                lensed = _taylens_current(self._lp, unlensed, lens_order, self.wcs)
            else:
                lensed = displace_map(unlensed, None, order=3, _plan=self._lp)
        beamed = maps.filter_map(lensed, self._kbeam_dev, self._fc)
        noise_map = self.ngen.get_map(seed=seed_noise)
        observed = beamed + noise_map
        if not isinstance(observed, devmap):
            observed = ndmap(np.asarray(observed), self.wcs)
        if return_intermediate:
            return [get_central(x, cfrac) if x is not None else None for x in [unlensed, kappa, lensed, beamed, noise_map, observed]]
        return get_central(observed, cfrac)
