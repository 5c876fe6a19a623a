"""ORACLE (test infrastructure) -- numpy restatement of the FFT-based callers around the hot path in
/root/reference/orphics/lensing.py: fkappa_to_fphi / kappa_to_phi (:651-665), flat_taylens
(:395-440) and FlatLensingSims.get_sim (:499-521) with flat_taylens in place of
pixell.lensing.displace_map (third-party spline remap, out of the path's scope).
The function bodies are pinned by tests/golden/lensing_refbody.npz (the reference's own flat_taylens / kappa_to_phi
run unmodified over oracle/enmap_np.py); PARITY UNPINNED for the pixell pieces underneath (enmap.fft(normalize='phys'),
laxes, pixshape)."""
from math import comb, factorial

import numpy as np

from . import enmap_np as enmap, maps_np as maps, theory as otheory


def fkappa_to_fphi(fkappa, modlmap):
    with np.errstate(divide="ignore", invalid="ignore"):
        kmap = np.nan_to_num(2. * fkappa / modlmap / (modlmap + 1.))     # lensing.py:663
    kmap[modlmap < 2.] = 0.
    return kmap


def kappa_to_phi(kappa, modlmap):
    f = lambda x: enmap.fft(enmap.ndmap(x, kappa.wcs), normalize="phys")
    invf = lambda x: enmap.ifft(enmap.ndmap(x, kappa.wcs), normalize="phys")
    return enmap.ndmap(np.asarray(invf(fkappa_to_fphi(np.asarray(f(kappa)), modlmap))).real, kappa.wcs)


def flat_taylens(phi, imap, taylor_order=5):
    """lensing.py:395-440."""
    wcs = phi.wcs
    f = lambda x: np.asarray(enmap.fft(enmap.ndmap(x, wcs), normalize="phys"))
    invf = lambda x: np.asarray(enmap.ifft(enmap.ndmap(x, wcs), normalize="phys"))
    kmap = f(phi)
    Ny, Nx = phi.shape
    ly_array, lx_array = np.asarray(enmap.lmap(phi.shape, wcs))
    alphaX = np.real(invf(1j * lx_array * kmap))
    alphaY = np.real(invf(1j * ly_array * kmap))
    iy, ix = np.mgrid[0:Ny, 0:Nx]
    py, px = enmap.extent(phi.shape, wcs) / np.array(phi.shape)
    alphaX0 = np.array(np.round(alphaX / px), dtype="int64")
    alphaY0 = np.array(np.round(alphaY / py), dtype="int64")
    dX, dY = alphaX - alphaX0 * px, alphaY - alphaY0 * py
    lensed = np.asarray(imap)[(iy + alphaY0) % Ny, (ix + alphaX0) % Nx].astype(np.float64)
    kmap = f(imap)
    for n in range(1, taylor_order):
        for k in range(n + 1):
            fac = 1j ** n * comb(n, k) * lx_array ** (n - k) * ly_array ** k / factorial(n)
            lensed = lensed + np.real(invf(fac * kmap))[(iy + alphaY0) % Ny, (ix + alphaX0) % Nx] * dX ** (n - k) * dY ** k
    return enmap.ndmap(lensed, wcs)


def grad_phi(phi):
    """(d phi/dy, d phi/dx) by FFT, (2, Ny, Nx): the deflection flat_taylens forms at lensing.py:414-419 and
    alpha_from_kappa's grad branch (lensing.py:443-449)."""
    wcs = phi.wcs
    f = lambda x: np.asarray(enmap.fft(enmap.ndmap(x, wcs), normalize="phys"))
    invf = lambda x: np.asarray(enmap.ifft(enmap.ndmap(x, wcs), normalize="phys"))
    kmap = f(phi)
    ly_array, lx_array = np.asarray(enmap.lmap(phi.shape, wcs))
    return np.stack([np.real(invf(1j * ly_array * kmap)), np.real(invf(1j * lx_array * kmap))])


def displace_bicubic(imap, phi):
    """NOT a reference restatement: CPU definition of the bicubic stand-in for pixell.lensing.displace_map
    (lensing.py:512) -- Keys cubic convolution (a = -1/2), periodic, at (iy + alphaY/py, ix + alphaX/px)."""
    a = np.asarray(imap, dtype=np.float64)
    Ny, Nx = a.shape[-2:]
    alphaY, alphaX = grad_phi(phi)
    py, px = enmap.extent(phi.shape, phi.wcs) / np.array(phi.shape)
    iy, ix = np.mgrid[0:Ny, 0:Nx]
    fy, fx = iy + alphaY / py, ix + alphaX / px
    by, bx = np.floor(fy), np.floor(fx)

    def weights(t):
        t2, t3 = t * t, t * t * t
        return [-0.5 * t3 + t2 - 0.5 * t, 1.5 * t3 - 2.5 * t2 + 1.0, -1.5 * t3 + 2.0 * t2 + 0.5 * t, 0.5 * t3 - 0.5 * t2]
    wy, wx = weights(fy - by), weights(fx - bx)
    y0, x0 = by.astype(np.int64) - 1, bx.astype(np.int64) - 1
    out = np.zeros(a.shape)
    for jy in range(4):
        r = 0.0
        for jx in range(4):
            r = r + wx[jx] * a[..., (y0 + jy) % Ny, (x0 + jx) % Nx]
        out = out + wy[jy] * r
    return enmap.ndmap(out, phi.wcs)


class FlatLensingSims:
    """lensing.py:458-521 (T-only or IQU), lensing by flat_taylens."""

    def __init__(self, shape, wcs, theory, beam_arcmin, noise_uk_arcmin, pol=False):
        if len(shape) < 3 and pol:
            shape = (3,) + tuple(shape)
        self.shape, self.wcs = shape, wcs
        self.modlmap = np.asarray(enmap.modlmap(shape, wcs))
        Ny, Nx = shape[-2:]
        ells = np.arange(0, self.modlmap.max(), 1)
        self.mgen = maps.MapGen(shape, wcs, otheory.power_from_theory(ells, theory, lensed=False, pol=pol))
        self.kgen = maps.MapGen(shape[-2:], wcs, theory.gCl("kk", self.modlmap).reshape((1, 1, Ny, Nx)))
        self.kbeam = maps.gauss_beam(self.modlmap, beam_arcmin)
        ncomp = 3 if pol else 1
        ps_noise = np.zeros((ncomp, ncomp, Ny, Nx))
        ps_noise[0, 0] = (noise_uk_arcmin * np.pi / 180. / 60.) ** 2.
        if pol:
            ps_noise[1, 1] = ps_noise[2, 2] = (np.sqrt(2.) * noise_uk_arcmin * np.pi / 180. / 60.) ** 2.
        self.ngen = maps.MapGen(shape, wcs, ps_noise)

    def get_sim(self, seed_cmb, seed_kappa, seed_noise, lens_order=5, skip_lensing=False):
        unlensed = self.mgen.get_map(seed=seed_cmb)
        if skip_lensing:
            lensed, kappa = unlensed, None
        else:
            kappa = self.kgen.get_map(seed=seed_kappa)
            phi = kappa_to_phi(kappa, self.modlmap)
            comps = np.asarray(unlensed).reshape((-1,) + tuple(self.shape[-2:]))
            lensed = np.stack([flat_taylens(phi, enmap.ndmap(c, self.wcs), lens_order) for c in comps]).reshape(np.shape(unlensed))
            lensed = enmap.ndmap(lensed, self.wcs)
        beamed = maps.filter_map(lensed, self.kbeam)
        noise = self.ngen.get_map(seed=seed_noise)
        return unlensed, kappa, lensed, beamed, noise, enmap.ndmap(np.asarray(beamed) + np.asarray(noise), self.wcs)


class SplitLensing:
    """lensing.py:959-1003: split-based four-point estimator.  qfrag(a, b) = kappa_hat(l) of the quadratic
    estimator with Fourier-space X leg a and Y leg b (lensing.py:973-976); cross_estimator combines the
    estimators of every pair of splits so that no noise bias from a single split survives."""

    def __init__(self, shape, wcs, qest, XY="TT", fourier_calc=None):
        from . import maps_np
        self.fc = fourier_calc if fourier_calc is not None else maps_np.FourierCalc(shape, wcs)
        self.qest, self.est = qest, XY

    def qpower(self, k1, k2):
        return self.fc.f2power(k1, k2)

    def qfrag(self, a, b):
        assert self.est == "TT"   # the reference's 'EE' branch is marked "wrong!" (lensing.py:977)
        return self.qest.kappa_from_map("TT", T2DData=np.array(a), T2DDataY=np.array(b), alreadyFTed=True, returnFt=True)

    def cross_estimator(self, ksplits):
        splits = np.asanyarray(ksplits)
        n = splits.shape[0]
        fn = float(n)
        s = np.mean(splits, axis=0)
        k = self.qfrag(s, s)
        kiisum, psum, psum2 = 0.0, 0.0, 0.0
        for i in range(n):
            mi = splits[i]
            ki = (self.qfrag(mi, s) + self.qfrag(s, mi)) / 2.0
            kii = self.qfrag(mi, mi)
            kiisum = kiisum + kii
            kic = ki - kii / fn
            psum = psum + self.qpower(kic, kic)
            for j in range(i + 1, n):
                mj = splits[j]
                kij = (self.qfrag(mi, mj) + self.qfrag(mj, mi)) / 2.0
                psum2 = psum2 + self.qpower(kij, kij)
        kc = k - kiisum / fn ** 2
        return (fn ** 4 * self.qpower(kc, kc) - 4.0 * fn ** 2 * psum + 4.0 * psum2) / fn / (fn - 1.0) / (fn - 2.0) / (fn - 3.0)
