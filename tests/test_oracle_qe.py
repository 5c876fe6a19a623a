"""CPU: physical validation of the (parity-unpinned) quadratic-estimator oracle -- the
acceptance test of tutorials/tt_verification.ipynb:597-617: <kappa_hat x kappa>/<kappa x kappa> -> 1."""
import numpy as np

from oracle import enmap_np as enmap, maps_np as maps, stats_np as stats, qe_np


class FirstOrderTheory:
    """'lensed' = unlensed: a first-order lensed sim has no lensed-power change at leading order."""

    def __init__(self, th):
        self.th = th

    def lCl(self, k, l):
        return self.th.uCl(k, l)

    def uCl(self, k, l):
        return self.th.uCl(k, l)


def lensed_first_order(T, kap, modl, LY, LX):
    kk = np.fft.fft2(kap)
    with np.errstate(all="ignore"):
        phik = np.nan_to_num(2 * kk / (modl * (modl + 1)))      # lensing.py:662-665
    phik[modl < 2] = 0
    kT = np.fft.fft2(T)
    d = lambda L, k: np.fft.ifft2(1j * L * k).real
    return T + d(LX, phik) * d(LX, kT) + d(LY, phik) * d(LY, kT)


def test_tt_estimator_has_unit_response(theory):
    npix, res = 128, 2.0
    shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
    modl = np.asarray(enmap.modlmap(shape, wcs))
    ly, lx = enmap.laxes(shape, wcs)
    LY, LX = np.meshgrid(ly, lx, indexing="ij")
    ells = np.arange(0, modl.max() + 1, 1.)
    mgT = maps.MapGen(shape, wcs, theory.uCl("TT", ells)[None, None])
    mgK = maps.MapGen(shape, wcs, theory.gCl("kk", ells)[None, None])
    nlev = (1.0 * np.pi / 180 / 60) ** 2
    mgN = maps.MapGen(shape, wcs, (np.zeros(ells.size) + nlev)[None, None])
    beam = maps.gauss_beam(modl, 1.5)
    tmask = maps.mask_kspace(shape, wcs, lmin=300, lmax=3000)
    kmask = maps.mask_kspace(shape, wcs, lmin=100, lmax=3000)
    q = qe_np.qest(shape, wcs, FirstOrderTheory(theory), noise2d=np.zeros(shape) + nlev, beam2d=beam, kmask=tmask, kmask_K=kmask)
    assert np.all(q.N.Nlkk["TT"] >= 0) and q.N.Nlkk["TT"][modl < 2].max() == 0
    fc = maps.FourierCalc(shape, wcs)
    b = stats.bin2D(modl, np.linspace(100, 3000, 6))
    rs = []
    for i in range(40):
        T = np.asarray(mgT.get_map(seed=100 + i))
        kap = np.asarray(mgK.get_map(seed=5000 + i))
        obs = np.fft.ifft2(np.fft.fft2(lensed_first_order(T, kap, modl, LY, LX)) * beam).real + np.asarray(mgN.get_map(seed=9000 + i))
        rec = q.kappa_from_map("TT", obs)
        pc = fc.power2d(enmap.ndmap(rec, wcs), enmap.ndmap(kap, wcs))[0]
        pi = fc.power2d(enmap.ndmap(kap, wcs))[0]
        rs.append(b.bin(pc)[1] / b.bin(pi)[1])
    rs = np.array(rs)
    mean, err = rs.mean(0), rs.std(0) / np.sqrt(len(rs))
    assert np.all(np.abs(mean - 1) < 4 * err + 0.03), (mean, err)
    assert abs(np.average(mean, weights=1 / err ** 2) - 1) < 0.05


def test_estimator_is_linear_and_returnft_consistent(theory):
    shape, wcs = maps.rect_geometry(width_arcmin=64 * 3.0, px_res_arcmin=3.0)
    modl = np.asarray(enmap.modlmap(shape, wcs))
    tmask = maps.mask_kspace(shape, wcs, lmin=200, lmax=2500)
    q = qe_np.qest(shape, wcs, theory, noise2d=np.zeros(shape) + 1e-8, beam2d=maps.gauss_beam(modl, 2.0), kmask=tmask,
                   kmask_P=tmask, kmask_K=tmask, pol=True, unlensed_equals_lensed=True)
    rng = np.random.RandomState(1)
    T, E, B = rng.standard_normal((3,) + shape)
    k = q.kappa_from_map("TT", T)
    np.testing.assert_allclose(q.kappa_from_map("TT", 2 * T), 4 * k, rtol=1e-12, atol=1e-12 * np.abs(k).max())  # quadratic
    kf = q.kappa_from_map("TT", np.fft.fft2(T), alreadyFTed=True, returnFt=True)
    np.testing.assert_allclose(np.fft.ifft2(kf).real, k, atol=1e-12 * np.abs(k).max())
    keb = q.kappa_from_map("EB", T, E, B)
    kbe = q.kappa_from_map("EB", np.fft.fft2(T), np.fft.fft2(E), np.fft.fft2(B), alreadyFTed=True)
    np.testing.assert_allclose(keb, kbe, atol=1e-12 * np.abs(keb).max())
    assert np.abs(keb).max() > 0
