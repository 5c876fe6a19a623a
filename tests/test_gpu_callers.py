"""GPU parity of the split / cross-spectrum callers through the C-ABI: maps.split_calc,
maps.noise_from_splits and lensing.SplitLensing.cross_estimator against the golden vectors produced by the
reference's own function bodies (tests/golden/make_golden_callers.py) and against the numpy oracle."""
import os

import numpy as np
import pytest

from conftest import ROOT, relerr
from oracle import enmap_np as oenmap, maps_np as omaps, qe_np, lensing_np

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "split_callers.npz"))


def geom():
    from orphics_b200 import maps
    shape, wcs = maps.rect_geometry(width_arcmin=48 * 2.0, px_res_arcmin=2.0, height_arcmin=32 * 2.0)
    assert tuple(shape) == (32, 48)
    return shape, wcs


@pytest.mark.parametrize("alt", [True, False])
def test_split_calc_matches_reference_golden(alt):
    from orphics_b200 import maps
    shape, wcs = geom()
    fc = maps.FourierCalc(shape, wcs)
    a, b = maps.ndmap(G["sc_isplits"], wcs), maps.ndmap(G["sc_jsplits"], wcs)
    t, c, n = maps.split_calc(a, b, a.mean(0), b.mean(0), fourier_calc=fc, alt=alt)
    assert t.shape == tuple(shape) and t.dtype == np.float64
    for got, key in ((t, "total"), (c, "cross"), (n, "noise")):
        assert relerr(got, G[f"sc_{key}_alt{int(alt)}"]) < 1e-10
    if alt:
        with pytest.raises(AssertionError):
            maps.split_calc(a, b[:3], a.mean(0), b.mean(0), fourier_calc=fc, alt=True)     # maps.py:2314


@pytest.mark.parametrize("ncomp", [1, 3])
@pytest.mark.parametrize("do_cross", [True, False])
def test_noise_from_splits_matches_reference_golden(ncomp, do_cross):
    from orphics_b200 import maps
    shape, wcs = geom()
    sh = (ncomp,) + tuple(shape) if ncomp > 1 else tuple(shape)
    splits = maps.ndmap(G[f"nfs_splits_c{ncomp}"], wcs)
    for dtype, tol in ((np.float64, 2e-6), (np.float32, 1e-5)):    # the reference casts the splits to float32
        fc = maps.FourierCalc(sh, wcs, dtype=dtype)
        noise, cteb = maps.noise_from_splits(splits, fourier_calc=fc, do_cross=do_cross)
        want = G[f"nfs_noise_c{ncomp}_x{int(do_cross)}"]
        assert np.shape(noise) == want.shape
        assert relerr(noise, want) < tol
        if do_cross:
            assert relerr(cteb, G[f"nfs_crossteb_c{ncomp}"]) < tol
        else:
            assert cteb is None
    # fp64 plan vs the oracle fed the same float32-rounded maps in double precision: 1e-10
    fc = maps.FourierCalc(sh, wcs)
    noise, cteb = maps.noise_from_splits(splits, fourier_calc=fc, do_cross=do_cross)
    so, wo = omaps.rect_geometry(width_arcmin=48 * 2.0, px_res_arcmin=2.0, height_arcmin=32 * 2.0)
    ofc = omaps.FourierCalc(sh, wo)
    s64 = np.asarray(splits).astype(np.float32).astype(np.float64).reshape((4, ncomp) + tuple(shape))
    ks = [np.fft.fft2(s) for s in s64]
    auto = sum(ofc.power2d(kmap=k)[0] for k in ks) / 4
    cross = sum(ofc.power2d(kmap=ks[i], kmap2=ks[j])[0] for i in range(4) for j in range(i + 1, 4)) / 6
    assert relerr(np.asarray(noise).reshape(np.shape(auto)), (auto - cross) / 4) < 1e-10


def test_noise_from_splits_six_components():
    """ncomp > 3 (two arrays' I,Q,U) with do_cross=False, docstring of maps.py:2345."""
    from orphics_b200 import maps
    shape, wcs = geom()
    rng = np.random.RandomState(3)
    splits = rng.standard_normal((3, 6) + tuple(shape))
    fc = maps.FourierCalc(tuple(shape), wcs)
    noise, cteb = maps.noise_from_splits(maps.ndmap(splits, wcs), fourier_calc=fc, do_cross=False)
    assert noise.shape == (6, 6) + tuple(shape) and cteb is None
    ks = np.fft.fft2(splits.astype(np.float32).astype(np.float64))
    nf = fc.normfact
    p = lambda x, y: np.real(np.conj(x) * y) * nf
    a, b = 1, 4
    auto = sum(p(ks[s, a], ks[s, b]) for s in range(3)) / 3
    cross = sum(p(ks[i, a], ks[j, b]) for i in range(3) for j in range(i + 1, 3)) / 3
    assert relerr(noise[a, b], (auto - cross) / 3) < 1e-10
    assert np.array_equal(noise[a, b], noise[b, a])


def test_split_lensing_cross_estimator_matches_reference_golden(theory):
    from orphics_b200 import maps, lensing, cosmology
    shape, wcs = maps.rect_geometry(width_arcmin=64 * 4.0, px_res_arcmin=4.0)
    modl = maps.Geometry.get(shape, wcs).modlmap()
    kw = dict(noise2d=np.zeros(shape) + (10.0 * np.pi / 180 / 60) ** 2, beam2d=maps.gauss_beam(modl, 5.0),
              kmask=maps.mask_kspace(shape, wcs, lmin=200, lmax=2000), kmask_K=maps.mask_kspace(shape, wcs, lmin=50, lmax=2500),
              unlensed_equals_lensed=True)
    q = lensing.qest(shape, wcs, cosmology.default_theory(), max_batch=8, **kw)
    assert relerr(q.N.AL["TT"], G["sl_AL"]) < 1e-9
    sl = lensing.SplitLensing(shape, wcs, q)
    got = sl.cross_estimator(G["sl_ksplits"])
    assert got.shape == tuple(shape)
    # compare the chain with identical normalisation input: the oracle's estimator with our A_L
    so, wo = omaps.rect_geometry(width_arcmin=64 * 4.0, px_res_arcmin=4.0)
    qo = qe_np.qest(so, wo, theory, **{k: np.asarray(v) if hasattr(v, "shape") else v for k, v in kw.items()})
    qo.N.AL["TT"] = np.asarray(q.N.AL["TT"])
    want = lensing_np.SplitLensing(so, wo, qo).cross_estimator(G["sl_ksplits"])
    assert relerr(got, want) < 1e-10
    assert relerr(got, G["sl_cross_estimator"]) < 1e-7        # golden used the oracle's own A_L (set-up agrees to 1e-9)
    # qfrag / qpower keep the reference's per-call signatures
    k01 = sl.qfrag(G["sl_ksplits"][0], G["sl_ksplits"][1])
    assert relerr(k01, qo.kappa_from_map("TT", G["sl_ksplits"][0], T2DDataY=G["sl_ksplits"][1], alreadyFTed=True, returnFt=True)) < 1e-10
    assert relerr(sl.qpower(k01, k01), omaps.FourierCalc(so, wo).f2power(k01, k01)) < 1e-12
    with pytest.raises(ValueError):
        sl.cross_estimator(G["sl_ksplits"][:3])


def test_split_lensing_fused_path_512(theory):
    """Same estimator on a power-of-two map (hand-written FFT passes, separate X/Y legs, device-resident
    batches) against the per-call composition of the reference's algorithm over our own qfrag."""
    from orphics_b200 import maps, lensing, cosmology
    shape, wcs = maps.rect_geometry(width_arcmin=512 * 2.0, px_res_arcmin=2.0)
    modl = maps.Geometry.get(shape, wcs).modlmap()
    q = lensing.qest(shape, wcs, cosmology.default_theory(), noise2d=np.zeros(shape) + (5.0 * np.pi / 180 / 60) ** 2,
                     beam2d=maps.gauss_beam(modl, 1.5), kmask=maps.mask_kspace(shape, wcs, lmin=300, lmax=2000),
                     kmask_K=maps.mask_kspace(shape, wcs, lmin=20, lmax=3500), unlensed_equals_lensed=True, max_batch=5)
    assert q.path("TT") == "fused"
    rng = np.random.RandomState(9)
    sig = rng.standard_normal(shape) * 50
    ks = np.stack([np.fft.fft2(sig + 10 * rng.standard_normal(shape)) for _ in range(4)])
    sl = lensing.SplitLensing(shape, wcs, q)
    got = sl.cross_estimator(ks)
    want = lensing_np.SplitLensing.cross_estimator(sl, ks)     # the restated loop, driving sl.qfrag / sl.qpower
    assert relerr(got, want) < 1e-10


@pytest.mark.parametrize("tag", ["2d", "1d", "pair"])
def test_ilc_matches_reference_functions(tag):
    """maps.silc / cilc / silc_noise / cilc_noise on the device vs goldens made by the reference's own functions:
    1e-10 on regular pixels, bit-exact nan_to_num / division-by-zero results on the degenerate ones."""
    from orphics_b200 import maps
    from test_oracle_callers import _ilc_check
    _ilc_check(1e-10, tag, lambda name, *a: getattr(maps, name)(*a))
    with pytest.raises(ValueError):
        maps.silc(np.zeros((2, 4, 4), dtype=complex), np.zeros((3, 3, 4, 4)))
