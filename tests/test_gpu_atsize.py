"""GPU parity AT THE BASELINE SIZES against the numpy oracle (SURVEY 8d gates):

  configs[1]  T-only 2048^2  sim -> taper -> power2d -> bin2D, 4 seeds in host-noise (numpy seed) mode
  configs[2]  IQU 2048^2 with the TEB rotation, all 6 binned spectra, 2 seeds
  configs[3]  TT quadratic estimator at 4096^2: A_L, kappa_hat(l), kappa map
  configs[4]  EB quadratic estimator at 4096^2 (fp64 and fp32) and a TT + EB spot check at 8192^2

The product builds its OWN set-up (covsqrt from the spectra, filters and A_L) -- nothing is injected from
the oracle -- and everything goes through the C ABI.  Tolerances: 1e-10 relative in fp64 (maps and Fourier
arrays relative to max|x|, bandpowers per spectrum), bit-exact slot indices and counts; the fp32 estimator is
held to the bound stated in include/orphx.h (ox_qeplan_create).  The oracle's FFTs run on all host cores
(scipy.fft workers) so that the whole file costs a few minutes of CPU.
"""
import os

import numpy as np
import pytest
import scipy.fft

from conftest import relerr
from oracle import enmap_np as oenmap, maps_np as omaps, stats_np as ostats, qe_np

pytestmark = pytest.mark.gpu
TOL64 = 1e-10
EDGES = np.arange(100, 3000, 40.0)          # tutorials/demo-grf.ipynb:159
KEDGES = np.linspace(20, 3500, 20)          # tutorials/tt_verification.ipynb:600
PAIRS = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]


@pytest.fixture(autouse=True)
def _oracle_on_all_cores():
    with scipy.fft.set_workers(max(1, min(32, os.cpu_count() or 1))):
        yield


def _sim_setup(pol, theory):
    from orphics_b200 import maps, stats, cosmology
    npix, res = 2048, 0.5
    shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res, pol=pol)
    so, wo = omaps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res, pol=pol)
    assert tuple(shape) == tuple(so)
    modl = np.asarray(oenmap.modlmap(so, wo))
    ells = np.arange(0, modl.max() + 1, 1.0)
    from oracle import theory as otheory
    ps_o = otheory.power_from_theory(ells, theory, lensed=True, pol=pol)
    ps = cosmology.power_from_theory(ells, cosmology.default_theory(), lensed=True, pol=pol)
    assert np.array_equal(ps, ps_o)
    return shape, wcs, so, wo, modl, ps


@pytest.mark.parametrize("pol,nseed", [(False, 4), (True, 2)])
def test_pipeline_2048_matches_oracle_on_numpy_seeds(pol, nseed, theory):
    """configs[1] / configs[2] at full size: MapGen.get_map (maps.py:1576-1587) -> x taper ->
    FourierCalc.power2d (maps.py:1639-1677) -> bin2D.bin (stats.py:790) on identical numpy seeds, through the
    hand-written fused kernels; stored maps and bandpowers vs the oracle, slot indices and counts bit-exact."""
    from orphics_b200 import maps, stats
    shape, wcs, so, wo, modl, ps = _sim_setup(pol, theory)
    og, ofc, ob = omaps.MapGen(so, wo, ps), omaps.FourierCalc(so, wo), ostats.bin2D(modl, EDGES)
    otaper = np.asarray(omaps.get_taper(so, wo)[0])
    mg = maps.MapGen(shape, wcs, ps, max_batch=nseed)                  # the product's own set-up
    assert relerr(mg.covsqrt, og.covsqrt) < 1e-12
    fc = maps.FourierCalc(shape, wcs, max_batch=nseed)
    b = stats.bin2D(fc.geometry.modlmap(), EDGES, geometry=fc.geometry)
    assert np.array_equal(fc.geometry.modlmap(), modl)
    assert np.array_equal(b.digitized, ob.digitized)
    assert np.array_equal(b.slot_counts, np.bincount(ob.digitized, minlength=len(EDGES) + 1))
    taper = np.asarray(maps.get_taper(shape, wcs)[0])
    assert np.array_equal(taper, otaper)
    pipe = maps.SimPipeline(mg, fc, b, window=taper)
    assert pipe.path == "fused"
    seeds = [1000 + i for i in range(nseed)]                           # sim i uses seed 1000+i
    bp = pipe.run(seeds, noise="numpy", keep_maps=True)
    stored = pipe.last_maps(nseed)
    pairs = PAIRS if pol else [(0, 0)]
    assert bp.shape == (nseed, len(pairs), len(EDGES) - 1)
    for i, seed in enumerate(seeds):
        mo = og.get_map(seed=seed)
        assert relerr(stored[i] if pol else stored[i, 0], mo) < TOL64
        p2o = ofc.power2d(oenmap.ndmap(np.asarray(mo) * otaper, wo))[0]
        want = np.array([ob.bin(p2o[a, c] if pol else p2o)[1] for a, c in pairs])
        auto = {0: 0, 1: 3, 2: 5}
        for s, (a, c) in enumerate(pairs):
            scale = np.sqrt(np.abs(want[auto[a] if pol else 0] * want[auto[c] if pol else 0]))
            err = np.max(np.abs(bp[i, s] - want[s]) / scale)
            assert err < TOL64, (seed, a, c, err)
    if not pol:
        # the per-call reference-signature chain on the same seed gives the same numbers
        m = mg.get_map(seed=seeds[0])
        c, p1 = b.bin(fc.power2d(m * taper)[0])
        assert np.max(np.abs(p1 - bp[0, 0]) / np.abs(bp[0, 0])) < TOL64


def _qe_setup(npix, theory, pol, dtype=np.float64, max_batch=1):
    from orphics_b200 import maps, lensing, cosmology
    res = 0.5
    shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
    so, wo = omaps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
    modl = np.asarray(oenmap.modlmap(so, wo))
    beam = omaps.gauss_beam(modl, 1.5)                                  # beam 1.5', noise 1 uK'
    n2d = np.zeros(so) + (1.0 * np.pi / 180 / 60) ** 2
    tm = np.asarray(omaps.mask_kspace(so, wo, lmin=300, lmax=2000))
    km = np.asarray(omaps.mask_kspace(so, wo, lmin=20, lmax=3500))
    kw = dict(noise2d=n2d, beam2d=beam, kmask=tm, noise2d_P=2 * n2d, kmask_P=tm, kmask_K=km, pol=pol, grad_cut=None,
              unlensed_equals_lensed=True, bigell=9000)               # tutorials/tt_verification.ipynb:81
    qo = qe_np.qest(so, wo, theory, **kw)
    q = lensing.qest(shape, wcs, cosmology.default_theory(), dtype=dtype, max_batch=max_batch, **kw)
    return shape, wcs, so, wo, modl, q, qo


def _observed_like(shape, rng, amp):
    """Band-limited Gaussian field with the amplitude of an observed CMB map (the estimator is quadratic in it)."""
    return rng.standard_normal(shape) * amp


def test_config3_tt_qe_4096_matches_oracle(theory):
    """configs[3] at full size: qest("TT") on a 4096^2 map: A_L, kappa_hat(l) (returnFt) and the kappa map."""
    shape, wcs, so, wo, modl, q, qo = _qe_setup(4096, theory, pol=False)
    assert q.path("TT") == "fused"
    assert relerr(q.N.AL["TT"], qo.N.AL["TT"]) < 1e-9               # set-up: 15 transforms of filters spanning 10 decades
    qo.N.AL["TT"] = np.asarray(q.N.AL["TT"])
    rng = np.random.RandomState(4096)
    T = _observed_like(shape, rng, 60.0)
    kfo = qo.kappa_from_map("TT", T, returnFt=True)
    kf = q.kappa_from_map("TT", T, returnFt=True)
    assert kf.shape == tuple(shape) and kf.dtype == np.complex128
    assert relerr(kf, kfo) < TOL64
    ko = np.asarray(oenmap.raw_ifft(kfo, normalize=True).real)
    k = q.kappa_from_map("TT", T)
    assert relerr(k, ko) < TOL64
    # binned kappa auto-spectrum of the reconstruction (tutorials/tt_verification.ipynb:612-616)
    from orphics_b200 import maps, stats
    fc = maps.FourierCalc(shape, wcs)
    b = stats.bin2D(fc.geometry.modlmap(), KEDGES, geometry=fc.geometry)
    ob = ostats.bin2D(modl, KEDGES)
    assert np.array_equal(b.digitized, ob.digitized)
    p1 = b.bin(fc.power2d(k)[0])[1]
    p1o = ob.bin(omaps.FourierCalc(so, wo).power2d(oenmap.ndmap(ko, wo))[0])[1]
    assert np.max(np.abs(p1 - p1o) / np.abs(p1o)) < TOL64


def test_config4_eb_qe_4096_fp64_and_fp32(theory):
    """configs[4] at 4096^2: qest("EB") from real E/B maps in fp64 (1e-10) and in the float32 mode (bound of
    include/orphx.h), same inputs."""
    from orphics_b200 import lensing, cosmology
    shape, wcs, so, wo, modl, q, qo = _qe_setup(4096, theory, pol=True)
    assert q.path("EB") == "fused_eb"
    assert relerr(q.N.AL["EB"], qo.N.AL["EB"]) < 1e-9
    qo.N.AL["EB"] = np.asarray(q.N.AL["EB"])
    rng = np.random.RandomState(8)
    E, B = _observed_like(shape, rng, 4.0), _observed_like(shape, rng, 1.0)
    kfo = qo.kappa_from_map("EB", None, E, B, returnFt=True)
    kf = q.kappa_from_map("EB", None, E, B, returnFt=True)
    assert relerr(kf, kfo) < TOL64
    ko = np.asarray(oenmap.raw_ifft(kfo, normalize=True).real)
    assert relerr(q.kappa_from_map("EB", None, E, B), ko) < TOL64
    del q
    q32 = lensing.qest(shape, wcs, cosmology.default_theory(), noise2d=qo.N.noise["TT"], noise2d_P=qo.N.noise["EE"],
                       beam2d=qo.N.beam, kmask=qo.N.fmask["TT"], kmask_P=qo.N.fmask["EE"], kmask_K=qo.N.fmaskK,
                       unlensed_equals_lensed=True, pol=True, dtype=np.float32)
    E32, B32 = E.astype(np.float32), B.astype(np.float32)
    want = qo.kappa_from_map("EB", None, E32.astype(np.float64), B32.astype(np.float64))
    got = q32.kappa_from_map("EB", None, E32, B32)
    assert got.dtype == np.float32
    assert relerr(got, want) < lensing.QE_FP32_BOUND, relerr(got, want)


def test_config4_qe_8192_spot_check(theory):
    """configs[4] spot check at 8192^2 (1 GiB per complex plane): TT and EB kappa_hat(l) on the hand-written FFT
    passes vs the oracle, fp64."""
    shape, wcs, so, wo, modl, q, qo = _qe_setup(8192, theory, pol=True)
    assert q.path("TT") == "fused" and q.path("EB") == "fused_eb"
    for XY in ("TT", "EB"):
        assert relerr(q.N.AL[XY], qo.N.AL[XY]) < 1e-9
        qo.N.AL[XY] = np.asarray(q.N.AL[XY])
    rng = np.random.RandomState(81)
    T = _observed_like(shape, rng, 60.0)
    assert relerr(q.kappa_from_map("TT", T, returnFt=True), qo.kappa_from_map("TT", T, returnFt=True)) < TOL64
    E, B = T * 0.07, _observed_like(shape, rng, 1.0)
    assert relerr(q.kappa_from_map("EB", None, E, B, returnFt=True), qo.kappa_from_map("EB", None, E, B, returnFt=True)) < TOL64


def test_tma_row_pass_stress_at_2048(monkeypatch):
    """256 maps of 2048^2 (four launches of 64: 221 tiles per CTA and launch, the benchmark's shape) through the persistent TMA row
    pass and through the one-tile kernel on the same Philox seeds: every bandpower of every map agrees to rounding (a single
    misread tile -- what a phase mix-up of the slot barriers produces -- moves a bandpower by ~1e-3)."""
    from orphics_b200 import maps, stats, cosmology
    shape, wcs = maps.rect_geometry(width_arcmin=2048 * 0.5, px_res_arcmin=0.5)
    g = maps.Geometry.get(shape, wcs)
    ps = cosmology.power_from_theory(np.arange(0, g.modlmap().max() + 1, 1.0), cosmology.default_theory(), lensed=True, pol=False)
    taper = np.asarray(maps.get_taper(shape, wcs)[0])
    out = {}
    for kb in ("legacy", "tma"):
        monkeypatch.setenv("ORPHX_KB", kb)
        mg = maps.MapGen(shape, wcs, ps, noise="philox_hermitian", max_batch=64)
        fc = maps.FourierCalc(shape, wcs, max_batch=64)
        b = stats.bin2D(g.modlmap(), EDGES, geometry=g)
        pipe = maps.SimPipeline(mg, fc, b, window=taper)
        out[kb] = pipe.run(range(5000, 5256), keep_maps=True)
        del pipe, mg, fc, b
    rel = np.abs(out["tma"] - out["legacy"]) / np.abs(out["legacy"])
    assert np.nanmax(rel) < 1e-12
