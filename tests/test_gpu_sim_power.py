"""GPU parity: MapGen / FourierCalc / fused pipeline through the C-ABI vs the numpy oracle
on identical inputs.  Tolerance: 1e-10 relative (fp64), 1e-5 (fp32) -- north_star."""
import numpy as np
import pytest

from conftest import relerr
from oracle import enmap_np as oenmap, maps_np as omaps, stats_np as ostats, theory as otheory, philox_np

pytestmark = pytest.mark.gpu
TOL64, TOL32 = 1e-10, 1e-5
EDGES = np.arange(100, 3000, 40.0)


def setup(npix, res, pol, theory):
    from orphics_b200 import maps
    w = npix * res
    shape, wcs = maps.rect_geometry(width_arcmin=w, px_res_arcmin=res, pol=pol)
    so, wo = omaps.rect_geometry(width_arcmin=w, px_res_arcmin=res, pol=pol)
    modl = np.asarray(oenmap.modlmap(so, wo))
    ells = np.arange(0, modl.max() + 1, 1.0)
    ps = otheory.power_from_theory(ells, theory, lensed=True, pol=pol)
    return shape, wcs, so, wo, modl, ps


def bp_close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))     # empty annuli are 0/0 in both
    scale = np.nanmax(np.abs(b), axis=-1, keepdims=True)
    assert np.nanmax(np.abs(a - b) / scale) < tol, np.nanmax(np.abs(a - b) / scale)


def test_covsqrt_setup_matches_oracle(theory):
    """MapGen.__init__ (maps.py:1573, enmap.spec2flat): the 1-D treatment (mode-weighted smoothing by FFT convolution,
    x Npix/area, matrix power) is host set-up and bit-identical to the oracle's (same scipy.fft convolution; two
    different FFT libraries would agree only to ~1e-9 of the peak, C_l l^2 spans ten decades); the per-pixel
    interpolation runs on the device.  So the product's OWN set-up -> map chain is held to the path's 1e-10."""
    from orphics_b200 import maps, enmap
    for pol in (False, True):
        for npix, res in ((128, 2.0), (512, 2.0)):
            shape, wcs, so, wo, modl, ps = setup(npix, res, pol, theory)
            mg = maps.MapGen(shape, wcs, ps)
            og = omaps.MapGen(so, wo, ps)
            assert relerr(mg.covsqrt, og.covsqrt) < 1e-12
            cov1 = oenmap.spec2flat_1d(so, wo, ps, 0.5)
            dev = enmap.Geometry.get(shape, wcs).interp_spec(cov1)
            assert relerr(dev, og.covsqrt) < 1e-13
        # end to end from cov: maps and bandpowers from the product's own covsqrt vs the oracle's
        for kw in (dict(), dict(scalar=True), dict(harm=True)):
            assert relerr(mg.get_map(seed=5, **kw), og.get_map(seed=5, **kw)) < TOL64


@pytest.mark.parametrize("npix,res", [(512, 2.0), (96, 3.0)])
def test_config1_T_map_seed_parity(npix, res, theory):
    """BASELINE config 1: single T map from the CAMB lensed Cls, same numpy seed as the
    reference: MapGen -> FourierCalc.power2d -> bin2D."""
    from orphics_b200 import maps, stats
    shape, wcs, so, wo, modl, ps = setup(npix, res, False, theory)
    og = omaps.MapGen(so, wo, ps)
    mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt))   # identical set-up input
    fc, ofc = maps.FourierCalc(shape, wcs), omaps.FourierCalc(so, wo)
    for seed in (1000, 7):
        m = mg.get_map(seed=seed)
        mo = og.get_map(seed=seed)
        assert m.shape == mo.shape == shape and m.dtype == np.float64
        assert relerr(m, mo) < TOL64
        p2d, k1, k2 = fc.power2d(m)
        p2o, ko, _ = ofc.power2d(mo)
        assert k1.dtype == np.complex128 and p2d.shape == shape
        assert relerr(k1, ko) < TOL64
        assert relerr(p2d, p2o) < TOL64
        b = stats.bin2D(fc.geometry.modlmap(), EDGES, geometry=fc.geometry)
        ob = ostats.bin2D(modl, EDGES)
        c, p1 = b.bin(p2d)
        co, p1o = ob.bin(p2o)
        bp_close(p1, p1o, TOL64)
        # fused (no p2d) path and the one-liner
        bp = fc.binned_power_batch(b, m)
        bp_close(bp[0, 0], p1o, TOL64)
        taper, w2 = maps.get_taper(shape, wcs)
        c2, pw = maps.binned_power(m, bin_edges=EDGES, fc=fc, binner=b, mask=np.asarray(taper))
        co2, pwo = omaps.binned_power(mo, EDGES, ob, ofc, mask=np.asarray(omaps.get_taper(so, wo)[0]))
        bp_close(pw, pwo, TOL64)
    # harm=True returns covsqrt*rand on the full plane
    kh = mg.get_map(seed=3, harm=True)
    assert relerr(kh, og.get_map(seed=3, harm=True)) < TOL64


def test_fouriercalc_api(theory):
    from orphics_b200 import maps
    shape, wcs, so, wo, modl, ps = setup(64, 4.0, False, theory)
    rng = np.random.RandomState(5)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    fc, ofc = maps.FourierCalc(shape, wcs), omaps.FourierCalc(so, wo)
    ao, bo = oenmap.ndmap(a, wo), oenmap.ndmap(b, wo)
    assert relerr(fc.fft(a), ofc.fft(ao)) < TOL64
    assert relerr(fc.iqu2teb(a, normalize=True), ofc.iqu2teb(ao, normalize=True)) < TOL64
    k = np.asarray(ofc.fft(ao))
    assert relerr(fc.ifft(k), ofc.ifft(k)) < TOL64
    kb = np.asarray(ofc.fft(bo))
    assert relerr(fc.f2power(k, kb), ofc.f2power(k, kb)) < TOL64
    assert relerr(fc.f2power(k, kb, pixel_units=True), ofc.f2power(k, kb, pixel_units=True)) < TOL64
    p, k1 = fc.f1power(a, kb)
    po, k1o = ofc.f1power(ao, kb)
    assert relerr(p, po) < TOL64 and relerr(k1, k1o) < TOL64
    p2d, k1, k2 = fc.power2d(a, b)
    p2o, k1o, k2o = ofc.power2d(ao, bo)
    assert relerr(p2d, p2o) < TOL64 and relerr(k2, k2o) < TOL64
    p2d, _, _ = fc.power2d(kmap=k, kmap2=kb)
    assert relerr(p2d, ofc.power2d(kmap=oenmap.ndmap(k, wo), kmap2=oenmap.ndmap(kb, wo))[0]) < TOL64
    f = maps.gauss_beam(modl, 5.0)
    assert relerr(maps.filter_map(maps.ndmap(a, wcs), f, fc), omaps.filter_map(ao, f)) < TOL64


@pytest.mark.parametrize("iau", [False, True])
def test_config3_IQU_seed_parity_six_spectra(iau, theory):
    """BASELINE config 3 (at an oracle-affordable size): IQU sims with the EB->QU rotation,
    all 6 auto/cross spectra, identical numpy seeds."""
    from orphics_b200 import maps, stats
    shape, wcs, so, wo, modl, ps = setup(128, 2.0, True, theory)
    og = omaps.MapGen(so, wo, ps)
    mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt))   # identical set-up input
    fc, ofc = maps.FourierCalc(shape, wcs, iau=iau), omaps.FourierCalc(so, wo, iau=iau)
    m = mg.get_map(seed=11, iau=iau)
    mo = og.get_map(seed=11, iau=iau)
    assert m.shape == (3, 128, 128)
    assert relerr(m, mo) < TOL64
    assert relerr(mg.get_map(seed=12, scalar=True), og.get_map(seed=12, scalar=True)) < TOL64
    p2d, k1, _ = fc.power2d(m)
    p2o, k1o, _ = ofc.power2d(mo)
    assert p2d.shape == (3, 3, 128, 128)
    assert relerr(k1, k1o) < TOL64
    for i in range(3):
        for j in range(3):
            assert np.max(np.abs(p2d[i, j] - p2o[i, j])) < TOL64 * np.max(np.abs(p2o[0, 0])) if i == 0 or j == 0 else True
            assert np.max(np.abs(p2d[i, j] - p2o[i, j])) < TOL64 * np.sqrt(np.max(np.abs(p2o[i, i])) * np.max(np.abs(p2o[j, j])))
    psk, _, _ = fc.power2d(m, skip_cross=True)
    assert np.all(psk[0, 1] == 0) and relerr(psk[1, 1], p2o[1, 1]) < TOL64
    edges = np.arange(200, 2600, 100.0)
    b = stats.bin2D(fc.geometry.modlmap(), edges, geometry=fc.geometry)
    ob = ostats.bin2D(modl, edges)
    bp = fc.binned_power_batch(b, m)[0]
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    for s, (i, j) in enumerate(pairs):
        ref = ob.bin(p2o[i, j])[1]
        scale = np.sqrt(np.abs(ob.bin(p2o[i, i])[1] * ob.bin(p2o[j, j])[1]))
        assert np.max(np.abs(bp[s] - ref) / scale) < TOL64
    # QU -> EB -> QU identity through the device: E,B of a sim built from E,B noise
    k = fc.iqu2teb(m, normalize=True)
    assert relerr(k, ofc.iqu2teb(mo, normalize=True)) < TOL64


@pytest.mark.parametrize("mode", ["philox", "philox_hermitian"])
@pytest.mark.parametrize("pol", [False, True])
def test_philox_modes_match_oracle_noise(mode, pol, theory):
    """Throughput noise modes: the device's Philox field reproduced in numpy and pushed
    through the restated reference algorithm agrees to 1e-10."""
    from orphics_b200 import maps
    shape, wcs, so, wo, modl, ps = setup(96, 3.0, pol, theory)
    og = omaps.MapGen(so, wo, ps)
    mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt), noise=mode, max_batch=3)
    nc = 3 if pol else 1
    seeds = [1000, 1001, 2 ** 40 + 5]
    ms = mg.get_maps(seeds)
    for i, s in enumerate(seeds):
        rand = philox_np.noise_field(s, nc, 96, 96, hermitian=(mode == "philox_hermitian"))
        rand = oenmap.ndmap(rand if pol else rand[0], wo)
        mo = og.map_from_noise(rand)
        if mode == "philox_hermitian":
            # Hermitian noise makes ifft(covsqrt*R) real already; the reference's .real then
            # keeps all of it, so the unit-mean |R|^2 convention gives the same spectrum
            pass
        assert relerr(ms[i], mo) < TOL64
    kh = mg.get_maps(seeds[:1], harm=True)
    rand = philox_np.noise_field(seeds[0], nc, 96, 96, hermitian=(mode == "philox_hermitian"))
    assert relerr(kh[0], og.map_from_noise(oenmap.ndmap(rand if pol else rand[0], wo), harm=True)) < TOL64


@pytest.mark.parametrize("path", ["cufft", "fused"])
@pytest.mark.parametrize("pol", [False, True])
def test_fused_pipeline_matches_oracle_and_accumulates_statistics(pol, path, theory, monkeypatch):
    """BASELINE config 2 path at test size: seeds -> bandpowers in one call, with taper,
    in seed-parity (numpy noise) and Philox modes; Statistics triple on the device.
    Both implementations: cuFFT passes, and the hand-written fused FFT kernels."""
    from orphics_b200 import maps, stats
    monkeypatch.setenv("ORPHX_PIPELINE", path)
    npix = 512 if path == "fused" else 256     # the fused kernels need ny >= 512
    shape, wcs, so, wo, modl, ps = setup(npix, 2.0, pol, theory)
    taper, w2 = maps.get_taper(shape, wcs)
    otaper = np.asarray(omaps.get_taper(so, wo)[0])
    edges = np.arange(200, 2600, 100.0)
    og, ofc, ob = omaps.MapGen(so, wo, ps), omaps.FourierCalc(so, wo), ostats.bin2D(modl, edges)
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)] if pol else [(0, 0)]

    def oracle_bp(m):
        p2o = ofc.power2d(oenmap.ndmap(np.asarray(m) * otaper, wo))[0]
        if not pol:
            return np.array([ob.bin(p2o)[1]])
        return np.array([ob.bin(p2o[i, j])[1] for i, j in pairs])

    for mode in ("numpy", "philox", "philox_hermitian"):
        mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt), noise=mode, max_batch=4)
        fc = maps.FourierCalc(shape, wcs, max_batch=4)
        b = stats.bin2D(fc.geometry.modlmap(), edges, geometry=fc.geometry)
        pipe = maps.SimPipeline(mg, fc, b, window=np.asarray(taper))
        assert pipe.path == path
        seeds = [1000 + i for i in range(6)]            # 6 sims with max_batch 4: two chunks
        bp = pipe.run(seeds, keep_maps=True)
        assert bp.shape == (6, len(pairs), len(edges) - 1)
        want = []
        for s in seeds:
            if mode == "numpy":
                mo = og.get_map(seed=s)
            else:
                rand = philox_np.noise_field(s, 3 if pol else 1, npix, npix, hermitian=(mode == "philox_hermitian"))
                mo = og.map_from_noise(oenmap.ndmap(rand if pol else rand[0], wo))
            want.append(oracle_bp(mo))
        if path == "fused":
            last = pipe.last_maps(2)                     # sims 5,6 (second chunk), before the taper
            assert relerr(last[1] if pol else last[1, 0], mo) < TOL64
        want = np.array(want)
        auto = {0: 0, 1: 3, 2: 5}
        for s, (i, j) in enumerate(pairs):
            scale = np.sqrt(np.abs(want[:, auto[i] if pol else 0] * want[:, auto[j] if pol else 0]))
            assert np.max(np.abs(bp[:, s] - want[:, s]) / scale) < TOL64
        N, S, Cm = pipe.stats()
        x = bp.reshape(6, -1)
        assert N == 6
        np.testing.assert_allclose(S, x.sum(0), rtol=1e-13)
        np.testing.assert_allclose(Cm, x.T @ x, rtol=1e-12, atol=1e-14 * np.abs(Cm).max())
        bp2 = pipe.run(seeds)
        assert np.array_equal(bp, bp2)                  # deterministic
        assert pipe.stats()[0] == 12


def test_fp32_mode_within_1e5(theory):
    from orphics_b200 import maps, stats
    shape, wcs, so, wo, modl, ps = setup(128, 2.0, False, theory)
    og, ofc = omaps.MapGen(so, wo, ps), omaps.FourierCalc(so, wo)
    mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt), dtype=np.float32)
    fc = maps.FourierCalc(shape, wcs, dtype=np.float32)
    m = mg.get_map(seed=21)
    mo = og.get_map(seed=21)
    assert m.dtype == np.float32
    assert relerr(m, mo) < TOL32
    p2d, k1, _ = fc.power2d(m)
    p2o, k1o, _ = ofc.power2d(oenmap.ndmap(np.asarray(m, dtype=np.float64), wo))
    assert p2d.dtype == np.float32 and k1.dtype == np.complex64
    assert relerr(k1, k1o) < TOL32 and relerr(p2d, p2o) < TOL32
    b = stats.bin2D(fc.geometry.modlmap(), EDGES, geometry=fc.geometry)
    bp = fc.binned_power_batch(b, m)[0, 0]
    bp_close(bp, ostats.bin2D(modl, EDGES).bin(p2o)[1], TOL32)


def test_sim_power_bin_recovers_theory_philox(theory):
    """Statistical acceptance (tutorials/demo-grf.ipynb:159-161) in the throughput mode."""
    from orphics_b200 import maps, stats
    shape, wcs, so, wo, modl, ps = setup(256, 2.0, False, theory)
    edges = np.arange(200, 3000, 80.0)
    for mode in ("philox", "philox_hermitian"):
        mg = maps.MapGen(shape, wcs, ps, noise=mode, max_batch=16)
        fc = maps.FourierCalc(shape, wcs, max_batch=16)
        b = stats.bin2D(fc.geometry.modlmap(), edges, geometry=fc.geometry)
        pipe = maps.SimPipeline(mg, fc, b)
        bp = pipe.run(range(64))[:, 0]
        ratio = bp.mean(0) / theory.lCl("TT", b.centers)
        assert abs(ratio.mean() - 1) < 0.01, ratio.mean()
        # exact expectation of the estimator: E[p2d] = covsqrt^2 * area / Npix, binned
        expect = b.bin(np.asarray(mg.covsqrt)[0, 0] ** 2 * fc.geometry.area / fc.geometry.npix)[1]
        nmodes = b.slot_counts[1:-1]
        sigma = np.sqrt(2.0 / (nmodes * 64.0))           # Gaussian-field bandpower scatter
        assert np.all(np.abs(bp.mean(0) / expect - 1) < 5 * sigma)


def test_large_map_properties_2048(theory):
    """BASELINE config 2 size (2048^2): size-independent properties instead of the oracle --
    Parseval through the device FFT, linearity of the binner, sum of bandpowers x counts =
    total power inside the binned annuli."""
    from orphics_b200 import maps, stats
    shape, wcs = maps.rect_geometry(width_arcmin=2048 * 0.5, px_res_arcmin=0.5)
    fc = maps.FourierCalc(shape, wcs, max_batch=2)
    g = fc.geometry
    modl = g.modlmap()
    ells = np.arange(0, modl.max() + 1, 1.0)
    from orphics_b200 import cosmology
    th = cosmology.default_theory()
    ps = cosmology.power_from_theory(ells, th, lensed=True, pol=False)
    mg = maps.MapGen(shape, wcs, ps, noise="philox", max_batch=2)
    m = mg.get_maps([1, 2])
    assert m.shape == (2, 2048, 2048)
    p2d, k, _ = fc.power2d(m[0])
    npix = 2048.0 ** 2
    assert abs(np.sum(np.abs(k) ** 2) / npix / np.sum(np.asarray(m[0]) ** 2) - 1) < 1e-12     # Parseval
    b = stats.bin2D(modl, EDGES, geometry=g)
    c, p1, n = b.bin(p2d, get_count=True)
    inside = (modl > EDGES[0]) & (modl <= EDGES[-1])
    assert n.sum() == inside.sum()
    assert abs(np.sum(p1 * n) / np.sum(np.asarray(p2d)[inside]) - 1) < 1e-12                 # checksum of checksums
    bp = fc.binned_power_batch(b, m)
    np.testing.assert_allclose(bp[0, 0], p1, rtol=1e-11)                                      # fused == unfused
    c, p1b = b.bin(2.5 * np.asarray(p2d))
    np.testing.assert_allclose(p1b, 2.5 * p1, rtol=1e-13)                                     # linearity
    ratio = bp[:, 0].mean(0) / th.lCl("TT", b.centers)
    assert abs(ratio.mean() - 1) < 0.02


def test_fused_and_cufft_paths_agree_at_2048(monkeypatch):
    """BASELINE config 2 size: the hand-written FFT path and the cuFFT path are two independent
    implementations of the same pipeline; their bandpowers agree to 1e-10 at 2048^2 (T and IQU),
    and the stored real-space maps agree with MapGen.get_maps."""
    from orphics_b200 import maps, stats, cosmology
    th = cosmology.default_theory()
    for pol, nb in ((False, 4), (True, 2)):
        shape, wcs = maps.rect_geometry(width_arcmin=2048 * 0.5, px_res_arcmin=0.5, pol=pol)
        g = maps.Geometry.get(shape, wcs)
        ps = cosmology.power_from_theory(np.arange(0, g.modlmap().max() + 1, 1.0), th, lensed=True, pol=pol)
        taper = np.asarray(maps.get_taper(shape, wcs)[0])
        out = {}
        for path in ("cufft", "fused"):
            monkeypatch.setenv("ORPHX_PIPELINE", path)
            mg = maps.MapGen(shape, wcs, ps, noise="philox_hermitian", max_batch=nb)
            fc = maps.FourierCalc(shape, wcs, max_batch=nb)
            b = stats.bin2D(g.modlmap(), EDGES, geometry=g)
            pipe = maps.SimPipeline(mg, fc, b, window=taper)
            assert pipe.path == path
            out[path] = pipe.run(range(50, 50 + nb), keep_maps=True)
            if path == "fused":
                stored = pipe.last_maps(nb)
                direct = mg.get_maps(range(50, 50 + nb))
                assert relerr(stored.reshape(np.shape(direct)), direct) < TOL64
        a, c = out["fused"], out["cufft"]
        auto = [0, 3, 5] if pol else [0]
        pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)] if pol else [(0, 0)]
        for s, (i, j) in enumerate(pairs):
            scale = np.sqrt(np.abs(c[:, auto[i]] * c[:, auto[j]]))
            assert np.max(np.abs(a[:, s] - c[:, s]) / scale) < TOL64


def test_fused_and_cufft_paths_agree_at_8192(monkeypatch):
    """The largest map of the BASELINE sweep (8192^2, one component): the hand-written kernels (ny = 8192 columns of 512
    threads, nx/2 = 4096 row transforms) against the cuFFT path, bandpowers and the stored map, float64."""
    from orphics_b200 import maps, stats, cosmology
    th = cosmology.default_theory()
    shape, wcs = maps.rect_geometry(width_arcmin=8192 * 0.5, px_res_arcmin=0.5)
    g = maps.Geometry.get(shape, wcs)
    ps = cosmology.power_from_theory(np.arange(0, g.modlmap().max() + 1, 1.0), th, lensed=True, pol=False)
    taper = np.asarray(maps.get_taper(shape, wcs)[0])
    out = {}
    mg0 = maps.MapGen(shape, wcs, ps, noise="philox_hermitian", max_batch=1)
    cs = np.asarray(mg0.covsqrt)
    del mg0
    for path in ("cufft", "fused"):
        monkeypatch.setenv("ORPHX_PIPELINE", path)
        mg = maps.MapGen(shape, wcs, covsqrt=cs, noise="philox_hermitian", max_batch=1)
        fc = maps.FourierCalc(shape, wcs, max_batch=1)
        b = stats.bin2D(g.modlmap(), EDGES, geometry=g)
        pipe = maps.SimPipeline(mg, fc, b, window=taper)
        assert pipe.path == path
        out[path] = pipe.run([77], keep_maps=True)
        if path == "fused":
            stored = pipe.last_maps(1)
            assert relerr(stored.reshape(shape), mg.get_maps([77])[0]) < TOL64
        del pipe, mg, fc, b
    a, c = out["fused"], out["cufft"]
    assert np.max(np.abs(a[:, 0] - c[:, 0]) / np.abs(c[:, 0])) < TOL64


@pytest.mark.parametrize("pol", [False, True])
def test_fp32_fused_pipeline_philox_within_1e5(pol, theory):
    """float32 mode on the hand-written kernels (single-precision Box-Muller, FFTs and binning) against the
    fp64 definition of the noise pushed through the restated reference algorithm: maps and bandpowers 1e-5."""
    from orphics_b200 import maps, stats
    npix = 512
    shape, wcs, so, wo, modl, ps = setup(npix, 2.0, pol, theory)
    edges = np.arange(200, 2600, 100.0)
    og, ofc, ob = omaps.MapGen(so, wo, ps), omaps.FourierCalc(so, wo), ostats.bin2D(modl, edges)
    mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt), noise="philox_hermitian", dtype=np.float32, max_batch=2)
    fc = maps.FourierCalc(shape, wcs, dtype=np.float32, max_batch=2)
    b = stats.bin2D(fc.geometry.modlmap(), edges, geometry=fc.geometry)
    pipe = maps.SimPipeline(mg, fc, b)
    assert pipe.path == "fused"
    bp = pipe.run([1000, 1001], keep_maps=True)
    last = pipe.last_maps(2)
    for i, s in enumerate((1000, 1001)):
        rand = philox_np.noise_field(s, 3 if pol else 1, npix, npix, hermitian=True)
        mo = og.map_from_noise(oenmap.ndmap(rand if pol else rand[0], wo))
        got = last[i] if pol else last[i, 0]
        assert got.dtype == np.float32
        assert relerr(got, mo) < TOL32
        p2o = ofc.power2d(mo)[0]
        want = ob.bin(p2o[0, 0] if pol else p2o)[1]
        assert np.max(np.abs(bp[i, 0] - want) / np.max(np.abs(want))) < TOL32


def test_multi_pow_on_device_matches_eigpow(theory):
    """enmap.multi_pow for a 2-D covariance (maps.py:1571) as a per-pixel Jacobi eigen-power on the device vs the
    oracle's numpy eigh: random SPD matrices, exactly singular ones (B = 0, |TE|^2 = TT EE), zero and negative
    eigenvalues (clipped for the square root), n = 1..3, integer and fractional exponents; then MapGen built from
    a 4-D covariance reproduces the map of the 3-D one."""
    from orphics_b200 import maps
    rng = np.random.RandomState(17)
    npix = 4000
    for n in (1, 2, 3):
        A = rng.standard_normal((npix, n, n))
        S = A @ A.transpose(0, 2, 1)
        # (ridge: condition numbers <~ 100, so that LAPACK's absolute eigenvalue accuracy does not limit the
        # comparison of the inverse powers at 1e-10)
        S = (S + 0.3 * np.eye(n) * np.trace(S, axis1=1, axis2=2)[:, None, None] / n) * rng.uniform(1e-3, 1e3, (npix, 1, 1))
        S[:50] = 0.0                                     # zero matrices
        if n > 1:
            S[50:100, -1, :] = 0.0; S[50:100, :, -1] = 0.0           # exactly singular (one vanishing component)
            v = rng.standard_normal((50, n, 1)); S[100:150] = v @ v.transpose(0, 2, 1)   # rank one
            S[150:200] -= 0.6 * np.eye(n) * np.trace(S[150:200], axis1=1, axis2=2)[:, None, None] / n   # indefinite
        mat = np.ascontiguousarray(S.transpose(1, 2, 0)).reshape(n, n, 40, 100)
        for e in (0.5, -0.5, -1.0, 2.0, 1.0):
            got = maps.multi_pow(mat, e)
            want = oenmap.eigpow(mat, e)
            scale = np.max(np.abs(want.reshape(n * n, -1)), axis=0).reshape(40, 100) + 1e-300
            assert np.max(np.abs(got - want) / scale) < 1e-10, (n, e)
    # MapGen from a 2-D (4-D array) covariance: same covsqrt, same map as from the 1-D spectra
    shape, wcs, so, wo, modl, ps = setup(128, 2.0, True, theory)
    og = omaps.MapGen(so, wo, ps)
    cs3 = np.asarray(og.covsqrt)
    cov4 = np.einsum("abyx,cbyx->acyx", cs3, cs3)         # covsqrt^2 per pixel, already in pixel units
    mg = maps.MapGen(shape, wcs, cov4, pixel_units=True)
    assert relerr(mg.covsqrt, cs3) < 1e-9
    assert relerr(mg.get_map(seed=3), og.get_map(seed=3)) < 1e-9


@pytest.mark.parametrize("pol", [False, True])
def test_tma_row_pass_matches_the_legacy_row_pass_and_the_oracle(pol, theory, monkeypatch):
    """The persistent TMA row kernels (ox_row_tma.cuh: cp.async.bulk.tensor tiles, three slots, two groups per CTA;
    ox_row_w32.cuh: the same with one warp per row and two radix-32 stages) agree with the one-tile-per-CTA kernel (the compiler contracts multiply-adds differently in
    the two instantiations, so agreement is to rounding, 1e-13, not bit for bit), on a 512 x 2048 patch (nx/2 = 1024:
    the 2048^2 configuration's row length), with the separable-window fast path and with the general 2-D window --
    and matches the oracle on numpy seeds."""
    from orphics_b200 import maps, stats
    ny, nx, res = 512, 2048, 1.0
    shape, wcs = maps.rect_geometry(width_arcmin=nx * res, px_res_arcmin=res, height_arcmin=ny * res, pol=pol)
    so, wo = omaps.rect_geometry(width_arcmin=nx * res, px_res_arcmin=res, height_arcmin=ny * res, pol=pol)
    assert tuple(shape[-2:]) == (ny, nx)
    modl = np.asarray(oenmap.modlmap(so, wo))
    ps = otheory.power_from_theory(np.arange(0, modl.max() + 1, 1.0), theory, lensed=True, pol=pol)
    og, ofc, ob = omaps.MapGen(so, wo, ps), omaps.FourierCalc(so, wo), ostats.bin2D(modl, EDGES)
    taper = np.asarray(maps.get_taper(shape, wcs)[0])
    nsim = 5                                                     # 5 planes x 128 row tiles: odd tile counts per CTA
    out = {}
    variants = ("tma", "tma_general_window", "tma4", "tma4_general_window", "w32", "w32_general_window")
    for kb in ("legacy",) + variants:
        monkeypatch.setenv("ORPHX_KB", kb.split("_")[0])
        monkeypatch.setenv("ORPHX_WINDOW_SEPARABLE", "0" if "general" in kb else "1")
        mg = maps.MapGen(shape, wcs, ps, noise="numpy", max_batch=nsim)
        fc = maps.FourierCalc(shape, wcs, max_batch=nsim)
        b = stats.bin2D(fc.geometry.modlmap(), EDGES, geometry=fc.geometry)
        pipe = maps.SimPipeline(mg, fc, b, window=taper)
        assert pipe.path == "fused"
        bp = pipe.run(range(40, 40 + nsim), keep_maps=True)
        out[kb] = (bp, pipe.last_maps(nsim))
    for kb in variants:
        assert np.array_equal(np.isnan(out[kb][0]), np.isnan(out["legacy"][0]))
        fin = np.isfinite(out["legacy"][0])
        assert np.max(np.abs(out[kb][0] - out["legacy"][0])[fin] / np.abs(out["legacy"][0])[fin].max()) < 1e-13
        assert relerr(out[kb][1], out["legacy"][1]) < 1e-13
    bp, stored = out["tma"]       # (the default)
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)] if pol else [(0, 0)]
    auto = {0: 0, 1: 3, 2: 5}
    for i in (0, nsim - 1):
        mo = og.get_map(seed=40 + i)
        assert relerr(stored[i] if pol else stored[i, 0], mo) < TOL64
        p2o = ofc.power2d(oenmap.ndmap(np.asarray(mo) * taper, wo))[0]
        want = np.array([ob.bin(p2o[a, c] if pol else p2o)[1] for a, c in pairs])
        for s, (a, c) in enumerate(pairs):
            scale = np.sqrt(np.abs(want[auto[a] if pol else 0] * want[auto[c] if pol else 0]))
            ok = np.isfinite(want[s])
            assert np.max(np.abs(bp[i, s] - want[s])[ok] / scale[ok]) < TOL64


@pytest.mark.parametrize("pol", [False, True])
def test_reference_options_real_noise_phys_normalisation_complex_filter(pol, theory):
    """Options of the reference signatures that round 1 refused: MapGen.get_map(real=True) (maps.py:1578: white noise drawn
    in real space, unitary transform), FourierCalc.iqu2teb(normalize='phys') (enmap.fft's physical normalisation, used at
    lensing.py:403,653) and maps.filter_map with a complex kfilter (maps.py:1923)."""
    from orphics_b200 import maps
    shape, wcs, so, wo, modl, ps = setup(96, 3.0, pol, theory)
    og, ofc = omaps.MapGen(so, wo, ps), omaps.FourierCalc(so, wo)
    mg = maps.MapGen(shape, wcs, covsqrt=np.asarray(og.covsqrt))
    fc = maps.FourierCalc(shape, wcs)
    for kw in (dict(real=True), dict(real=True, scalar=True), dict(real=True, harm=True)):
        assert relerr(mg.get_map(seed=8, **kw), og.get_map(seed=8, **kw)) < TOL64, kw
    m = og.get_map(seed=9)
    for rot in (True, False):
        got = fc.iqu2teb(maps.ndmap(np.asarray(m), wcs), normalize="phys", rot=rot)
        assert relerr(got, ofc.iqu2teb(m, normalize="phys", rot=rot)) < TOL64
    rng = np.random.RandomState(4)
    kf = np.exp(-(modl / 2000.0) ** 2) * np.exp(1j * rng.uniform(0, 2 * np.pi, size=modl.shape))     # no symmetry at all
    got = maps.filter_map(maps.ndmap(np.asarray(m), wcs), kf)
    assert got.dtype == np.float64 and relerr(got, omaps.filter_map(m, kf)) < TOL64
    assert relerr(maps.filter_map(maps.ndmap(np.asarray(m), wcs), kf.real + 0j), omaps.filter_map(m, kf.real)) < TOL64
    if pol:
        # a (Q,U) pair on its own: the reference rotates the LAST two components whatever their number (maps.py:1615)
        qu = np.asarray(m)[1:]
        fc2, ofc2 = maps.FourierCalc((2,) + tuple(shape[-2:]), wcs), omaps.FourierCalc((2,) + tuple(so[-2:]), wo)
        assert relerr(fc2.iqu2teb(maps.ndmap(qu, wcs), normalize=False), ofc2.iqu2teb(oenmap.ndmap(qu, wo), normalize=False)) < TOL64
        p2, k2, _ = fc2.power2d(maps.ndmap(qu, wcs))
        op2, ok2, _ = ofc2.power2d(oenmap.ndmap(qu, wo))
        assert np.shape(p2) == (2, 2) + tuple(shape[-2:]) and relerr(p2, op2) < TOL64 and relerr(k2, ok2) < TOL64
        assert relerr(p2[0, 0], fc.power2d(maps.ndmap(np.asarray(m), wcs))[0][1, 1]) < TOL64       # EE from (Q,U) == EE from (I,Q,U)
