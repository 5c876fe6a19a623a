"""Device-resident maps behind the reference signatures (enmap.devmap): the chained per-call sequence of the
reference's tutorials (tutorials/demo-grf.ipynb:52-161: get_map -> * taper -> power2d -> bin) must give the same
numbers whether the intermediate maps stay in HBM or go through numpy arrays, and a devmap must behave as the
numpy ndmap it stands for."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EDGES = np.arange(100, 3000, 40)


@pytest.fixture()
def setup():
    from orphics_b200 import maps, stats, cosmology, enmap
    npix = 256
    shape, wcs = maps.rect_geometry(width_arcmin=npix * 2.0, px_res_arcmin=2.0)
    theory = cosmology.default_theory()
    ells = np.arange(0, 6000, 1)
    ps = cosmology.power_from_theory(ells, theory, lensed=True, pol=False)
    return dict(maps=maps, stats=stats, enmap=enmap, shape=shape, wcs=wcs, ps=ps, theory=theory)


def _chain(S, device, noise="philox", seeds=(3, 4)):
    maps, stats, enmap = S["maps"], S["stats"], S["enmap"]
    old = enmap.DEVICE_RESIDENT
    enmap.DEVICE_RESIDENT = device
    try:
        mg = maps.MapGen(S["shape"], S["wcs"], S["ps"], noise=noise)
        fc = maps.FourierCalc(S["shape"], S["wcs"])
        taper, w2 = maps.get_taper(S["shape"], S["wcs"])
        binner = stats.bin2D(fc.geometry.modlmap(), EDGES)
        out = []
        for seed in seeds:
            m = mg.get_map(seed=seed)
            assert isinstance(m, enmap.devmap) == device
            p2d, k1, k2 = fc.power2d(m * taper)
            assert isinstance(p2d, enmap.devmap) == device
            cents, p1d = binner.bin(p2d / w2)
            out.append((np.asarray(m), np.asarray(p2d), np.asarray(k1), p1d))
        return out
    finally:
        enmap.DEVICE_RESIDENT = old


@pytest.mark.parametrize("noise", ["philox", "numpy"])
def test_chained_calls_agree_with_the_host_round_trip(setup, noise):
    dev = _chain(setup, True, noise)
    host = _chain(setup, False, noise)
    for d, h in zip(dev, host):
        for a, b in zip(d, h):
            assert a.shape == b.shape and a.dtype == b.dtype
            assert np.array_equal(a, b, equal_nan=True)


def test_devmap_behaves_as_the_ndmap_it_stands_for(setup):
    maps, enmap = setup["maps"], setup["enmap"]
    mg = maps.MapGen(setup["shape"], setup["wcs"], setup["ps"], noise="philox")
    m = mg.get_map(seed=11)
    assert isinstance(m, enmap.devmap)
    h = np.array(m)
    assert m.shape == h.shape and m.dtype == h.dtype and m.ndim == 2 and m.size == h.size and m.wcs is setup["wcs"]
    rng = np.random.RandomState(0)
    w = rng.uniform(0.5, 1.5, size=h.shape)
    wd = enmap.devmap.from_host(w)
    # device arithmetic == numpy arithmetic (one rounding per element either way)
    for got, want in [(m * w, h * w), (w * m, w * h), (m * wd, h * w), (m + wd, h + w), (m - wd, h - w), (wd - m, w - h),
                      (m / wd, h / w), (w / m, w / h), (m * 2.5, h * 2.5), (2.5 * m, 2.5 * h), (m / 3.0, h / 3.0),
                      (1.0 - m, 1.0 - h), (m + 1, h + 1), (np.multiply(w, m), w * h), (np.add(m, w), h + w)]:
        assert isinstance(got, enmap.devmap)
        assert np.array_equal(np.asarray(got), want)
    # everything else is numpy's: results are host arrays with numpy's values
    assert np.array_equal(m ** 2, h ** 2) and np.array_equal(-m, -h) and np.array_equal(abs(m), abs(h))
    assert np.array_equal(m > 0, h > 0) and np.array_equal(m[3:7, ::2], h[3:7, ::2]) and m[5, 6] == h[5, 6]
    assert m.mean() == h.mean() and np.sum(m) == np.sum(h) and np.array_equal(np.sqrt(abs(m)), np.sqrt(abs(h)))
    assert np.array_equal(m.copy(), h) and np.array_equal(m.reshape(-1), h.reshape(-1)) and len(m) == len(h)
    assert np.array_equal(m * np.float32(2), h * np.float32(2))          # promotion rules stay numpy's
    assert np.array_equal(m * w[0], h * w[0])                             # general broadcasting too
    assert np.array_equal(m * np.arange(h.shape[1]), h * np.arange(h.shape[1]))
    # item assignment edits the host copy; the next device use sees it
    m2 = m.copy()
    m2[0, :] = 7.0
    h2 = h.copy()
    h2[0, :] = 7.0
    assert np.array_equal(np.asarray(m2 * wd), h2 * w)
    assert np.array_equal(np.asarray(m), h)                               # the original is untouched
    # a ufunc writing INTO a devmap runs on the host copy; the device copy follows on next use
    m3 = m.copy()
    np.multiply(h, 3.0, out=m3)
    assert np.array_equal(np.asarray(m3 * wd), (h * 3.0) * w)
    # what np.asarray hands out is read-only (it is the cached host copy): editing goes through the devmap or a copy
    with pytest.raises(ValueError):
        np.asarray(m)[0, 0] = 1.0
    c = np.array(m)
    c[0, 0] = 1.0                                                         # np.array(m) is a private, writable copy
    assert np.asarray(m)[0, 0] == h[0, 0]
    # from_host keeps a private copy: editing the caller's array afterwards reaches neither copy
    w2 = w.copy()
    d2 = enmap.devmap.from_host(w2)
    w2[:] = 0.0
    assert np.array_equal(np.asarray(d2), w) and np.array_equal(np.asarray(m * d2), h * w)


def test_component_stacks_and_views(setup):
    maps, enmap, cosmology = setup["maps"], setup["enmap"], __import__("orphics_b200.cosmology", fromlist=["x"])
    shape = (3,) + tuple(setup["shape"])
    ps = cosmology.power_from_theory(np.arange(0, 6000, 1), setup["theory"], lensed=True, pol=True)
    mg = maps.MapGen(shape, setup["wcs"], ps, noise="philox")
    fc = maps.FourierCalc(shape, setup["wcs"])
    taper, w2 = maps.get_taper(shape, setup["wcs"])
    m = mg.get_map(seed=5)
    h = np.asarray(m)
    assert m.shape == shape
    mt = m * taper                                    # (3,Ny,Nx) x (Ny,Nx) on the device
    assert isinstance(mt, enmap.devmap) and np.array_equal(np.asarray(mt), h * np.asarray(taper))
    assert isinstance(m[1], enmap.devmap) and np.array_equal(np.asarray(m[1]), h[1]) and np.array_equal(np.asarray(m[-1]), h[2])
    p2d, k1, k2 = fc.power2d(mt)
    enmap.DEVICE_RESIDENT = False
    try:
        p2d_h, k1_h, _ = fc.power2d(h * np.asarray(taper))
    finally:
        enmap.DEVICE_RESIDENT = True
    assert np.array_equal(np.asarray(p2d), p2d_h) and np.array_equal(np.asarray(k1), k1_h)
    # complex k-map (op) real filter on the device; f2power / ifft / iqu2teb take devmaps
    kf = maps.gauss_beam(fc.geometry.modlmap(), 1.5)
    kk = k1[0] * kf
    assert isinstance(kk, enmap.devmap) and np.array_equal(np.asarray(kk), k1_h[0] * kf)
    pw = fc.f2power(k1[0], kk)
    assert isinstance(pw, enmap.devmap)
    assert np.allclose(np.asarray(pw), np.real(np.conj(k1_h[0]) * (k1_h[0] * kf)) * fc.normfact, rtol=1e-13, atol=0)
    back = fc.ifft(fc.fft(m[0]))
    assert np.max(np.abs(np.asarray(back).real - h[0])) < 1e-10 * np.max(np.abs(h[0]))
    f1 = maps.filter_map(m[0], kf, fc=maps.FourierCalc(setup["shape"], setup["wcs"]))
    enmap.DEVICE_RESIDENT = False
    try:
        f1_h = maps.filter_map(enmap.ndmap(h[0], setup["wcs"]), kf, fc=maps.FourierCalc(setup["shape"], setup["wcs"]))
    finally:
        enmap.DEVICE_RESIDENT = True
    assert isinstance(f1, enmap.devmap) and np.array_equal(np.asarray(f1), np.asarray(f1_h))


def test_estimator_consumes_device_kmaps(setup):
    from orphics_b200 import lensing
    maps, enmap = setup["maps"], setup["enmap"]
    shape, wcs, theory = setup["shape"], setup["wcs"], setup["theory"]
    fc = maps.FourierCalc(shape, wcs)
    modl = fc.geometry.modlmap()
    kbeam = maps.gauss_beam(modl, 1.5)
    n2d = modl * 0 + (1.0 * np.pi / 180. / 60.) ** 2
    tmask = maps.mask_kspace(shape, wcs, lmin=300, lmax=2000)
    kmask = maps.mask_kspace(shape, wcs, lmin=20, lmax=3000)
    q = lensing.qest(shape, wcs, theory, noise2d=n2d, beam2d=kbeam, kmask=tmask, kmask_K=kmask, unlensed_equals_lensed=True)
    mg = maps.MapGen(shape, wcs, setup["ps"], noise="philox")
    m = mg.get_map(seed=21)
    _, kT, _ = fc.power2d(m)
    assert isinstance(kT, enmap.devmap)
    rec = q.kappa_from_map("TT", kT, alreadyFTed=True)
    rec_m = q.kappa_from_map("TT", m)
    assert isinstance(rec, enmap.devmap) and isinstance(rec_m, enmap.devmap)
    enmap.DEVICE_RESIDENT = False
    try:
        rec_h = q.kappa_from_map("TT", np.asarray(kT), alreadyFTed=True)
        rec_mh = q.kappa_from_map("TT", np.asarray(m))
    finally:
        enmap.DEVICE_RESIDENT = True
    assert np.array_equal(np.asarray(rec), np.asarray(rec_h)) and np.array_equal(np.asarray(rec_m), np.asarray(rec_mh))
    # the tutorial's cross-power line (tt_verification.ipynb:612): power2d(recon, kappa) with device maps
    p, _, _ = fc.power2d(rec, m)
    p_h = fc.power2d(np.asarray(rec), np.asarray(m))[0]
    assert np.array_equal(np.asarray(p), np.asarray(p_h))
