"""GPU parity: lensing.qest through the C-ABI vs the numpy oracle on identical inputs (kappa maps and
kappa_hat(l) <= 1e-10 relative in fp64, 1e-5 in fp32), TT on the half-plane path and on the general
c2c path, EB, alreadyFTed / returnFt / separate Y leg / batches / mean-field stack."""
import numpy as np
import pytest

from conftest import relerr
from oracle import enmap_np as oenmap, maps_np as omaps, qe_np

pytestmark = pytest.mark.gpu
TOL64, TOL32 = 1e-10, 1e-5


def setup(npix, res, theory, masked=True):
    from orphics_b200 import maps, lensing, cosmology
    w = npix * res
    shape, wcs = maps.rect_geometry(width_arcmin=w, px_res_arcmin=res)
    so, wo = omaps.rect_geometry(width_arcmin=w, px_res_arcmin=res)
    modl = np.asarray(oenmap.modlmap(so, wo))
    beam = omaps.gauss_beam(modl, 1.5)
    n2d = np.zeros(so) + (1.0 * np.pi / 180 / 60) ** 2
    if masked:
        tm = np.asarray(omaps.mask_kspace(so, wo, lmin=300, lmax=2000))
        km = np.asarray(omaps.mask_kspace(so, wo, lmin=20, lmax=3500))
    else:
        tm = km = None
    kw = dict(noise2d=n2d, beam2d=beam, kmask=tm, noise2d_P=2 * n2d, kmask_P=tm, kmask_K=km, pol=True, grad_cut=None,
              unlensed_equals_lensed=True, bigell=9000)               # tutorials/tt_verification.ipynb:81
    qo = qe_np.qest(so, wo, theory, **kw)
    q = lensing.qest(shape, wcs, cosmology.default_theory(), max_batch=3, **kw)
    return shape, wcs, so, wo, q, qo


@pytest.mark.parametrize("masked", [True, False])
def test_filters_and_normalisation_match_oracle(masked, theory):
    shape, wcs, so, wo, q, qo = setup(128, 2.0, theory, masked)
    for XY in ("TT", "EB"):
        assert relerr(q.N.WXY(XY), qo.N.WXY(XY)) < 1e-14
        assert relerr(q.N.Nlkk[XY], qo.N.Nlkk[XY]) < 1e-9
        assert relerr(q.N.AL[XY], qo.N.AL[XY]) < 1e-9
    assert q._plans["TT"][1] == masked        # half-plane path only when the filters vanish at Nyquist


@pytest.mark.parametrize("masked", [True, False])
def test_device_setup_matches_host_setup(masked, theory, monkeypatch):
    """The device set-up (interpolated 2-D spectra, ox_qe_filter, ox_qe_norm: SURVEY 8f-2) against the numpy set-up of
    round 1 (ORPHX_QE_SETUP=host) and the oracle: filters to rounding, A_L to the conditioning of its FFT convolutions,
    same path selection, same kappa."""
    from orphics_b200 import lensing, enmap
    shape, wcs, so, wo, q, qo = setup(128, 2.0, theory, masked)
    assert q.N.device and isinstance(q.N.AL["TT"], enmap.devmap)
    monkeypatch.setenv("ORPHX_QE_SETUP", "host")
    _, _, _, _, qh, _ = setup(128, 2.0, theory, masked)
    assert not qh.N.device
    for XY in ("TT", "EB"):
        assert relerr(q.N.WXY(XY), qh.N.WXY(XY)) < 1e-14 and relerr(q.N.WXY(XY), qo.N.WXY(XY)) < 1e-14
        YY = XY[1] + XY[1]
        assert relerr(q.N.WY(YY), qh.N.WY(YY)) < 1e-14 and relerr(q.N.WY(YY), qo.N.WY(YY)) < 1e-14
        assert relerr(q.N.AL[XY], qh.N.AL[XY]) < 1e-10 and relerr(q.N.Nlkk[XY], qh.N.Nlkk[XY]) < 1e-10
        assert q.path(XY) == qh.path(XY)
    for k in ("TT", "EE", "TE"):
        assert relerr(q.N.lClFid2d[k], qo.N.lCl[k]) < 1e-14
    T = np.random.RandomState(5).standard_normal(shape) * 50
    assert relerr(q.kappa_from_map("TT", T), qh.kappa_from_map("TT", T)) < 1e-9


@pytest.mark.parametrize("masked,npix,path", [(True, 128, "half"), (False, 128, "c2c"), (True, 512, "fused")])
def test_tt_kappa_matches_oracle(masked, npix, path, theory):
    shape, wcs, so, wo, q, qo = setup(npix, 2.0, theory, masked)
    assert q.path("TT") == path
    qo.N.AL["TT"] = np.asarray(q.N.AL["TT"])          # identical set-up input: compare the per-map chain only
    rng = np.random.RandomState(3)
    T = rng.standard_normal(shape) * 50
    T2 = rng.standard_normal(shape) * 50
    k, ko = q.kappa_from_map("TT", T), qo.kappa_from_map("TT", T)
    assert k.shape == shape and k.dtype == np.float64
    assert relerr(k, ko) < TOL64
    kT = np.fft.fft2(T)
    assert relerr(q.kappa_from_map("TT", kT, alreadyFTed=True), ko) < TOL64
    kf, kfo = q.kappa_from_map("TT", T, returnFt=True), qo.kappa_from_map("TT", T, returnFt=True)
    assert kf.dtype == np.complex128 and relerr(kf, kfo) < TOL64
    assert relerr(q.reconstruct("TT", T, T2DDataY=T2), qo.kappa_from_map("TT", T, T2DDataY=T2)) < TOL64   # separate Y leg
    # batch + mean-field stack (Statistics.add_stack semantics)
    stack = np.stack([T, T2, T + T2])
    q.reset_meanfield("TT")
    kb = q.kappa_from_maps("TT", stack, returnFt=True, accumulate_meanfield=True)
    want = np.stack([qo.kappa_from_map("TT", m, returnFt=True) for m in stack])
    assert relerr(kb, want) < TOL64
    acc, cnt = q.meanfield("TT")
    assert cnt == 3
    tot = want.sum(0)                                 # the stack keeps the Hermitian part (= what the kappa MAP sees)
    iy, ix = (-np.arange(shape[0])) % shape[0], (-np.arange(shape[1])) % shape[1]
    herm = 0.5 * (tot + np.conj(tot[iy][:, ix]))
    assert relerr(acc, herm[:, :shape[1] // 2 + 1]) < TOL64
    kb2 = q.kappa_from_maps("TT", stack, accumulate_meanfield=True)
    assert q.meanfield("TT")[1] == 6
    assert relerr(kb2, np.stack([qo.kappa_from_map("TT", m) for m in stack])) < TOL64


@pytest.mark.parametrize("npix,path", [(128, "c2c"), (512, "fused_eb")])
def test_eb_kappa_matches_oracle(npix, path, theory):
    from orphics_b200 import maps
    shape, wcs, so, wo, q, qo = setup(npix, 2.0, theory, True)
    assert q.path("EB") == path
    qo.N.AL["EB"] = np.asarray(q.N.AL["EB"])
    rng = np.random.RandomState(4)
    iqu = rng.standard_normal((3,) + shape) * np.array([50., 3., 3.])[:, None, None]
    fc = maps.FourierCalc((3,) + shape, wcs)
    _, kteb, _ = fc.power2d(iqu)                                   # tutorials/tt_verification.ipynb:606
    ofc = omaps.FourierCalc((3,) + so, wo)
    _, ktebo, _ = ofc.power2d(oenmap.ndmap(iqu, wo))
    assert relerr(kteb, ktebo) < TOL64
    k = q.kappa_from_map("EB", kteb[0], kteb[1], kteb[2], alreadyFTed=True)      # ipynb:610
    ko = qo.kappa_from_map("EB", np.asarray(ktebo[0]), np.asarray(ktebo[1]), np.asarray(ktebo[2]), alreadyFTed=True)
    assert relerr(k, ko) < TOL64
    kf = q.kappa_from_map("EB", kteb[0], kteb[1], kteb[2], alreadyFTed=True, returnFt=True)
    assert relerr(kf, qo.kappa_from_map("EB", ktebo[0], ktebo[1], ktebo[2], alreadyFTed=True, returnFt=True)) < TOL64
    E, B = rng.standard_normal((2,) + shape)
    assert relerr(q.kappa_from_map("EB", None, E, B), qo.kappa_from_map("EB", None, E, B)) < TOL64     # real E/B maps in
    # k-maps that are NOT Hermitian (transforms of complex fields) must take the reference's c2c chain on every path
    kE = np.fft.fft2(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    kB = np.fft.fft2(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    got = q.kappa_from_map("EB", None, kE, kB, alreadyFTed=True, returnFt=True)
    assert relerr(got, qo.kappa_from_map("EB", None, kE, kB, alreadyFTed=True, returnFt=True)) < TOL64
    gotT = q.kappa_from_map("TT", kE, alreadyFTed=True, returnFt=True)
    qo.N.AL["TT"] = np.asarray(q.N.AL["TT"])
    assert relerr(gotT, qo.kappa_from_map("TT", kE, alreadyFTed=True, returnFt=True)) < TOL64
    # batches with the mean-field stack
    Es, Bs = rng.standard_normal((2, 3) + shape)
    q.reset_meanfield("EB")
    kb = q.kappa_from_maps("EB", Es, Bs, returnFt=True, accumulate_meanfield=True)
    want = np.stack([qo.kappa_from_map("EB", None, e, b, returnFt=True) for e, b in zip(Es, Bs)])
    assert relerr(kb, want) < TOL64
    acc, cnt = q.meanfield("EB")
    iy, ix = (-np.arange(shape[0])) % shape[0], (-np.arange(shape[1])) % shape[1]
    tot = want.sum(0)
    herm = 0.5 * (tot + np.conj(tot[iy][:, ix]))
    assert cnt == 3 and relerr(acc, herm[:, :shape[1] // 2 + 1]) < TOL64


def test_tt_fused_equals_cufft_rectangular(theory, monkeypatch):
    """Size-independent check at a size the oracle does not reach: the hand-written FFT passes and the
    cuFFT chain are two implementations of the same estimator (1024 x 2048, batch, mean field)."""
    from orphics_b200 import maps, lensing, cosmology
    ny, nx, res = 1024, 2048, 1.0
    shape, wcs = maps.rect_geometry(width_arcmin=nx * res, px_res_arcmin=res, height_arcmin=ny * res)
    assert tuple(shape) == (ny, nx)
    modl = maps.Geometry.get(shape, wcs).modlmap()
    kw = dict(noise2d=np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2, beam2d=maps.gauss_beam(modl, 1.5),
              kmask=maps.mask_kspace(shape, wcs, lmin=300, lmax=2000), kmask_K=maps.mask_kspace(shape, wcs, lmin=20, lmax=3500),
              unlensed_equals_lensed=True, max_batch=2)
    qf = lensing.qest(shape, wcs, cosmology.default_theory(), **kw)
    assert qf.path("TT") == "fused"
    monkeypatch.setenv("ORPHX_QE", "cufft")
    qc = lensing.qest(shape, wcs, cosmology.default_theory(), **kw)
    assert qc.path("TT") == "half"
    monkeypatch.delenv("ORPHX_QE")
    rng = np.random.RandomState(11)
    T = rng.standard_normal((2,) + tuple(shape)) * 50
    for q in (qf, qc):
        q.reset_meanfield("TT")
    a = qf.kappa_from_maps("TT", T, returnFt=True, accumulate_meanfield=True)
    b = qc.kappa_from_maps("TT", T, returnFt=True, accumulate_meanfield=True)
    assert relerr(a, b) < TOL64
    assert relerr(qf.meanfield("TT")[0], qc.meanfield("TT")[0]) < TOL64
    assert relerr(qf.kappa_from_maps("TT", T), qc.kappa_from_maps("TT", T)) < TOL64
    assert relerr(qf.kappa_from_map("TT", np.fft.fft2(T[0]), alreadyFTed=True), qc.kappa_from_map("TT", T[0])) < TOL64


@pytest.mark.parametrize("npix", [128, 512])
def test_fp32_mode(npix, theory):
    from orphics_b200 import lensing, cosmology
    shape, wcs, so, wo, q, qo = setup(npix, 2.0, theory, True)
    q32 = lensing.qest(shape, wcs, cosmology.default_theory(), noise2d=qo.N.noise["TT"], beam2d=qo.N.beam,
                       kmask=qo.N.fmask["TT"], kmask_K=qo.N.fmaskK, unlensed_equals_lensed=True, dtype=np.float32)
    rng = np.random.RandomState(5)
    T = (rng.standard_normal(shape) * 50).astype(np.float32)
    assert q32.path("TT") == ("fused" if npix >= 512 else "half")
    k = q32.kappa_from_map("TT", T)
    assert k.dtype == np.float32
    assert relerr(k, qo.kappa_from_map("TT", T.astype(np.float64))) < TOL32
    # EB in float32 (BASELINE configs[4]): real E/B maps in, kappa map and kappa_hat(l) out
    q32p = lensing.qest(shape, wcs, cosmology.default_theory(), noise2d=qo.N.noise["TT"], noise2d_P=qo.N.noise["EE"], beam2d=qo.N.beam,
                        kmask=qo.N.fmask["TT"], kmask_P=qo.N.fmask["EE"], kmask_K=qo.N.fmaskK, unlensed_equals_lensed=True,
                        pol=True, dtype=np.float32, max_batch=2)
    assert q32p.path("EB") == ("fused_eb" if npix >= 512 else "c2c")
    qo.N.AL["EB"] = np.asarray(q32p.N.AL["EB"], dtype=np.float64)
    E, B = (rng.standard_normal((2, 2) + shape) * 3).astype(np.float32)
    kb = q32p.kappa_from_maps("EB", E, B)
    want = np.stack([qo.kappa_from_map("EB", None, e.astype(np.float64), b.astype(np.float64)) for e, b in zip(E, B)])
    assert kb.dtype == np.float32 and relerr(kb, want) < TOL32
    kf = q32p.kappa_from_maps("EB", E, B, returnFt=True)
    wantf = np.stack([qo.kappa_from_map("EB", None, e.astype(np.float64), b.astype(np.float64), returnFt=True) for e, b in zip(E, B)])
    assert kf.dtype == np.complex64 and relerr(kf, wantf) < TOL32


def test_qe_recovers_input_kappa_statistically(theory):
    """tutorials/tt_verification.ipynb:597-617 on the device: unit response to first-order lensing."""
    from orphics_b200 import maps, lensing, stats
    from test_oracle_qe import FirstOrderTheory, lensed_first_order
    npix, res = 256, 1.5
    shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
    fc = maps.FourierCalc(shape, wcs, max_batch=16)
    g = fc.geometry
    modl = g.modlmap()
    LY, LX = np.meshgrid(g.ly, g.lx, indexing="ij")
    ells = np.arange(0, modl.max() + 1, 1.)
    nlev = (1.0 * np.pi / 180 / 60) ** 2
    mk = lambda ps: maps.MapGen(shape, wcs, ps[None, None], noise="philox", max_batch=16)
    mgT, mgK, mgN = mk(theory.uCl("TT", ells)), mk(theory.gCl("kk", ells)), mk(np.zeros(ells.size) + nlev)
    beam = maps.gauss_beam(modl, 1.5)
    tmask = maps.mask_kspace(shape, wcs, lmin=300, lmax=3000)
    kmask = maps.mask_kspace(shape, wcs, lmin=100, lmax=3000)
    q = lensing.qest(shape, wcs, FirstOrderTheory(theory), noise2d=np.zeros(shape) + nlev, beam2d=beam, kmask=tmask,
                     kmask_K=kmask, max_batch=16)
    b = stats.bin2D(modl, np.linspace(100, 3000, 8), geometry=g)
    nsim = 48
    rs = []
    for s0 in range(0, nsim, 16):
        seeds = list(range(s0, s0 + 16))
        Ts, Ks, Ns = mgT.get_maps(seeds), mgK.get_maps([s + 5000 for s in seeds]), mgN.get_maps([s + 9000 for s in seeds])
        obs = np.stack([np.fft.ifft2(np.fft.fft2(lensed_first_order(np.asarray(T), np.asarray(K), modl, LY, LX)) * beam).real + np.asarray(N)
                        for T, K, N in zip(Ts, Ks, Ns)])
        rec = q.kappa_from_maps("TT", obs)
        pc = fc.binned_power_batch(b, rec, np.asarray(Ks))[:, 0]
        pi = fc.binned_power_batch(b, np.asarray(Ks))[:, 0]
        rs.append(pc / pi)
    rs = np.concatenate(rs)
    mean, err = rs.mean(0), rs.std(0) / np.sqrt(nsim)
    assert np.all(np.abs(mean - 1) < 4 * err + 0.03), (mean, err)


@pytest.mark.parametrize("pol", [False, True])
def test_flat_lensing_sims_match_oracle(pol, theory):
    """lensing.FlatLensingSims.get_sim (lensing.py:499-521) with the FFT-based flat_taylens
    (lensing.py:395-440) as the lensing step: every intermediate matches the oracle on identical seeds."""
    from orphics_b200 import maps, lensing, cosmology
    from oracle import lensing_np
    shape, wcs = maps.rect_geometry(width_arcmin=128 * 2.0, px_res_arcmin=2.0)
    so, wo = omaps.rect_geometry(width_arcmin=128 * 2.0, px_res_arcmin=2.0)
    sims = lensing.FlatLensingSims(shape, wcs, cosmology.default_theory(), 1.5, 1.0, pol=pol)
    osims = lensing_np.FlatLensingSims(so, wo, theory, 1.5, 1.0, pol=pol)
    got = sims.get_sim(seed_cmb=1, seed_kappa=2, seed_noise=3, lens_order=4, return_intermediate=True)
    want = osims.get_sim(1, 2, 3, lens_order=4)
    for name, a, b in zip(("unlensed", "kappa", "lensed", "beamed", "noise", "observed"), got, want):
        assert np.shape(a) == np.shape(b), name
        assert relerr(a, b) < TOL64, (name, relerr(a, b))
    phi = lensing.kappa_to_phi(got[1], sims.modlmap)
    assert relerr(phi, lensing_np.kappa_to_phi(want[1], osims.modlmap)) < TOL64
    lens = lensing.flat_taylens(maps.ndmap(np.asarray(want[1]) * 0 + np.asarray(phi), wcs), maps.ndmap(np.asarray(want[0]).reshape((-1, 128, 128))[0], wcs), 5)
    olens = lensing_np.flat_taylens(oenmap.ndmap(np.asarray(phi), wo), oenmap.ndmap(np.asarray(want[0]).reshape((-1, 128, 128))[0], wo), 5)
    assert relerr(lens, olens) < TOL64                          # identical inputs: the transform chain itself
    skipped = sims.get_sim(seed_cmb=1, seed_noise=3, skip_lensing=True, cfrac=0.5)
    assert skipped.shape[-2:] == (64, 64)
    # deflection field and the bicubic displacement (the stand-in for pixell's displace_map) against their numpy
    # definitions, on identical inputs
    ophi = oenmap.ndmap(np.asarray(phi), wo)
    alpha = lensing.alpha_from_kappa(phi=maps.ndmap(np.asarray(phi), wcs))
    assert relerr(alpha, lensing_np.grad_phi(ophi)) < TOL64
    cmb = np.asarray(want[0])
    disp = lensing.displace_map(maps.ndmap(cmb, wcs), maps.ndmap(np.asarray(phi), wcs), order=3)
    odisp = lensing_np.displace_bicubic(cmb, ophi)
    assert np.shape(disp) == np.shape(cmb) and relerr(disp, odisp) < TOL64
    # the two lensing operations agree where they should: to the interpolation error of a cubic on these maps
    tay = lensing.flat_taylens(maps.ndmap(np.asarray(phi), wcs), maps.ndmap(cmb, wcs), 5)
    assert relerr(disp, tay) < 5e-2
    bsims = lensing.FlatLensingSims(shape, wcs, cosmology.default_theory(), 1.5, 1.0, pol=pol, lensing="bicubic")
    bobs = bsims.get_sim(seed_cmb=1, seed_kappa=2, seed_noise=3)
    assert relerr(bobs, got[5]) < 5e-2


def test_tma_c2r_pass_of_the_estimator_matches_the_one_tile_kernel(monkeypatch):
    """The estimator's c2r row pass at nx = 4096 runs on the persistent TMA kernel with 2-row tiles (ox_row_tma.cuh).  On
    512 x 4096 x 8 realisations (41 tiles per CTA) the first version returned sporadic wrong realisations: a group that ran
    ahead took a slot's still-pending previous fill for its own (mbarrier waits only know a phase's parity); the per-slot
    fill counters make that impossible.  Same bits as the one-tile kernel, several repetitions."""
    from orphics_b200 import maps, lensing, cosmology
    ny, nx, nb = 512, 4096, 8
    shape, wcs = maps.rect_geometry(width_arcmin=nx * 0.5, px_res_arcmin=0.5, height_arcmin=ny * 0.5)
    assert tuple(shape) == (ny, nx)
    modl = maps.Geometry.get(shape, wcs).modlmap()
    q = lensing.qest(shape, wcs, cosmology.default_theory(), noise2d=np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2,
                     beam2d=maps.gauss_beam(modl, 1.5), kmask=maps.mask_kspace(shape, wcs, lmin=300, lmax=2000),
                     kmask_K=maps.mask_kspace(shape, wcs, lmin=20, lmax=3500), unlensed_equals_lensed=True, max_batch=nb)
    assert q.path("TT") == "fused"
    T = np.random.RandomState(1).standard_normal((nb,) + tuple(shape)) * 50
    monkeypatch.setenv("ORPHX_KB", "legacy")
    want = np.asarray(q.kappa_from_maps("TT", T, returnFt=True))
    monkeypatch.setenv("ORPHX_KB", "tma")
    for _ in range(4):
        assert np.array_equal(np.asarray(q.kappa_from_maps("TT", T, returnFt=True)), want)


def test_hermitian_check_is_exhaustive(theory):
    """A k-map that is Hermitian everywhere except in ONE pixel pair (on a row that a sampled check would skip) must
    not take the half-plane path: the result has to be the reference's full-plane chain on that input."""
    shape, wcs, so, wo, q, qo = setup(512, 2.0, theory, True)
    assert q.path("TT") == "fused"
    qo.N.AL["TT"] = np.asarray(q.N.AL["TT"])
    rng = np.random.RandomState(9)
    kT = np.fft.fft2(rng.standard_normal(shape) * 50)
    iy, ix = 37, 11                                   # 37 is not a multiple of 16; |l| ~ 800 is inside the TT filter
    assert 300 < np.hypot(q.geometry.ly[iy], q.geometry.lx[ix]) < 2000
    kT[iy, ix] *= 3.0                                 # breaks k(p') = conj k(p) for this pair only
    got = q.kappa_from_map("TT", kT, alreadyFTed=True, returnFt=True)
    want = qo.kappa_from_map("TT", kT, alreadyFTed=True, returnFt=True)
    assert relerr(got, want) < TOL64
    # and the symmetric input still takes the fast path with the same answer
    kS = np.fft.fft2(np.fft.ifft2(kT).real)
    assert relerr(q.kappa_from_map("TT", kS, alreadyFTed=True, returnFt=True),
                  qo.kappa_from_map("TT", kS, alreadyFTed=True, returnFt=True)) < TOL64
