"""CPU tests: the oracle against the reference-made golden vectors and the reference's
known answers; host logic of the product that needs no device."""
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import enmap_np as oenmap, maps_np as omaps, stats_np as ostats, theory as otheory, philox_np


@pytest.mark.parametrize("case", ["fourier", "trimquirk", "onedge", "odd"])
def test_bin2d_oracle_matches_reference_golden(case):
    z = load_golden(f"bin2d_{case}.npz")
    b = ostats.bin2D(z["modrmap"], z["edges"])
    assert b.digitized.dtype == np.int64
    assert np.array_equal(b.digitized, z["digitized"])
    assert np.array_equal(b.centers, z["centers"])
    with np.errstate(all="ignore"):
        c, r, n = b.bin(z["data"], get_count=True)
        assert np.array_equal(n, z["count"]) and n.dtype == z["count"].dtype
        assert np.array_equal(r, z["res"], equal_nan=True)
        c, r, n = b.bin(z["data"], weights=z["weights"], get_count=True)
        assert np.array_equal(n, z["count_w"]) and np.array_equal(r, z["res_w"], equal_nan=True)
        if "data_nan" in z:
            c, r, n = b.bin(z["data_nan"], mask_nan=True, get_count=True)
            assert np.array_equal(n, z["count_nan"]) and np.array_equal(r, z["res_nan"], equal_nan=True)
        if "res_err" in z:
            c, r, s = b.bin(z["data"], err=True)
            assert np.array_equal(r, z["res_err"], equal_nan=True)
            assert np.array_equal(s, z["std_err"], equal_nan=True)


def test_trim_quirk_is_reproduced():
    # stats.py:796-797: with no pixel above the last edge the last real bin is dropped
    z = load_golden("bin2d_trimquirk.npz")
    assert z["res"].size < z["centers"].size
    assert z["res"].size == int(z["digitized"].max()) + 1 - 2


@pytest.mark.parametrize("P", [1, 2, 3])
def test_statistics_triple_matches_reference_golden(P):
    z = load_golden(f"statistics_P{P}.npz")
    parts = []
    for r in range(P):
        s = ostats.StatsTriple()
        for x in z[f"x{r}"]:
            s.add("v", x)
        parts.append(s)
    tot = ostats.StatsTriple.merge(parts)
    assert tot.N["v"] == int(z["N"])
    np.testing.assert_allclose(tot.mean("v"), z["mean"], rtol=1e-14)
    np.testing.assert_allclose(tot.cov("v"), z["cov"], rtol=1e-12, atol=1e-14)


def test_mpi_distribute_matches_reference_golden():
    z = load_golden("mpi_distribute.npz")
    from orphics_b200 import mpi as pmpi
    for n, c in z["cases"]:
        for impl in (ostats.mpi_distribute, pmpi.mpi_distribute):
            num, tasks = impl(int(n), int(c))
            assert np.array_equal(num, z[f"n_{n}_{c}"])
            assert np.array_equal([t[0] for t in tasks], z[f"first_{n}_{c}"])
            assert sum(len(t) for t in tasks) == n


def test_rect_geometry_known_answers():
    # tutorials/demo-grf.ipynb:97 ; mapwork.ipynb:61 ; "Correlated maps.ipynb":104
    shape, wcs = omaps.rect_geometry(width_deg=10.0, px_res_arcmin=1.0)
    assert shape == (600, 600)
    np.testing.assert_allclose(wcs.cdelt, [1 / 60.0, 1 / 60.0], rtol=1e-12)
    np.testing.assert_allclose(wcs.crval, [0, 0], atol=1e-12)
    np.testing.assert_allclose(wcs.crpix, [300.5, 300.5], rtol=1e-12)
    assert omaps.rect_geometry(width_deg=20.0, px_res_arcmin=0.5)[0] == (2400, 2400)
    assert omaps.rect_geometry(width_deg=25.0, px_res_arcmin=2.0)[0] == (750, 750)
    assert omaps.rect_geometry(width_arcmin=2048 * 0.5, px_res_arcmin=0.5, pol=True)[0] == (3, 2048, 2048)


def test_product_host_geometry_matches_oracle():
    from orphics_b200 import enmap as penmap
    for w, r in ((10 * 60.0, 1.0), (512 * 2.0, 2.0), (300.0, 0.5)):
        so, wo = omaps.rect_geometry(width_arcmin=w, px_res_arcmin=r)
        am = np.pi / 180 / 60
        sp, wp = penmap.geometry([[-w / 2 * am, -w / 2 * am], [w / 2 * am, w / 2 * am]], r * am)
        assert so == sp
        np.testing.assert_array_equal(wo.cdelt, wp.cdelt)
        np.testing.assert_array_equal(wo.crpix, wp.crpix)
        for m in ("cylindrical", "intermediate"):
            np.testing.assert_array_equal(oenmap.extent(so, wo, method=m), penmap.extent(sp, wp, method=m))
            assert oenmap.area(so, wo, m) == penmap.area(sp, wp, m)
            for a, b in zip(oenmap.laxes(so, wo, m), penmap.laxes(sp, wp, m)):
                np.testing.assert_array_equal(a, b)


def test_theory_loader_product_matches_oracle_and_file():
    from orphics_b200 import cosmology as pcos
    to, tp = otheory.load_theory(), pcos.default_theory()
    ells = np.arange(0, 9500, 0.5)
    for k in ("TT", "EE", "BB", "TE"):
        np.testing.assert_array_equal(to.lCl(k, ells), tp.lCl(k, ells))
        np.testing.assert_array_equal(to.uCl(k, ells), tp.uCl(k, ells))
    np.testing.assert_array_equal(to.gCl("kk", ells), tp.gCl("kk", ells))
    # first row of data/cosmo2017_10K_acc3_lensedCls.dat: L=2 TT=0.10580E+04
    assert to.lCl("TT", 2.0) == pytest.approx(0.10580e4 * 2 * np.pi / 6.0, rel=1e-12)
    assert to.lCl("TT", 1.0) == 0.0 and to.lCl("TT", 9001.0) == 0.0
    root = "/root/reference/data/cosmo2017_10K_acc3"
    if os.path.exists(root + "_lensedCls.dat"):
        tf = otheory.load_theory(root)
        for k in ("TT", "EE", "BB", "TE"):
            np.testing.assert_array_equal(tf.lCl(k, ells), to.lCl(k, ells))
            np.testing.assert_array_equal(tf.uCl(k, ells), to.uCl(k, ells))
        np.testing.assert_array_equal(tf.gCl("kk", ells), to.gCl("kk", ells))


def test_oracle_sim_power_bin_recovers_theory(theory):
    # the notebooks' acceptance test (tutorials/demo-grf.ipynb:159-161): binned / theory -> 1
    shape, wcs = omaps.rect_geometry(width_arcmin=256 * 2.0, px_res_arcmin=2.0)
    modl = oenmap.modlmap(shape, wcs)
    ells = np.arange(0, modl.max() + 1, 1.0)
    ps = otheory.power_from_theory(ells, theory, lensed=True, pol=False)
    mg = omaps.MapGen(shape, wcs, ps)
    fc = omaps.FourierCalc(shape, wcs)
    b = ostats.bin2D(modl, np.arange(200, 3000, 80.0))
    rs = []
    for i in range(24):
        p2d, _, _ = fc.power2d(mg.get_map(seed=1000 + i))
        c, p1 = b.bin(p2d)
        rs.append(p1 / theory.lCl("TT", c))
    rs = np.array(rs)
    assert abs(rs.mean() - 1) < 0.02


def test_oracle_pol_roundtrip_and_parseval(theory):
    shape, wcs = omaps.rect_geometry(width_arcmin=64 * 2.0, px_res_arcmin=2.0, pol=True)
    rng = np.random.RandomState(3)
    m = oenmap.ndmap(rng.standard_normal(shape), wcs)
    fc = omaps.FourierCalc(shape, wcs)
    k = fc.iqu2teb(m, normalize=True)
    back = oenmap.harm2map(k)
    np.testing.assert_allclose(back, m, atol=1e-12)      # QU -> EB -> QU identity
    np.testing.assert_allclose(np.sum(np.abs(k) ** 2), np.sum(m ** 2), rtol=1e-12)  # Parseval (unitary)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = tuple(int(x) for x in philox_np.philox4x32_10(*c, *k))
        assert got == want
    f = philox_np.noise_field(7, 2, 48, 40, hermitian=True)
    iy, ix = np.mgrid[0:48, 0:40]
    assert np.array_equal(f[:, (-iy) % 48, (-ix) % 40], np.conj(f))
    g = philox_np.noise_field(7, 1, 128, 128)
    assert abs(g.real.std() - 1) < 0.02 and abs(g.imag.std() - 1) < 0.02


def test_cosine_window_product_matches_oracle():
    from orphics_b200 import maps as pmaps
    for args in ((64, 48, 10, 7, 0, 0), (50, 50, 6, 6, 2, 3), (32, 40, 0, 5, 1, 0)):
        np.testing.assert_array_equal(omaps.cosine_window(*args), pmaps.cosine_window(*args))
    np.testing.assert_array_equal(omaps.gauss_beam(np.arange(0, 5000.0, 7), 1.5), pmaps.gauss_beam(np.arange(0, 5000.0, 7), 1.5))
