"""GPU parity: stats.bin2D through the C-ABI vs the reference-made golden vectors and the
oracle.  Slot indices and counts are bit-exact; sums within 1e-10 (the device sums in a
different but fixed order than np.bincount)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import enmap_np as oenmap, maps_np as omaps, stats_np as ostats

pytestmark = pytest.mark.gpu
RTOL = 1e-10  # north_star fp64 tolerance


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    np.testing.assert_allclose(a[m], b[m], rtol=rtol, atol=rtol * np.max(np.abs(b[m])) if m.any() else 0)


@pytest.mark.parametrize("case", ["fourier", "trimquirk", "onedge", "odd"])
def test_bin2d_matches_reference_golden(case):
    from orphics_b200 import stats
    z = load_golden(f"bin2d_{case}.npz")
    b = stats.bin2D(z["modrmap"], z["edges"])
    assert b.digitized.dtype == np.int64
    assert np.array_equal(b.digitized, z["digitized"])          # bit-exact indices
    assert np.array_equal(b.centers, z["centers"])
    with np.errstate(all="ignore"):
        c, r, n = b.bin(z["data"], get_count=True)
    assert n.dtype == np.int64 and np.array_equal(n, z["count"])  # bit-exact counts
    close(r, z["res"])
    c, r, n = b.bin(z["data"], weights=z["weights"], get_count=True)
    close(n, z["count_w"])
    close(r, z["res_w"])
    if "data_nan" in z:
        c, r, n = b.bin(z["data_nan"], mask_nan=True, get_count=True)
        assert np.array_equal(n, z["count_nan"])
        close(r, z["res_nan"])
    if "res_err" in z:
        c, r, s = b.bin(z["data"], err=True)
        close(r, z["res_err"])
        close(s, z["std_err"], rtol=1e-9)
    c2, r2 = b.bin(z["data"])
    assert np.array_equal(c2, z["centers"])
    close(r2, z["res"])
    # float32 data path
    c3, r3 = b.bin(z["data"].astype(np.float32))
    close(r3, ostats.bin2D(z["modrmap"], z["edges"]).bin(z["data"].astype(np.float32).astype(np.float64))[1], rtol=1e-6)


def test_bin2d_is_deterministic():
    from orphics_b200 import stats
    z = load_golden("bin2d_fourier.npz")
    b = stats.bin2D(z["modrmap"], z["edges"])
    r0 = b.bin(z["data"])[1]
    for _ in range(5):
        assert np.array_equal(b.bin(z["data"])[1], r0, equal_nan=True)


@pytest.mark.parametrize("npix,res", [(512, 2.0), (2048, 0.5)])
def test_device_modlmap_digitize_counts_bit_exact(npix, res):
    """BASELINE configs 1 and 2 geometries: modlmap, slot indices and per-bin mode counts
    identical to numpy / orphics.stats.bin2D semantics."""
    from orphics_b200 import enmap, maps, stats
    shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
    so, wo = omaps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
    assert shape == so == (npix, npix)
    modl_o = np.asarray(oenmap.modlmap(so, wo))
    g = enmap.Geometry.get(shape, wcs)
    modl = g.modlmap()
    assert np.array_equal(modl, modl_o)                           # bit-identical |l|
    for edges in (np.arange(100, 3000, 40.0), np.linspace(20, 3500, 20), np.array([0.0, modl_o[0, 1], modl_o[1, 1], 5000.0])):
        ob = ostats.bin2D(modl_o, edges)
        for b in (stats.bin2D(modl, edges), stats.bin2D(modl, edges, geometry=g)):
            assert np.array_equal(b.digitized, ob.digitized)
            assert np.array_equal(b.slot_counts, np.bincount(ob.digitized, minlength=len(edges) + 1))
    rng = np.random.RandomState(0)
    data = rng.standard_normal(shape) * (1 + modl_o)
    edges = np.arange(100, 3000, 40.0)
    ob = ostats.bin2D(modl_o, edges)
    b = stats.bin2D(modl, edges, geometry=g)
    c, r, n = b.bin(data, get_count=True)
    co, ro, no = ob.bin(data, get_count=True)
    assert np.array_equal(n, no) and np.array_equal(c, co)
    close(r, ro)
    # batched
    stack = np.stack([data, 2 * data, data ** 2])
    cb, rb = b.bin_batch(stack)
    for i in range(3):
        close(rb[i], ob.bin(stack[i])[1])


def test_mask_kspace_and_rotmat_match_oracle():
    from orphics_b200 import enmap, maps
    shape, wcs = maps.rect_geometry(width_arcmin=256 * 1.5, px_res_arcmin=1.5, pol=True)
    so, wo = omaps.rect_geometry(width_arcmin=256 * 1.5, px_res_arcmin=1.5, pol=True)
    for kw in (dict(lmin=300, lmax=2000), dict(lmin=20, lmax=3500, lxcut=50), dict(lycut=90.0), dict(lmax=1234.5, lxcut=10, lycut=20)):
        assert np.array_equal(np.asarray(maps.mask_kspace(shape, wcs, **kw)), np.asarray(omaps.mask_kspace(so, wo, **kw)))
    for iau in (False, True):
        rot = enmap.Geometry.get(shape, wcs).rotmat(iau)
        ro = oenmap.queb_rotmat(oenmap.lmap(so, wo), iau=iau)
        np.testing.assert_allclose(rot, ro, atol=1e-14)
