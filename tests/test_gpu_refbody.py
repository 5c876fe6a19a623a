"""GPU parity against golden vectors produced by the reference's OWN MapGen / FourierCalc / helper bodies
(tests/golden/make_golden_maps.py; the pixell layer underneath is the oracle's stand-in): maps on numpy seeds,
k-maps, 2-D spectra incl. the TEB matrix layout, f1power, ifft, taper, masks, beam, filter_map -- 1e-10, masks
and index-like outputs bit-exact."""
import os

import numpy as np
import pytest

from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "maps_refbody.npz"))
TOL = 1e-10


@pytest.mark.parametrize("pol", [False, True])
def test_mapgen_fouriercalc_match_reference_bodies(pol):
    from orphics_b200 import maps
    tag = "IQU" if pol else "T"
    shape, wcs = maps.rect_geometry(width_arcmin=96 * 2.0, px_res_arcmin=2.0, height_arcmin=64 * 2.0, pol=pol)
    assert tuple(shape[-2:]) == (64, 96)
    mg0 = maps.MapGen(shape, wcs, G[f"{tag}_ps"])                      # set-up from the 1-D spectra
    assert relerr(mg0.covsqrt, G[f"{tag}_covsqrt"]) < 5e-9             # (FFT-library conditioning of spec2flat's smoothing)
    mg = maps.MapGen(shape, wcs, covsqrt=G[f"{tag}_covsqrt"])          # identical set-up input from here on
    fc = maps.FourierCalc(shape, wcs)
    m1, m2 = mg.get_map(seed=11), mg.get_map(seed=12)
    assert relerr(m1, G[f"{tag}_map11"]) < TOL and relerr(m2, G[f"{tag}_map12"]) < TOL
    assert relerr(mg.get_map(seed=11, harm=True), G[f"{tag}_harm11"]) < TOL
    p2d, k1, _ = fc.power2d(G[f"{tag}_map11"])
    assert relerr(p2d, G[f"{tag}_p2d"]) < TOL and relerr(k1, G[f"{tag}_k1"]) < TOL
    assert relerr(fc.power2d(G[f"{tag}_map11"], G[f"{tag}_map12"])[0], G[f"{tag}_p2d_cross"]) < TOL
    assert relerr(fc.power2d(G[f"{tag}_map11"], pixel_units=True)[0], G[f"{tag}_p2d_pix"]) < TOL
    if pol:
        assert relerr(mg.get_map(seed=11, scalar=True), G["IQU_map11_scalar"]) < TOL
        assert relerr(mg.get_map(seed=11, iau=True), G["IQU_map11_iau"]) < TOL
        skip = fc.power2d(G["IQU_map11"], skip_cross=True)[0]
        assert relerr(skip, G["IQU_p2d_skip"]) < TOL and not skip[0, 1].any()
        assert relerr(fc.power2d(G["IQU_map11"], rot=False)[0], G["IQU_p2d_norot"]) < TOL
        assert relerr(fc.iqu2teb(G["IQU_map11"]), G["IQU_teb_unitary"]) < TOL
    else:
        assert relerr(fc.f1power(G["T_map12"], G["T_k1"])[0], G["T_f1power"]) < TOL
        assert relerr(fc.ifft(G["T_k1"]), G["T_ifft"]) < TOL and relerr(fc.fft(G["T_map11"]), G["T_fft"]) < TOL
        assert abs(fc.normfact / float(G["T_normfact"]) - 1) < 1e-14


def test_helpers_match_reference_bodies():
    from orphics_b200 import maps
    shape, wcs = maps.rect_geometry(width_arcmin=96 * 2.0, px_res_arcmin=2.0, height_arcmin=64 * 2.0)
    modl = maps.Geometry.get(shape, wcs).modlmap()
    taper, w2 = maps.get_taper(shape, wcs)
    assert relerr(taper, G["T_taper"]) < 1e-15 and abs(w2 / float(G["T_w2"]) - 1) < 1e-14
    assert relerr(maps.get_taper(shape, wcs, taper_percent=20.0, pad_percent=5.0, weight=np.abs(G["T_map11"]))[0], G["T_taper_weight"]) < 1e-15
    assert relerr(maps.cosine_window(37, 51, lenApodY=5, lenApodX=9, padY=2, padX=0), G["T_window_odd"]) < 1e-15
    assert relerr(maps.gauss_beam(modl, 1.5), G["T_beam"]) < 1e-14
    assert np.array_equal(np.asarray(maps.mask_kspace(shape, wcs, lmin=300, lmax=2000)), G["T_mask_l"])
    assert np.array_equal(np.asarray(maps.mask_kspace(shape, wcs, lxcut=90, lycut=50, lmax=4000)), G["T_mask_xy"])
    assert relerr(maps.filter_map(maps.ndmap(G["T_map11"], wcs), G["T_beam"] * G["T_mask_l"]), G["T_filtered"]) < TOL


def test_lensing_callers_match_reference_bodies():
    """lensing.flat_taylens / kappa_to_phi / fkappa_to_fphi on device FFTs vs the reference's own bodies."""
    from orphics_b200 import maps, lensing
    GL = np.load(os.path.join(ROOT, "tests", "golden", "lensing_refbody.npz"))
    shape, wcs = maps.rect_geometry(width_arcmin=64 * 2.0, px_res_arcmin=2.0, height_arcmin=48 * 2.0)
    assert tuple(shape) == (48, 64)
    modl = maps.Geometry.get(shape, wcs).modlmap()
    phi, fphi = lensing.kappa_to_phi(maps.ndmap(GL["kappa"], wcs), modl, return_fphi=True)
    assert relerr(phi, GL["phi"]) < TOL
    assert relerr(fphi, GL["fphi"]) < TOL      # carries enmap.fft(normalize='phys')'s (pixsize/Npix)^1/2
    # the device route (half-plane transforms, ox_lens_kappa_to_phi), from a host and from a device-resident map
    assert relerr(lensing.kappa_to_phi(maps.ndmap(GL["kappa"], wcs), modl), GL["phi"]) < TOL
    assert relerr(lensing.kappa_to_phi(maps.devmap.from_host(GL["kappa"], wcs), modl), GL["phi"]) < TOL
    assert relerr(lensing.fkappa_to_fphi(np.fft.fft2(GL["kappa"]), modl), GL["fk2fp"]) < 1e-13
    for order in (2, 5):
        got = lensing.flat_taylens(maps.ndmap(GL["phis"], wcs), maps.ndmap(GL["imap"], wcs), order)
        assert relerr(got, GL[f"lensed_o{order}"]) < TOL
        got = lensing.flat_taylens(maps.devmap.from_host(GL["phis"], wcs), maps.devmap.from_host(GL["imap"], wcs), order)
        assert relerr(got, GL[f"lensed_o{order}"]) < TOL
