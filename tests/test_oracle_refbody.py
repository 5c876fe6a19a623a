"""CPU: the oracle's restatement of the orphics-owned logic (MapGen, FourierCalc, tapers, masks, beam, filter_map)
against golden vectors produced by the reference's OWN class and function bodies executed over the same
pixell stand-in (tests/golden/make_golden_maps.py).  Agreement here means the restatement adds no
discrepancy of its own on top of the (unpinnable) pixell layer."""
import os

import numpy as np
import pytest

from conftest import ROOT, relerr
from oracle import enmap_np as oenmap, maps_np as omaps

G = np.load(os.path.join(ROOT, "tests", "golden", "maps_refbody.npz"))
TOL = 1e-13


def geom(pol):
    shape, wcs = omaps.rect_geometry(width_arcmin=96 * 2.0, px_res_arcmin=2.0, height_arcmin=64 * 2.0, pol=pol)
    assert tuple(shape[-2:]) == (64, 96)
    return shape, wcs


@pytest.mark.parametrize("pol", [False, True])
def test_mapgen_and_fouriercalc_restatement(pol):
    tag = "IQU" if pol else "T"
    shape, wcs = geom(pol)
    mg = omaps.MapGen(shape, wcs, G[f"{tag}_ps"])
    fc = omaps.FourierCalc(shape, wcs)
    assert relerr(mg.covsqrt, G[f"{tag}_covsqrt"]) < TOL
    m1, m2 = mg.get_map(seed=11), mg.get_map(seed=12)
    assert relerr(m1, G[f"{tag}_map11"]) < TOL and relerr(m2, G[f"{tag}_map12"]) < TOL
    assert relerr(mg.get_map(seed=11, harm=True), G[f"{tag}_harm11"]) < TOL
    p2d, k1, _ = fc.power2d(m1)
    assert relerr(p2d, G[f"{tag}_p2d"]) < TOL and relerr(k1, G[f"{tag}_k1"]) < TOL
    assert relerr(fc.power2d(m1, m2)[0], G[f"{tag}_p2d_cross"]) < TOL
    assert relerr(fc.power2d(m1, pixel_units=True)[0], G[f"{tag}_p2d_pix"]) < TOL
    if pol:
        assert relerr(mg.get_map(seed=11, scalar=True), G["IQU_map11_scalar"]) < TOL
        assert relerr(mg.get_map(seed=11, iau=True), G["IQU_map11_iau"]) < TOL
        assert relerr(fc.power2d(m1, skip_cross=True)[0], G["IQU_p2d_skip"]) < TOL
        assert relerr(fc.power2d(m1, rot=False)[0], G["IQU_p2d_norot"]) < TOL
        assert relerr(fc.iqu2teb(m1), G["IQU_teb_unitary"]) < TOL
    else:
        assert relerr(fc.f1power(m2, k1)[0], G["T_f1power"]) < TOL
        assert relerr(fc.ifft(k1), G["T_ifft"]) < TOL and relerr(fc.fft(m1), G["T_fft"]) < TOL
        assert abs(fc.normfact / float(G["T_normfact"]) - 1) < 1e-15


def test_helpers_restatement():
    shape, wcs = geom(False)
    modl = np.asarray(oenmap.modlmap(shape, wcs))
    taper, w2 = omaps.get_taper(shape, wcs)
    assert np.array_equal(np.asarray(taper), G["T_taper"]) and w2 == float(G["T_w2"])
    tw = omaps.get_taper(shape, wcs, taper_percent=20.0, pad_percent=5.0, weight=np.abs(G["T_map11"]))[0]
    assert relerr(tw, G["T_taper_weight"]) < 1e-15
    assert np.array_equal(omaps.cosine_window(37, 51, lenApodY=5, lenApodX=9, padY=2, padX=0), G["T_window_odd"])
    assert np.array_equal(omaps.gauss_beam(modl, 1.5), G["T_beam"])
    assert np.array_equal(np.asarray(omaps.mask_kspace(shape, wcs, lmin=300, lmax=2000)), G["T_mask_l"])
    assert np.array_equal(np.asarray(omaps.mask_kspace(shape, wcs, lxcut=90, lycut=50, lmax=4000)), G["T_mask_xy"])
    assert relerr(omaps.filter_map(oenmap.ndmap(G["T_map11"], wcs), G["T_beam"] * G["T_mask_l"]), G["T_filtered"]) < TOL


GL = np.load(os.path.join(ROOT, "tests", "golden", "lensing_refbody.npz"))


def test_lensing_callers_restatement():
    """oracle/lensing_np.py vs the reference's own flat_taylens / kappa_to_phi / fkappa_to_fphi bodies."""
    from oracle import lensing_np
    shape, wcs = omaps.rect_geometry(width_arcmin=64 * 2.0, px_res_arcmin=2.0, height_arcmin=48 * 2.0)
    modl = np.asarray(oenmap.modlmap(shape, wcs))
    assert relerr(lensing_np.kappa_to_phi(oenmap.ndmap(GL["kappa"], wcs), modl), GL["phi"]) < TOL
    assert relerr(lensing_np.fkappa_to_fphi(np.fft.fft2(GL["kappa"]), modl), GL["fk2fp"]) < TOL
    for order in (2, 5):
        got = lensing_np.flat_taylens(oenmap.ndmap(GL["phis"], wcs), oenmap.ndmap(GL["imap"], wcs), order)
        assert relerr(got, GL[f"lensed_o{order}"]) < 1e-12
    assert float(GL["max_shift_pixels"]) > 3      # the integer part of the Taylens displacement is exercised
