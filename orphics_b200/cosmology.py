"""Theory-spectrum inputs of the hot path: the part of orphics.cosmology that turns the
CAMB tables in data/ into C_l arrays (loadTheorySpectraFromCAMB cosmology.py:863-946,
default_theory :850-852, power_from_theory :1270-1280).  Host set-up code (1-D arrays)."""
import os

import numpy as np

_PACKED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cosmo2017_10K_acc3.npz")


class TheorySpectra:
    """Linear interpolators of tabulated spectra, zero outside the table (the behaviour of
    pyfisher.TheorySpectra that cosmology.py:888-943 relies on)."""

    def __init__(self):
        self._tables = {"l": {}, "u": {}, "g": {}}
        self.dimensionless = False

    def loadCls(self, ell, Cl, XY="TT", lensed=False, interporder="linear", lpad=9000, fill_zero=True):
        keep = ell < lpad
        self._tables["l" if lensed else "u"][XY] = (np.array(ell[keep], dtype=float), np.array(Cl[keep], dtype=float))

    def loadGenericCls(self, ell, Cl, keyName, lpad=9000):
        keep = ell < lpad
        self._tables["g"][keyName] = (np.array(ell[keep], dtype=float), np.array(Cl[keep], dtype=float))

    def _get(self, which, XY, ell):
        ell = np.asarray(ell, dtype=float)
        if XY == "ET":
            XY = "TE"
        if XY in ("EB", "BE", "TB", "BT"):
            return ell * 0.
        x, y = self._tables[which][XY]
        return np.interp(ell, x, y, left=0., right=0.)

    def lCl(self, XY, ell):
        return self._get("l", XY, ell)

    def uCl(self, XY, ell):
        return self._get("u", XY, ell)

    def gCl(self, keyName, ell):
        return self._get("g", keyName, ell)


def _read_tables(cambRoot):
    if cambRoot is None:
        z = np.load(_PACKED)
        return z["lensedCls"].T, z["scalCls"].T, z["lenspotentialCls"].T
    return (np.loadtxt(cambRoot + "_lensedCls.dat", unpack=True, usecols=[0, 1, 2, 3, 4]),
            np.loadtxt(cambRoot + "_scalCls.dat", unpack=True, usecols=[0, 1, 2, 3]),
            np.loadtxt(cambRoot + "_lenspotentialCls.dat", unpack=True, usecols=[0, 5]))


def loadTheorySpectraFromCAMB(cambRoot=None, unlensedEqualsLensed=False, useTotal=False, TCMB=2.7255e6, lpad=9000,
                              get_dimensionless=True, skip_lens=False, dells=False, scalcls=True):
    """D_l tables -> C_l interpolators.  cambRoot=None uses the packed copy of
    data/cosmo2017_10K_acc3_* shipped with the package."""
    if useTotal or not scalcls:
        raise NotImplementedError("only the scalCls/lensedCls tables are supported")
    if not (get_dimensionless):
        TCMB = 1.
    lens, scal, pot = _read_tables(cambRoot)
    theory = TheorySpectra()
    ell = np.array(lens[0])
    mult = (2. * np.pi / ell / (ell + 1.) if not (dells) else 1) / TCMB ** 2.
    lensed = {"TT": lens[1] * mult, "EE": lens[2] * mult, "BB": lens[3] * mult, "TE": lens[4] * mult}
    for k in ("TT", "TE", "EE", "BB"):
        theory.loadCls(ell, lensed[k], k, lensed=True, lpad=lpad)
    if not (skip_lens):
        theory.loadGenericCls(np.array(pot[0]), 2. * np.pi * np.array(pot[1]) / 4., "kk", lpad=lpad)
    if unlensedEqualsLensed:
        for k in ("TT", "TE", "EE", "BB"):
            theory.loadCls(ell, lensed[k], k, lensed=False, lpad=lpad)
    else:
        ell = np.array(scal[0])
        mult = (2. * np.pi / ell / (ell + 1.) if not (dells) else 1) / TCMB ** 2.
        un = {"TT": scal[1] * mult, "EE": scal[2] * mult, "TE": scal[3] * mult, "BB": scal[2] * 0.}
        for k in ("TT", "TE", "EE", "BB"):
            theory.loadCls(ell, un[k], k, lensed=False, lpad=lpad)
    theory.dimensionless = get_dimensionless
    return theory


def default_theory(lpad=9000, root=None):
    return loadTheorySpectraFromCAMB(root, unlensedEqualsLensed=False, useTotal=False, TCMB=2.7255e6, lpad=lpad,
                                     get_dimensionless=False)


def power_from_theory(ells, theory, lensed=True, pol=False):
    ncomp = 3 if pol else 1
    cfunc = theory.lCl if lensed else theory.uCl
    ps = np.zeros((ncomp, ncomp,) + np.shape(ells))
    ps[0, 0] = cfunc('TT', ells)
    if pol:
        ps[1, 1] = cfunc('EE', ells)
        ps[2, 2] = cfunc('BB', ells)
        ps[0, 1] = cfunc('TE', ells)
        ps[1, 0] = cfunc('TE', ells)
    return ps
