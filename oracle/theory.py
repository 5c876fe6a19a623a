"""ORACLE (test infrastructure) -- restatement of the theory-spectra loader.

Follows /root/reference/orphics/cosmology.py:863-946 (loadTheorySpectraFromCAMB),
:850-852 (default_theory) and :1270-1280 (power_from_theory).  The interpolator
class lives in the third-party ``pyfisher`` (cosmology.py:888; absent, unpinned):
``TheorySpectra.loadCls(ell,Cl,XY,lensed,interporder="linear",lpad)`` keeps
ell<lpad and evaluates with interp1d(bounds_error=False, fill_value=0) --
restated as np.interp(left=0,right=0).  PARITY UNPINNED for that class; the
arithmetic D_l -> C_l is pinned by the reference lines cited.
"""
import os

import numpy as np

_PACKED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "orphics_b200", "data",
                       "cosmo2017_10K_acc3.npz")


class TheorySpectra:
    def __init__(self):
        self._l = {}
        self._u = {}
        self._g = {}

    @staticmethod
    def _key(XY):
        return "TE" if XY == "ET" else XY

    def loadCls(self, ell, Cl, XY, lensed=False, lpad=9000):
        sel = ell < lpad
        (self._l if lensed else self._u)[XY] = (np.array(ell[sel]), np.array(Cl[sel]))

    def loadGenericCls(self, ell, Cl, key, lpad=9000):
        sel = ell < lpad
        self._g[key] = (np.array(ell[sel]), np.array(Cl[sel]))

    def _eval(self, table, XY, ell):
        XY = self._key(XY)
        ell = np.asarray(ell, dtype=np.float64)
        if XY in ("EB", "BE", "TB", "BT"):
            return ell * 0.0
        x, y = table[XY]
        return np.interp(ell, x, y, left=0.0, right=0.0)

    def lCl(self, XY, ell):
        return self._eval(self._l, XY, ell)

    def uCl(self, XY, ell):
        return self._eval(self._u, XY, ell)

    def gCl(self, key, ell):
        return self._eval(self._g, key, ell)


def _columns(cambRoot):
    if cambRoot is None:
        z = np.load(_PACKED)
        return z["lensedCls"].T, z["scalCls"].T, z["lenspotentialCls"].T
    lens = np.loadtxt(cambRoot + "_lensedCls.dat", unpack=True, usecols=[0, 1, 2, 3, 4])
    scal = np.loadtxt(cambRoot + "_scalCls.dat", unpack=True, usecols=[0, 1, 2, 3])
    pot = np.loadtxt(cambRoot + "_lenspotentialCls.dat", unpack=True, usecols=[0, 5])
    return lens, scal, pot


def load_theory(cambRoot=None, lpad=9000, unlensedEqualsLensed=False):
    """cosmology.py:863-946 with get_dimensionless=False (TCMB=1, :875),
    useTotal=False, scalcls=True -- i.e. default_theory() (:850-852).
    cambRoot=None reads the packed raw columns made by tools/pack_camb.py."""
    lens, scal, pot = _columns(cambRoot)
    th = TheorySpectra()
    ell, tt, ee, bb, te = [np.array(c) for c in lens]
    mult = 2.0 * np.pi / ell / (ell + 1.0)                              # :892-893
    for name, c in (("TT", tt), ("TE", te), ("EE", ee), ("BB", bb)):
        th.loadCls(ell, c * mult, name, lensed=True, lpad=lpad)         # :894-901
    elldd, cldd = pot
    th.loadGenericCls(elldd, 2.0 * np.pi * cldd / 4.0, "kk", lpad=lpad)  # :905-907,:912
    if unlensedEqualsLensed:                                            # :915-920
        for name, c in (("TT", tt), ("TE", te), ("EE", ee), ("BB", bb)):
            th.loadCls(ell, c * mult, name, lensed=False, lpad=lpad)
    else:                                                               # :923-930
        ell, tt, ee, te = [np.array(c) for c in scal]
        mult = 2.0 * np.pi / ell / (ell + 1.0)
        for name, c in (("TT", tt), ("TE", te), ("EE", ee), ("BB", ee * 0.0)):
            th.loadCls(ell, c * mult, name, lensed=False, lpad=lpad)
    return th


def power_from_theory(ells, theory, lensed=True, pol=False):
    """cosmology.py:1270-1280."""
    ells = np.asarray(ells)
    ncomp = 3 if pol else 1
    cfunc = theory.lCl if lensed else theory.uCl
    ps = np.zeros((ncomp, ncomp) + ells.shape)
    ps[0, 0] = cfunc("TT", ells)
    if pol:
        ps[1, 1] = cfunc("EE", ells)
        ps[2, 2] = cfunc("BB", ells)
        ps[0, 1] = cfunc("TE", ells)
        ps[1, 0] = cfunc("TE", ells)
    return ps
