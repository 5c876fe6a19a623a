"""ORACLE (test infrastructure) -- numpy restatement of the historical orphics
``lensing.qest`` (Hu-Okamoto / Hu-DeDeo-Vale flat-sky quadratic estimator).

PARITY UNPINNED: the class is absent from the /root/reference snapshot.  Only its call
sites survive: tutorials/tt_verification.ipynb:81 (constructor), :608,:610
(``kappa_from_map("TT", kT, alreadyFTed=True)``, ``("EB", kT,kE,kB, alreadyFTed=True)``),
lensing.py:973-976 (``SplitLensing.qfrag``) and the conventions corroborated in-tree:
phi = 2 kappa / (L(L+1)) (lensing.py:662-665), raw fft / ifft(normalize=True)
(lensing.py:20,826,892), strict mask inequalities (maps.py:1941,1943).  The arithmetic below
follows the historical Estimator/QuadNorm classes and is validated
physically in tests/test_oracle_qe.py: A_L against a brute-force O(N^2) sum, unit response
<kappa_hat kappa>/<kappa kappa> -> 1 on gradient-order lensed sims.
"""
import numpy as np

from . import enmap_np as enmap


def _fft(a):
    return enmap.raw_fft(a)


def _ifft(a):
    return enmap.raw_ifft(a, normalize=True)


def _fmask(arr, mask):
    arr = arr.copy()
    if mask is not None:
        arr[np.asarray(mask) < 1.e-3] = 0.
    return arr


class QuadNorm:
    """Filters and normalisation on the 2-D Fourier grid."""

    def __init__(self, shape, wcs, theory, noise2d, noise2d_P, noise2d_B, beam2d, kmask, kmask_P, kmask_K,
                 grad_cut=None, unlensed_equals_lensed=False, bigell=9000, method="cylindrical"):
        self.shape, self.wcs = shape[-2:], wcs
        ly, lx = enmap.laxes(shape, wcs, method)
        self.lyMap, self.lxMap = np.meshgrid(ly, lx, indexing="ij")
        self.modLMap = np.sqrt(self.lxMap ** 2 + self.lyMap ** 2)
        self.thetaMap = np.arctan2(self.lyMap, self.lxMap)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.nan_to_num(1. / self.modLMap, posinf=0., neginf=0.)
        self.lxHatMap, self.lyHatMap = self.lxMap * inv, self.lyMap * inv
        py, px = enmap.extent(shape, wcs, method=method) / np.array(shape[-2:])
        self.pixScaleY, self.pixScaleX = py, px
        self.bigell = bigell
        self.gradCut = bigell if grad_cut is None else grad_cut
        L = self.modLMap
        self.uCl = {k: (theory.lCl(k, L) if unlensed_equals_lensed else theory.uCl(k, L)) for k in ("TT", "EE", "BB", "TE")}
        self.lCl = {k: theory.lCl(k, L) for k in ("TT", "EE", "BB", "TE")}
        z = np.zeros(self.shape)
        self.noise = {"TT": z + (0. if noise2d is None else noise2d)}
        self.noise["EE"] = 2. * self.noise["TT"] if noise2d_P is None else z + noise2d_P
        self.noise["BB"] = self.noise["EE"] if noise2d_B is None else z + noise2d_B
        self.beam = z + (1. if beam2d is None else beam2d)
        self.fmask = {"TT": kmask, "EE": kmask_P, "BB": kmask_P}
        self.fmaskK = kmask_K
        self.Nlkk, self.AL = {}, {}

    def WXY(self, XY):
        X, Y = XY
        if Y == "B":
            Y = "E"
        grad = X + Y
        with np.errstate(divide="ignore", invalid="ignore"):
            tot = self.lCl[X + X] * self.beam ** 2. + self.noise[X + X]
            W = _fmask(np.nan_to_num(self.uCl[grad] / tot, posinf=0., neginf=0.) * self.beam, self.fmask[X + X])
        W[self.modLMap > self.gradCut] = 0.
        W[self.modLMap >= self.bigell] = 0.
        return W

    def WY(self, YY):
        with np.errstate(divide="ignore", invalid="ignore"):
            tot = self.lCl[YY] * self.beam ** 2. + self.noise[YY]
            W = _fmask(np.nan_to_num(1. / tot, posinf=0., neginf=0.) * self.beam, self.fmask[YY])
        W[self.modLMap >= self.bigell] = 0.
        return W

    def getNlkk2d(self, XY):
        """Returns the multiplier applied in kappa_from_map, N_L * 2/(L(L+1)); stores N_L^{kk}."""
        lx, ly, L = self.lxMap, self.lyMap, self.modLMap
        terms = []
        if XY == "TT":
            C = self.uCl["TT"]
            WXY = self.WXY("TT") * self.beam
            WY = self.WY("TT") * self.beam
            r = 2. ** 0.25
            for e1, e2 in ((lx, lx), (ly, ly), (r * lx, r * ly)):
                preF, preG = e1 * e2 * C * WXY, WY
                preFX, preGX = e1 * WXY, e2 * C * WY
                terms.append(e1 * e2 * _fft(_ifft(preF) * _ifft(preG) + _ifft(preFX) * _ifft(preGX)))
        elif XY == "EB":
            C = self.uCl["EE"]
            lxh, lyh = self.lxHatMap, self.lyHatMap
            s2, c2 = 2. * lxh * lyh, lyh * lyh - lxh * lxh
            fF = (s2 ** 2., c2 ** 2., 1.j * np.sqrt(2.) * s2 * c2)
            fG = (c2 ** 2., s2 ** 2., 1.j * np.sqrt(2.) * s2 * c2)
            WXY = self.WXY("EB") * self.beam
            WY = self.WY("BB") * self.beam
            for ellsq in (lx * lx, ly * ly, np.sqrt(2.) * lx * ly):
                preF, preG = ellsq * C * WXY, WY
                for a, b in zip(fF, fG):
                    terms.append(ellsq * _fft(_ifft(preF * a) * _ifft(preG * b)))
        else:
            raise NotImplementedError(XY)
        ALinv = np.real(np.sum(terms, axis=0))
        with np.errstate(divide="ignore", invalid="ignore"):
            alval = np.nan_to_num(1. / ALinv, posinf=0., neginf=0.)
        alval = _fmask(alval, self.fmaskK)
        NL = (L ** 2.) * ((L + 1.) ** 2.) * alval / 4.
        NL[(L >= self.bigell) | (L < 2.)] = 0.
        ret = np.nan_to_num(NL.real * self.pixScaleX * self.pixScaleY)
        self.Nlkk[XY] = ret.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            self.AL[XY] = ret * 2. * np.nan_to_num(1. / L / (L + 1.), posinf=0., neginf=0.)
        return self.AL[XY]


class qest:
    """qest(shape,wcs,theory,noise2d=,beam2d=,kmask=,noise2d_P=,kmask_P=,kmask_K=,pol=,grad_cut=,
    unlensed_equals_lensed=,bigell=) as constructed at tutorials/tt_verification.ipynb:81."""

    def __init__(self, shape, wcs, theory, noise2d=None, beam2d=None, kmask=None, noise2d_P=None, kmask_P=None,
                 kmask_K=None, pol=False, grad_cut=None, unlensed_equals_lensed=False, bigell=9000, noise2d_B=None,
                 method="cylindrical"):
        self.shape, self.wcs = shape, wcs
        self.N = QuadNorm(shape, wcs, theory, noise2d, noise2d_P, noise2d_B, beam2d, kmask, kmask_P, kmask_K,
                          grad_cut, unlensed_equals_lensed, bigell, method)
        self.pol = pol
        self.phaseY = np.cos(2. * self.N.thetaMap) + 1.j * np.sin(2. * self.N.thetaMap)
        self.N.getNlkk2d("TT")
        if pol:
            self.N.getNlkk2d("EB")

    def kappa_from_map(self, XY, T2DData, E2DData=None, B2DData=None, T2DDataY=None, E2DDataY=None, B2DDataY=None,
                       alreadyFTed=False, returnFt=False):
        f = (lambda a: None if a is None else np.asarray(a)) if alreadyFTed else (lambda a: None if a is None else _fft(a))
        kX = {"T": f(T2DData), "E": f(E2DData), "B": f(B2DData)}
        kY = {"T": f(T2DDataY), "E": f(E2DDataY), "B": f(B2DDataY)}
        for k in kY:
            if kY[k] is None:
                kY[k] = kX[k]
        X, Y = XY
        N = self.N
        WXY, WY = N.WXY(XY), N.WY(Y + Y)
        lx, ly = N.lxMap, N.lyMap
        phaseY = self.phaseY if (Y in ("E", "B")) else 1.
        phaseB = 1.j if Y == "B" else 1.
        high_star = _ifft(kY[Y] * WY * phaseY * phaseB).conjugate()
        kPx = _fft(_ifft(1.j * lx * kX[X] * WXY * phaseY) * high_star)
        kPy = _fft(_ifft(1.j * ly * kX[X] * WXY * phaseY) * high_star)
        raw = _ifft(1.j * lx * kPx + 1.j * ly * kPy).real
        kappaft = -_fmask(np.nan_to_num(N.AL[XY]) * _fft(raw), N.fmaskK)
        if returnFt:
            return kappaft
        return enmap.ndmap(_ifft(kappaft).real, self.wcs)

    reconstruct = kappa_from_map


class _TableNorm:
    """QuadNorm stand-in holding ready-made filter tables (bench.py's cpu_baseline: the per-realisation chain is
    what is timed, so the set-up arithmetic above is skipped and the filters are handed in)."""

    def __init__(self, shape, wcs, tables, AL, kmask_K, method="cylindrical"):
        ly, lx = enmap.laxes(shape, wcs, method)
        self.lyMap, self.lxMap = np.meshgrid(ly, lx, indexing="ij")
        self.thetaMap = np.arctan2(self.lyMap, self.lxMap)
        self._t, self.AL, self.fmaskK = dict(tables), dict(AL), kmask_K

    def WXY(self, XY):
        return self._t["WXY_" + XY]

    def WY(self, YY):
        return self._t["WY_" + YY]


def qest_from_tables(shape, wcs, tables, AL, kmask_K=None, method="cylindrical"):
    """qest whose kappa_from_map runs the chain above on given tables: tables = {"WXY_TT": .., "WY_TT": ..,
    "WXY_EB": .., "WY_BB": ..} (full-plane float64), AL = {"TT": .., "EB": ..}."""
    q = qest.__new__(qest)
    q.shape, q.wcs = shape, wcs
    q.N = _TableNorm(shape, wcs, tables, AL, kmask_K, method)
    q.pol = "EB" in AL
    q.phaseY = np.cos(2. * q.N.thetaMap) + 1.j * np.sin(2. * q.N.thetaMap)
    return q
