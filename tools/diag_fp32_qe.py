"""Diagnostic (GPU): where does the float32 quadratic estimator lose accuracy?  Prints max-norm relative errors
against the float64 oracle chain for TT / EB at 512^2 on the hand-written path and on the cuFFT chain, and of the
kappa_hat(l) per annulus."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orphics_b200 import maps, lensing, cosmology  # noqa: E402
from oracle import qe_np, maps_np as omaps, enmap_np as oenmap, theory as otheory  # noqa: E402

npix = int(sys.argv[1]) if len(sys.argv) > 1 else 512
res = 2.0 if npix <= 512 else 0.5
th = otheory.load_theory()
shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
so, wo = omaps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
modl = np.asarray(oenmap.modlmap(so, wo))
n2d = np.zeros(so) + (1.0 * np.pi / 180 / 60) ** 2
tm = np.asarray(omaps.mask_kspace(so, wo, lmin=300, lmax=2000))
km = np.asarray(omaps.mask_kspace(so, wo, lmin=20, lmax=3500))
kw = dict(noise2d=n2d, beam2d=omaps.gauss_beam(modl, 1.5), kmask=tm, noise2d_P=2 * n2d, kmask_P=tm, kmask_K=km, pol=True,
          unlensed_equals_lensed=True)
import scipy.fft
with scipy.fft.set_workers(16):
    qo = qe_np.qest(so, wo, th, **kw)
rng = np.random.RandomState(5)
T = (rng.standard_normal(shape) * 50).astype(np.float32)
E, B = (rng.standard_normal((2,) + tuple(shape)) * 3).astype(np.float32)


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


for env in ("", "cufft"):
    if env:
        os.environ["ORPHX_QE"] = env
    else:
        os.environ.pop("ORPHX_QE", None)
    q32 = lensing.qest(shape, wcs, cosmology.default_theory(), dtype=np.float32, **kw)
    for XY in ("TT", "EB"):
        qo.N.AL[XY] = np.asarray(q32.N.AL[XY], dtype=np.float64)
    with scipy.fft.set_workers(16):
        wantT = qo.kappa_from_map("TT", T.astype(np.float64), returnFt=True)
        wantE = qo.kappa_from_map("EB", None, E.astype(np.float64), B.astype(np.float64), returnFt=True)
    for XY, got, want in (("TT", q32.kappa_from_map("TT", T, returnFt=True), wantT),
                          ("EB", q32.kappa_from_map("EB", None, E, B, returnFt=True), wantE)):
        got = np.asarray(got, dtype=np.complex128)
        kmap_err = rel(np.fft.ifft2(got).real, np.fft.ifft2(want).real)
        line = f"{npix} path={q32.path(XY):9s} {XY} kappa_hat relerr {rel(got, want):.3e}  kappa map relerr {kmap_err:.3e}  per-L:"
        for lo, hi in ((20, 100), (100, 500), (500, 1500), (1500, 3500)):
            sel = (modl > lo) & (modl <= hi)
            line += f" ({lo},{hi}] {np.sqrt(np.sum(np.abs(got[sel] - want[sel]) ** 2) / np.sum(np.abs(want[sel]) ** 2)):.2e}"
        print(line, flush=True)
    del q32
# plain float32 transforms: FourierCalc.fft on the device vs numpy float64
fc32 = maps.FourierCalc(shape, wcs, dtype=np.float32)
k = np.asarray(fc32.fft(T), dtype=np.complex128)
print("float32 FourierCalc.fft relerr", rel(k, np.fft.fft2(T.astype(np.float64))), flush=True)
