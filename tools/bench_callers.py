"""Timing of the split / cross-spectrum callers through their reference-facing Python signatures
(host numpy arrays in and out, so PCIe staging is inside the timed region) next to the numpy oracle on one host
core.  Usage: python tools/bench_callers.py [npix] [nsplits]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orphics_b200 import _capi, maps, lensing, cosmology  # noqa: E402
from oracle import maps_np as omaps, qe_np, lensing_np, theory as otheory, enmap_np as oenmap  # noqa: E402

npix = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
_capi.require_device()
shape, wcs = maps.rect_geometry(width_arcmin=npix * 0.5, px_res_arcmin=0.5)
so, wo = omaps.rect_geometry(width_arcmin=npix * 0.5, px_res_arcmin=0.5)
rng = np.random.RandomState(0)
N = npix * npix
res = {"npix": npix, "nsplits": n, "device": _capi.device_name()}


def timeit(f, reps=3):
    f()
    _capi.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    _capi.synchronize()
    return (time.perf_counter() - t0) / reps


# noise_from_splits, IQU
splits = rng.standard_normal((n, 3, npix, npix))
fc3 = maps.FourierCalc((3,) + tuple(shape), wcs)
t = timeit(lambda: maps.noise_from_splits(maps.ndmap(splits, wcs), fourier_calc=fc3))
t0 = time.perf_counter()
omaps.noise_from_splits(oenmap.ndmap(splits, wo), fourier_calc=omaps.FourierCalc((3,) + tuple(so), wo))
res["noise_from_splits_IQU"] = {"ours_s": t, "oracle_1core_s": time.perf_counter() - t0,
                                "bytes_h2d": splits.astype(np.float64).nbytes, "bytes_d2h": 2 * 9 * N * 8}
# split_calc
ks = (rng.standard_normal((n, npix, npix)) + 1j * rng.standard_normal((n, npix, npix)))
fc1 = maps.FourierCalc(shape, wcs)
t = timeit(lambda: maps.split_calc(ks, ks, ks.mean(0), ks.mean(0), fourier_calc=fc1))
t0 = time.perf_counter()
omaps.split_calc(ks, ks, ks.mean(0), ks.mean(0), fourier_calc=omaps.FourierCalc(so, wo))
res["split_calc"] = {"ours_s": t, "oracle_1core_s": time.perf_counter() - t0, "bytes_h2d": 2 * ks.nbytes + 2 * N * 16, "bytes_d2h": 3 * N * 8}
# SplitLensing.cross_estimator
modl = maps.Geometry.get(shape, wcs).modlmap()
kw = dict(noise2d=np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2, beam2d=maps.gauss_beam(modl, 1.5),
          kmask=maps.mask_kspace(shape, wcs, lmin=300, lmax=2000), kmask_K=maps.mask_kspace(shape, wcs, lmin=20, lmax=3500),
          unlensed_equals_lensed=True)
q = lensing.qest(shape, wcs, cosmology.default_theory(), max_batch=8, **kw)
sl = lensing.SplitLensing(shape, wcs, q)
l0 = _capi.launch_count()
t = timeit(lambda: sl.cross_estimator(ks), reps=2)
nl = (_capi.launch_count() - l0) // 3
qo = qe_np.qest(so, wo, otheory.load_theory(), **{k: np.asarray(v) if hasattr(v, "shape") else v for k, v in kw.items()})
t0 = time.perf_counter()
qo.kappa_from_map("TT", ks[0], T2DDataY=ks[1], alreadyFTed=True, returnFt=True)
tq = time.perf_counter() - t0
nq = 1 + 3 * n + n * (n - 1)
res["split_lensing_cross_estimator"] = {"ours_s": t, "qe_path": q.path("TT"), "estimators": nq, "gpu_launches": nl,
                                        "oracle_1core_s_extrapolated": tq * nq, "oracle_one_qfrag_s": tq,
                                        "bytes_h2d": (n + 1) * N * 16, "bytes_d2h": N * 8}
print(json.dumps(res))
