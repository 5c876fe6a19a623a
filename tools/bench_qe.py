"""Secondary benchmark (BASELINE configs[3]/[4]): TT / EB quadratic-estimator reconstructions per second
with inputs resident in HBM, and the fraction of the SURVEY 8d roofline (TT: 38 s N bytes per
realisation incl. the mean-field accumulate; EB ~ 76 s N).  Usage: python tools/bench_qe.py [npix] [batch] [f64|f32] [TT|EB]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orphics_b200 import _capi, maps, lensing, cosmology  # noqa: E402
from orphics_b200._capi import lib, check, ptr  # noqa: E402

npix = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dt = sys.argv[3] if len(sys.argv) > 3 else "f64"
est = sys.argv[4] if len(sys.argv) > 4 else "TT"
rdt = np.float32 if dt == "f32" else np.float64
cdt = np.complex64 if dt == "f32" else np.complex128
s = 4 if dt == "f32" else 8
shape, wcs = maps.rect_geometry(width_arcmin=npix * 0.5, px_res_arcmin=0.5)
th = cosmology.default_theory()
g = maps.Geometry.get(shape, wcs)
modl = g.modlmap()
beam = maps.gauss_beam(modl, 1.5)
n2d = np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2
tmask = maps.mask_kspace(shape, wcs, lmin=300, lmax=2000)
kmask = maps.mask_kspace(shape, wcs, lmin=20, lmax=3500)
q = lensing.qest(shape, wcs, th, noise2d=n2d, beam2d=beam, kmask=tmask, kmask_P=tmask, kmask_K=kmask, pol=(est == "EB"),
                 unlensed_equals_lensed=True, dtype=rdt, max_batch=nb)
h, real_path = q._plans[est]
rng = np.random.RandomState(0)
N = npix * npix
if est == "TT":
    x = _capi.DeviceBuffer(nb * N * s).upload((rng.standard_normal((nb, npix, npix)) * 50).astype(rdt))
    y = None
    already = 0
else:
    mk = lambda: (rng.standard_normal((nb, npix, npix)) + 1j * rng.standard_normal((nb, npix, npix))).astype(cdt)
    x = _capi.DeviceBuffer(nb * N * 2 * s).upload(mk())
    y = _capi.DeviceBuffer(nb * N * 2 * s).upload(mk())
    already = 1
out = _capi.DeviceBuffer(nb * N * 2 * s)


def step():
    check(lib.ox_qe_reconstruct(h, C.c_void_p(x.ptr), C.c_void_p(y.ptr) if y else None, _capi.OX_DEVICE, nb, already, 1, 1,
                                C.c_void_p(out.ptr), _capi.OX_DEVICE))


for _ in range(3):
    step()
_capi.synchronize()
t = _capi.Timer()
K = 10
l0 = _capi.launch_count()
t.start()
for _ in range(K):
    step()
t.stop()
ms = t.elapsed_ms()
rate = K * nb / (ms / 1e3)
bytes_per = (38 if est == "TT" else 76) * s * N
peak = 6550.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"metric": f"{est} QE reconstructions/s at {npix}^2 {dt} (kappa_hat(l) + mean-field accumulate, inputs in HBM)",
                  "value": rate, "ms_per_realisation": ms / (K * nb), "batch": nb, "half_plane_path": bool(real_path),
                  "algorithmic_bytes_per_realisation": bytes_per, "roofline_frac": rate * bytes_per / 1e9 / peak,
                  "peak_gbs": peak, "gpu_launches": _capi.launch_count() - l0, "device": _capi.device_name()}))
