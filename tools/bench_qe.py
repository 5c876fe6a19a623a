"""Secondary benchmark (BASELINE configs[3]/[4]): TT / EB quadratic-estimator reconstructions per second
with inputs resident in HBM, and the fraction of the SURVEY 8d roofline (TT: 38 s N bytes per
realisation incl. the mean-field accumulate; EB ~ 76 s N).  Usage: python tools/bench_qe.py [npix] [batch] [f64|f32] [TT|EB] [nreal]
Under torchrun (WORLD_SIZE > 1) the nreal realisations (default 10 steps x batch per rank) are split over the
ranks with the reference's rule (mpi.py:78-91) and the packed mean-field stack + count are summed with one NCCL
all-reduce (ox_qe_meanfield_allreduce) inside the timed region (BASELINE configs[3]: 512 realisations plus mean-field allreduce)."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orphics_b200 import _capi, maps, lensing, cosmology, mpi  # noqa: E402
from orphics_b200._capi import lib, check, ptr  # noqa: E402

npix = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dt = sys.argv[3] if len(sys.argv) > 3 else "f64"
est = sys.argv[4] if len(sys.argv) > 4 else "TT"
rank, local, ws = mpi.init_process_group()
_capi.require_device()
_capi.set_device(local)
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
h = q._plans[est][0]
rng = np.random.RandomState(0)
N = npix * npix
if est == "TT":
    x = _capi.DeviceBuffer(nb * N * s).upload((rng.standard_normal((nb, npix, npix)) * 50).astype(rdt))
    y = None
    already = 0
else:
    # real E and B maps in HBM (their transforms are Hermitian, as those of iqu2teb are); the SURVEY 8d
    # accounting (76 s N) includes the two input transforms
    x = _capi.DeviceBuffer(nb * N * s).upload((rng.standard_normal((nb, npix, npix)) * 3).astype(rdt))
    y = _capi.DeviceBuffer(nb * N * s).upload((rng.standard_normal((nb, npix, npix)) * 3).astype(rdt))
    already = 0
out = _capi.DeviceBuffer(nb * N * 2 * s)


def step():
    check(lib.ox_qe_reconstruct(h, C.c_void_p(x.ptr), C.c_void_p(y.ptr) if y else None, _capi.OX_DEVICE, nb, already, 1, 1,
                                C.c_void_p(out.ptr), _capi.OX_DEVICE))


nreal = int(sys.argv[5]) if len(sys.argv) > 5 else 10 * nb * ws
mine = len(mpi.mpi_distribute(nreal, ws)[1][rank])
K = max(1, mine // nb)
comm = mpi.NcclComm(rank, ws)     # data plane: one ncclAllReduce of the packed [stack | count] issued by liborphx.so
if ws > 1:
    import torch
    import torch.distributed as dist
for _ in range(3):
    step()
q.allreduce_meanfield(est, comm)
check(lib.ox_qe_meanfield_reset(h))
_capi.synchronize()
if ws > 1:
    dist.barrier()
    torch.cuda.synchronize()
t = _capi.Timer()
l0 = _capi.launch_count()
t.start()
for _ in range(K):
    step()
t.stop()
ms = t.elapsed_ms()
ar_ms = 0.0
if ws > 1:
    t2 = _capi.Timer()
    t2.start()
    q.allreduce_meanfield(est, comm)     # Statistics.add_stack / allreduce (stats.py:1227-1228)
    t2.stop()
    ar_ms = t2.elapsed_ms()
    tt = torch.tensor([ms + ar_ms], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
else:
    total_ms = ms
count = q.meanfield_count(est)
rate = ws * K * nb / (total_ms / 1e3)
bytes_per = (38 if est == "TT" else 76) * s * N
peak = 6550.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
if rank == 0:
    print(json.dumps({"metric": f"{est} QE reconstructions/s at {npix}^2 {dt} (kappa_hat(l) + mean-field accumulate, inputs in HBM)",
                      "value": rate, "n_gpus": ws, "realisations": ws * K * nb, "meanfield_count_after_allreduce": count,
                      "ms_per_realisation_per_gpu": ms / (K * nb), "meanfield_allreduce_ms": ar_ms, "batch": nb,
                      "path": q.path(est), "algorithmic_bytes_per_realisation": bytes_per,
                      "roofline_frac": rate / ws * bytes_per / 1e9 / peak, "peak_gbs": peak,
                      "gpu_launches": _capi.launch_count() - l0, "device": _capi.device_name()}))
if ws > 1:
    dist.destroy_process_group()
