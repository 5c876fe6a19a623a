"""Worker of tests/test_gpu_dist.py: launched by torch.distributed.run with 2 ranks, one per GPU.  Every rank runs
its contiguous shard of the seeds (mpi.py:78-91) through the fused pipeline and the TT estimator, then the packed
Statistics triple and the packed mean-field stack are summed with ONE ncclAllReduce each, issued by liborphx.so
(ox_pipeline_allreduce / ox_qe_meanfield_allreduce).  Rank 0 repeats the whole job alone and compares."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orphics_b200 import _capi, maps, stats, mpi, cosmology, lensing  # noqa: E402

rank, local, ws = mpi.init_process_group()
_capi.set_device(local)
comm = mpi.NcclComm(rank, ws)
npix, res, nsim, B = 512, 2.0, 12, 4
shape, wcs = maps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res)
th = cosmology.default_theory()
g = maps.Geometry.get(shape, wcs)
modl = g.modlmap()
ps = cosmology.power_from_theory(np.arange(0, modl.max() + 1, 1.0), th, lensed=True, pol=False)
edges = np.arange(200, 2600, 100.0)


def run(seeds):
    mg = maps.MapGen(shape, wcs, ps, noise="philox", max_batch=B)
    fc = maps.FourierCalc(shape, wcs, max_batch=B)
    b = stats.bin2D(modl, edges, geometry=g)
    pipe = maps.SimPipeline(mg, fc, b, window=np.asarray(maps.get_taper(shape, wcs)[0]))
    bp = pipe.run(seeds, keep_maps=True)
    return pipe, bp


_, tasks = mpi.mpi_distribute(nsim, ws)
pipe, bp = run([1000 + i for i in tasks[rank]])
pipe.allreduce(comm)
N, S, Cm = pipe.stats()

kw = dict(noise2d=np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2, beam2d=maps.gauss_beam(modl, 1.5),
          kmask=maps.mask_kspace(shape, wcs, lmin=300, lmax=2000), kmask_K=maps.mask_kspace(shape, wcs, lmin=20, lmax=3500),
          unlensed_equals_lensed=True, max_batch=B)
q = lensing.qest(shape, wcs, th, **kw)
mgq = maps.MapGen(shape, wcs, ps, noise="philox", max_batch=B)


def qe_stack(qq, seeds):
    qq.reset_meanfield("TT")
    for c0 in range(0, len(seeds), B):
        qq.kappa_from_maps("TT", mgq.get_maps(seeds[c0:c0 + B]), returnFt=True, accumulate_meanfield=True)


qe_stack(q, [1000 + i for i in tasks[rank]])
q.allreduce_meanfield("TT", comm)
acc, cnt = q.meanfield("TT")
_capi.synchronize()

# the host-side Statistics class with the same communicator: labels on a subset of ranks (orphics/tests/test_stats.py's
# closed forms in P), ONE packed all-reduce through ox_comm_allreduce_f64
st = stats.Statistics(comm=True if ws > 1 else None, nccl=comm)
for i in range(rank + 1):
    st.add("vec", np.arange(3.0) + rank)
    st.add_stack("map", np.full((4, 5), float(rank + 1)))
if rank == ws - 1:
    st.add("last_only", np.array([7.0, 9.0]))
st.allreduce()
P = ws
exp_n = P * (P + 1) // 2
exp_sum = sum((r + 1) * (np.arange(3.0) + r) for r in range(P))
st_ok = (st.count("vec") == exp_n and np.array_equal(st._sum["vec"], exp_sum) and st.count("last_only") == 1
         and np.array_equal(st._sum["last_only"], [7.0, 9.0]) and st.stack_count("map") == exp_n
         and np.array_equal(st._stack["map"], np.full((4, 5), float(sum((r + 1) ** 2 for r in range(P))))))

ok = bool(st_ok)
report = {"ws": ws, "nccl": comm.nccl_version(), "statistics_class_ok": bool(st_ok)}
if rank == 0:
    pipe1, bp1 = run([1000 + i for i in range(nsim)])
    N1, S1, C1 = pipe1.stats()
    report["N"] = [N, N1]
    report["sum_err"] = float(np.max(np.abs(S - S1) / np.abs(S1)))
    report["cross_err"] = float(np.max(np.abs(Cm - C1)) / np.max(np.abs(C1)))
    ok &= (N == N1 == nsim) and report["sum_err"] < 1e-13 and report["cross_err"] < 1e-13
    ok &= bool(np.array_equal(bp, bp1[:len(tasks[0])]))          # per-sim bandpowers do not depend on the sharding
    q1 = lensing.qest(shape, wcs, th, quadnorm=q.N, **kw)
    qe_stack(q1, [1000 + i for i in range(nsim)])
    acc1, cnt1 = q1.meanfield("TT")
    report["mf_count"] = [cnt, cnt1]
    report["mf_err"] = float(np.max(np.abs(acc - acc1)) / np.max(np.abs(acc1)))
    ok &= (cnt == cnt1 == nsim) and report["mf_err"] < 1e-13
    report["ok"] = bool(ok)
    print("NCCL_WORKER " + json.dumps(report), flush=True)
comm.free()
if ws > 1:
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
