#!/usr/bin/env python
"""Benchmark of the north-star path: maps/sec through sim -> FFT -> power2d -> bin2D at
2048^2 fp64 (BASELINE.json metric; headline workload = configs[1], T-only GRF sims from the CAMB
lensed TT spectrum at 0.5 arcmin, 72 bandpowers), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, liborphx.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle) on host cores

A step = one batch of --batch maps per GPU through ox_pipeline_run.  On power-of-two maps the
pipeline is three hand-written kernels (orphics_b200/csrc/ox_fused.cu):
  K_A Philox noise x covsqrt -> inverse FFT along y        (writes the transposed half plane)
  K_B inverse FFT along x -> real map (stored) x taper -> forward FFT along x
  K_C forward FFT along y -> |k|^2 -> deterministic annular binning -> bandpowers
followed by the Statistics triple; at the end the packed [N | SUM | CROSS] is summed over the ranks
with ONE ncclAllReduce issued by liborphx.so (ox_pipeline_allreduce).  torch.distributed is only the
rendezvous (NCCL unique id, barrier, max over ranks of the timings).

The JSON line also carries a "configs" block with the other BASELINE configurations measured the same
way (value, e2e, roofline of the dominant kernel, cpu_baseline):
  configs[2]  IQU 2048^2 with the TEB rotation, 6 binned spectra
  configs[3]  TT quadratic estimator on 4096^2 maps, 512 realisations sharded over the GPUs + mean-field all-reduce
  configs[4]  EB quadratic estimator at 8192^2, fp64 and fp32
(--configs none skips them, --configs 2,3 selects).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EDGES = np.arange(100, 3000, 40.0)      # tutorials/demo-grf.ipynb:159
KEDGES = np.linspace(20, 3500, 20)      # tutorials/tt_verification.ipynb:600
SEED0 = 1000                            # sim i uses seed 1000+i
METRIC = "maps/sec sim->FFT->power2d->bin2D at 2048^2 fp64"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128, help="timed steps (128 x 64 maps ~ 0.5 s on one B200)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="maps per step per GPU")
    ap.add_argument("--npix", type=int, default=2048)
    ap.add_argument("--res", type=float, default=0.5, help="pixel size, arcmin")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--noise", default="philox_hermitian", choices=["philox", "philox_hermitian"],
                    help="philox_hermitian draws the Hermitian half plane directly (N normals per map); "
                         "philox draws the reference's full complex plane (4N normals per map)")
    ap.add_argument("--no-keep-maps", action="store_true", help="do not store the real-space maps in HBM")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (variants, per-call API)")
    ap.add_argument("--pol", action="store_true", help="headline workload = IQU sims, 6 spectra (configs[2])")
    ap.add_argument("--no-window", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=24, help="maps timed for cpu_baseline: ~10 s on one host core (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="all", help="'all', 'none' or a list out of 2,3,4: the other BASELINE configs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --steps x --batch maps per GPU; strong: --total maps shared by the GPUs (configs[1] literal)")
    ap.add_argument("--total", type=int, default=1024, help="maps of the whole job under --scaling strong")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled through NVML from a thread of this process every
    ~4 ms while the timed region runs (nvidia-smi's own loop is too coarse for a 70 ms region)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake", 0x80), ("sync_boost", 0x10))

    def __init__(self, gpu_index, period=0.004):
        self.rows, self.gpu, self.period, self.h, self.err = [], gpu_index, period, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:           # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), float(pw), int(rs)))
            except Exception as e:       # noqa: BLE001
                self.err = repr(e)
                break
            self._stop.wait(self.period)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)], "samples": 0}
        self._stop.set()
        self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"], "samples": 0}
        sm = np.array([r[0] for r in self.rows])
        pw = np.array([r[1] for r in self.rows])
        bits = 0
        busy = np.ones(len(sm), dtype=bool)      # the sampler only runs inside the timed region: every sample is under load
        for r, b in zip(self.rows, busy):
            if b:
                bits |= r[2] & ~0x1              # (bit 0 = "GPU idle", not a throttle reason)
        reasons = [name for name, bit in self.REASONS if bits & bit]
        return {"sm_mhz": float(np.median(sm[busy])), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "power_w_max": float(pw.max()), "samples": int(len(sm)), "samples_under_load": int(busy.sum())}


# --------------------------------------------------------------------------- CPU reference path (oracle)
_ORACLE = {}


def _oracle_setup(npix, res, pol, window):
    from oracle import enmap_np as oenmap, maps_np as omaps, stats_np as ostats, theory as otheory
    so, wo = omaps.rect_geometry(width_arcmin=npix * res, px_res_arcmin=res, pol=pol)
    modl = np.asarray(oenmap.modlmap(so, wo))
    th = otheory.load_theory()
    ps = otheory.power_from_theory(np.arange(0, modl.max() + 1, 1.0), th, lensed=True, pol=pol)
    _ORACLE.update(mg=omaps.MapGen(so, wo, ps), fc=omaps.FourierCalc(so, wo), b=ostats.bin2D(modl, EDGES), pol=pol,
                   taper=np.asarray(omaps.get_taper(so, wo)[0]) if window else None)


def _oracle_one(seed):
    """The reference path for one map: MapGen.get_map (maps.py:1576) x taper ->
    FourierCalc.power2d (maps.py:1639) -> bin2D.bin (stats.py:790)."""
    o = _ORACLE
    m = o["mg"].get_map(seed=seed)
    if o["taper"] is not None:
        m = m * o["taper"]
    p2d = o["fc"].power2d(m)[0]
    if o["pol"]:
        return np.array([o["b"].bin(p2d[i, j])[1] for i in range(3) for j in range(i, 3)])
    return o["b"].bin(p2d)[1]


def cpu_baseline_sample(npix, res, pol, window, nmaps):
    """1 core, nmaps maps of the same workload (set-up excluded, 1 warm-up map)."""
    _oracle_setup(npix, res, pol, window)
    _oracle_one(SEED0)
    t0 = time.perf_counter()
    for i in range(nmaps):
        _oracle_one(SEED0 + i)
    dt = time.perf_counter() - t0
    return {"value": nmaps / dt, "unit": "maps/s", "cores": 1, "kind": "port",
            "sample": f"{nmaps} maps of {npix}^2 {'IQU' if pol else 'T'} through the numpy oracle "
                      f"(MapGen.get_map [numpy MT19937 full-plane noise] -> taper -> power2d -> bin2D.bin), 1 process / 1 thread, set-up excluded"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (numpy oracle; the reference
    itself cannot be imported: pixell is absent) on all host cores, one process per core
    with realisations split as the reference does under MPI (mpi.py:78-91).  A step is the same
    --batch maps as the GPU arm's step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    ncpu = host_cores()
    mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    per_proc_gb = 1.5 * (args.npix / 2048.0) ** 2 * (3 if args.pol else 1)
    P = max(1, int(min(ncpu, 64, args.batch, mem_gb * 0.5 / per_proc_gb)))
    ctx = mp.get_context("fork")
    window = not args.no_window
    with ctx.Pool(P, initializer=_oracle_setup, initargs=(args.npix, args.res, args.pol, window)) as pool:
        step_maps = args.batch
        for w in range(args.warmup):
            pool.map(_oracle_one, [SEED0 + i for i in range(P)], chunksize=1)      # warm-up: one map per process
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_oracle_one, [SEED0 + k * step_maps + i for i in range(step_maps)], chunksize=1)
        dt = time.perf_counter() - t0
    value = args.steps * step_maps / dt
    sample = (f"{step_maps} maps per step over {P} processes (1 FFT thread each) of {args.npix}^2 "
              f"{'IQU' if args.pol else 'T'}; numpy-oracle restatement of MapGen.get_map -> taper -> power2d -> bin2D.bin")
    cfg = workload_config(args, step_maps, 1)
    cfg["noise"] = "numpy MT19937 full complex plane (np.random.seed + rand_gauss_harm, maps.py:1577-1578)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "maps/s", "cores": P, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, batch, ngpu):
    return {"workload": f"configs[1]: batch of T-only {args.npix}x{args.npix} ({args.res}') GRF sims from "
                        "cosmo2017_10K_acc3 lensed TT, cosine taper, 72 binned TT bandpowers (edges arange(100,3000,40))"
            if not args.pol else
            f"configs[2]: IQU {args.npix}x{args.npix} ({args.res}') sims with TEB rotation, 6 binned auto/cross spectra",
            "maps_per_step_per_gpu": batch, "global_batch": batch * ngpu, "npix": args.npix, "ncomp": 3 if args.pol else 1,
            "nbins": len(EDGES) - 1, "noise": args.noise, "window": not args.no_window,
            "maps_materialised_in_hbm": not args.no_keep_maps,
            "parallelism": f"realisations sharded over {ngpu} GPU(s) (mpi_distribute rule), one ncclAllReduce of the packed Statistics triple",
            "l2": "working set per step (batch x 67 MB of maps+Fourier planes) >> 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------- this repo
class Ctx:
    """Per-process state shared by the measurements: rank, communicator, peak, timing helpers."""

    def __init__(self):
        from orphics_b200 import _capi, mpi
        self.capi, self.mpi = _capi, mpi
        self.rank, self.local, self.ws = mpi.init_process_group()
        self.dist = self.torch = None
        if self.ws > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
        _capi.require_device()
        _capi.set_device(self.local)
        self.comm = mpi.NcclComm(self.rank, self.ws)     # data plane: ncclAllReduce issued by liborphx.so
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = ("MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks
                         else "B200_PROFILING.md fallback 6650 GB/s (of fallback)")
        self.traffic = {}
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        except Exception:
            pass

    def barrier(self):
        self.capi.synchronize()
        if self.ws > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.ws == 1:
            return float(v)
        t = self.torch.tensor([v], device=f"cuda:{self.local}", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, nsteps, finish=None):
        """Device time (CUDA events on the library stream) of nsteps calls of step(k) [+ finish()], bracketed by a
        barrier + synchronize on both sides; max over ranks.  Returns (ms, launches)."""
        self.barrier()
        tm = self.capi.Timer()
        l0 = self.capi.launch_count()
        tm.start()
        for k in range(nsteps):
            step(k)
        if finish is not None:
            finish()
        tm.stop()
        ms = tm.elapsed_ms()
        launches = self.capi.launch_count() - l0
        self.barrier()
        return self.max_over_ranks(ms), launches

    def wall(self, step, nsteps, finish=None):
        """Wall-clock seconds of nsteps host-facing calls (each synchronises on its own D2H); max over ranks."""
        self.barrier()
        t0 = time.perf_counter()
        for k in range(nsteps):
            step(k)
        if finish is not None:
            finish()
        self.capi.synchronize()
        dt = time.perf_counter() - t0
        self.barrier()
        return self.max_over_ranks(dt)

    def roofline(self, stage_ms, alg_bytes, key, note=None):
        """roofline object of the dominant stage: achieved = algorithmic bytes per launch / its CUDA-event time."""
        dom = max((k for k in stage_ms if alg_bytes.get(k)), key=lambda k: stage_ms[k])
        gbs = alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
        tr = self.traffic.get(key, {}).get(dom)
        r = {"bound": "hbm", "kernel": dom, "achieved": gbs, "peak": self.peak, "unit": "GB/s", "frac": gbs / self.peak,
             "traffic": tr, "peak_source": self.peak_src, "algorithmic_bytes_per_launch": alg_bytes[dom],
             "ms_per_launch": stage_ms[dom]}
        if tr:
            r["traffic_over_algorithmic"] = tr / alg_bytes[dom]
        if note:
            r["note"] = note
        return r

    def stages(self, stage_ms, alg_bytes):
        out = {}
        for k, ms in stage_ms.items():
            b = alg_bytes.get(k)
            gbs = b / (ms * 1e-3) / 1e9 if (b and ms > 0) else None
            out[k] = {"ms_per_launch": ms, "algorithmic_bytes_per_launch": b, "achieved_gbs": gbs,
                      "frac": gbs / self.peak if gbs else None}
        return out


def profile_stages(capi, fn, reps=3):
    """Average per-stage device times (ms) of fn() from the library's stage marks."""
    acc = {}
    for _ in range(reps):
        with capi.StageProfile() as p:
            fn()
        for k, v in p.totals().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    return acc


def build_pipeline(args, npix, pol, B, dtype, noise):
    from orphics_b200 import maps, stats, cosmology
    shape, wcs = maps.rect_geometry(width_arcmin=npix * args.res, px_res_arcmin=args.res, pol=pol)
    th = cosmology.default_theory()
    g = maps.Geometry.get(shape, wcs)
    modl = g.modlmap()
    ps = cosmology.power_from_theory(np.arange(0, modl.max() + 1, 1.0), th, lensed=True, pol=pol)
    mg = maps.MapGen(shape, wcs, ps, noise=noise, dtype=dtype, max_batch=B)
    fc = maps.FourierCalc(shape, wcs, dtype=dtype, max_batch=B)
    binner = stats.bin2D(modl, EDGES, geometry=g)
    window = None if args.no_window else np.asarray(maps.get_taper(shape, wcs)[0])
    pipe = maps.SimPipeline(mg, fc, binner, window=window)
    return dict(shape=shape, wcs=wcs, th=th, g=g, ps=ps, mg=mg, fc=fc, binner=binner, window=window, pipe=pipe)


def pipeline_alg_bytes(npix, nc, s, keep, B):
    N = npix * npix
    return {"K_A sim+col_ifft": nc * s * N * B, "K_B row_c2r+taper+r2c": nc * (2 + keep) * s * N * B,
            "K_C col_fft+power+bin": (nc * s * N + N) * B}


def run_headline(cx, args):
    capi, mpi = cx.capi, cx.mpi
    rank, ws = cx.rank, cx.ws
    npix, B, W = args.npix, args.batch, max(args.warmup, 3)
    pol = args.pol
    dtype = np.float32 if args.dtype == "f32" else np.float64
    P = build_pipeline(args, npix, pol, B, dtype, args.noise)
    pipe, window, binner, th = P["pipe"], P["window"], P["binner"], P["th"]
    mode = capi.NOISE_MODES[args.noise]
    # the real-space maps are materialised in HBM (SURVEY 8d anti-short-circuit rule) unless asked not to
    flags = pipe._flags(False, False, keep_maps=not args.no_keep_maps)

    # shard: contiguous blocks of realisations per rank (mpi.py:78-91)
    if args.scaling == "strong":
        total = args.total
        K = max(1, total // (ws * B))
        total = K * B * ws
    else:
        K = args.steps
        total = ws * K * B
    _, tasks = mpi.mpi_distribute(total, ws)
    my = np.array(tasks[rank], dtype=np.int64) + SEED0
    seeds_pin = capi.PinnedArray((K, B), np.int64)
    seeds_pin.array[:] = my.reshape(K, B)
    warm_pin = capi.PinnedArray((B,), np.int64)
    warm_pin.array[:] = np.arange(B) + 7
    out_pin = capi.PinnedArray((B, pipe.nspec, pipe.nbins), np.float64)

    def reduce_stats():
        pipe.allreduce(cx.comm)          # ONE ncclAllReduce of [N | SUM | CROSS] (stats.py:1215-1217), no-op at 1 rank

    # ---- device-resident throughput
    for _ in range(W):
        pipe.run_raw(warm_pin.array, B, mode, flags, None)
    reduce_stats()
    pipe.reset_stats()
    cx.barrier()
    sampler = ClockSampler(cx.local)
    sampler.start()
    ms, launches = cx.timed(lambda k: pipe.run_raw(seeds_pin.array[k], B, mode, flags, None), K, reduce_stats)
    clocks = sampler.stop()
    if ws > 1:
        allc = [None] * ws
        cx.dist.all_gather_object(allc, clocks)
        if rank == 0:
            sms = [c["sm_mhz"] for c in allc if c.get("sm_mhz")]
            clocks = dict(allc[0])
            clocks["per_rank_sm_mhz"] = [c.get("sm_mhz") for c in allc]
            clocks["per_rank_samples"] = [c.get("samples") for c in allc]
            clocks["reasons"] = sorted(set(r for c in allc for r in c.get("reasons", [])))
            if sms:
                clocks["sm_mhz"] = float(min(sms))
    N_stat, S, Cm = pipe.stats()
    value = total / (ms / 1e3)
    # collective alone (device time of one more all-reduce of the packed triple)
    ar_ms = None
    if ws > 1:
        ar_ms, _ = cx.timed(lambda k: reduce_stats(), 1)
        pipe.reset_stats()

    # ---- configs[1] literal: a FIXED batch of 1024 maps shared by the GPUs (strong scaling: 16 steps on one GPU, 2 on
    # eight), timed from the first launch to the end of the all-reduce of the Statistics triple
    strong = None
    if args.scaling == "weak" and not args.no_extras:
        Ks = max(1, 1024 // (ws * B))
        pipe.reset_stats()
        ms_s, _ = cx.timed(lambda k: pipe.run_raw(seeds_pin.array[k % K], B, mode, flags, None), Ks, reduce_stats)
        strong = {"maps_total": int(Ks * B * ws), "steps_per_gpu": int(Ks), "ms_total": ms_s, "value": Ks * B * ws / (ms_s / 1e3),
                  "unit": "maps/s", "note": "strong scaling of configs[1]: 1024 maps shared by the GPUs, all-reduce included"}
        pipe.reset_stats()

    # ---- end to end through the Python API with host buffers: pinned seeds in, bandpowers out, every step
    e2e = None
    if not args.no_e2e:
        pipe.reset_stats()
        dt = cx.wall(lambda k: pipe.run_raw(seeds_pin.array[k], B, mode, flags, capi.ptr(out_pin.array)), K, reduce_stats)
        e2e = {"value": total / dt, "unit": "maps/s", "h2d_bytes_per_step": int(B * 8),
               "d2h_bytes_per_step": int(out_pin.array.nbytes),
               "note": "SimPipeline.run_raw: seeds from pinned host memory -> bandpowers in pinned host memory each step; "
                       "the noise itself is drawn on the device (Philox), as the reference draws it in-process"}

    # ---- the reference's per-call API (rank 0, a few maps): MapGen.get_map -> x taper -> FourierCalc.power2d -> bin2D.bin
    percall = None
    if rank == 0 and not args.no_e2e and not args.no_extras and not pol:
        percall = per_call_api(cx, args, P, dtype)

    if rank != 0:
        return None

    # ---- roofline: per-stage device times from CUDA events between the stages (rank 0)
    s = 4 if args.dtype == "f32" else 8
    Npx = npix * npix
    nc = 3 if pol else 1
    keep = 0 if args.no_keep_maps else 1
    st_ms = profile_stages(capi, lambda: pipe.run_raw(seeds_pin.array[0], B, mode, flags, None), reps=5)
    if pipe.path == "fused":
        alg = pipeline_alg_bytes(npix, nc, s, keep, B)
        note = ("per-kernel bytes are the fused kernel's own minimum traffic; see roofline_pipeline for the BASELINE metric's fraction "
                "and roofline_true for the bytes the three kernels really move")
    else:
        alg = {}
        note = "cuFFT path"
    key = f"{'IQU' if pol else 'T'}{npix}_{args.dtype}"
    st_ms = {k: v for k, v in st_ms.items() if k in alg} if alg else st_ms
    roofline = cx.roofline(st_ms, alg, key, note) if alg else None
    pipe_bytes = (10 * s + 1) * Npx if not pol else (30 * s + 1) * Npx
    pipe_gbs = value / ws * pipe_bytes / 1e9
    roofline_pipeline = {"bound": "hbm", "algorithmic_bytes_per_map": pipe_bytes, "achieved": pipe_gbs, "peak": cx.peak,
                         "unit": "GB/s", "frac": pipe_gbs / cx.peak,
                         "note": "(10s+1)N per T map (SURVEY 8d: each 2-D FFT charged two passes); the taper RMW pass (2sN) is real traffic that the formula does not credit"}
    true_bytes = sum(alg.values()) / B if alg else None
    roofline_true = None
    if true_bytes:
        g2 = value / ws * true_bytes / 1e9
        roofline_true = {"bytes_per_map_minimal_of_the_three_kernels": true_bytes, "achieved": g2, "frac": g2 / cx.peak,
                         "unit": "GB/s", "note": "true HBM utilisation: the fused kernels' own minimal traffic x maps/s over the measured copy peak"}

    variants = None
    if ws == 1 and not args.no_extras:
        def timed(vmode, vflags, steps=4):
            pipe.run_raw(warm_pin.array, B, vmode, vflags, None)
            ms2, _ = cx.timed(lambda k: pipe.run_raw(seeds_pin.array[k % K], B, vmode, vflags, None), steps)
            return steps * B / (ms2 / 1e3)
        variants = {
            "philox_fullplane_noise_maps_per_s": timed(capi.NOISE_PHILOX, flags),
            "philox_hermitian_noise_maps_per_s": timed(capi.NOISE_PHILOX_HERMITIAN, flags),
            "without_storing_maps_maps_per_s": timed(mode, flags & ~capi.FLAG_KEEP_MAPS),
            "note": "same pipeline, 4 steps each: the reference's full complex-plane noise (4N normals/map) vs the "
                    "Hermitian half-plane draw (N normals/map); and without materialising the real-space maps in HBM"}
        pipe.reset_stats()

    cpu = None
    if ws == 1 and args.cpu_sample > 0:
        cpu = cpu_baseline_sample(npix, args.res, pol, not args.no_window, args.cpu_sample)

    with np.errstate(all="ignore"):
        bp_mean = S[:pipe.nbins] / max(N_stat, 1)
        w2 = float(np.mean(window ** 2)) if window is not None else 1.0
        ratio = float(np.nanmean(bp_mean / w2 / th.lCl("TT", binner.centers)))

    line = {
        "metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": ws, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, B, ws),
        "clocks": clocks, "e2e": e2e, "e2e_per_call_api": percall, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_pipeline": roofline_pipeline, "roofline_true": roofline_true,
        "stages": cx.stages(st_ms, alg), "cpu_baseline": cpu,
        "collective": {"what": "one ncclAllReduce(sum) of the packed float64 [N | SUM | CROSS] issued by liborphx.so (ox_pipeline_allreduce)",
                       "bytes": int(8 * (1 + pipe.dim + pipe.dim ** 2)), "ms": ar_ms, "nccl_version": cx.comm.nccl_version()},
        "check": {"stat_N": int(N_stat), "mean_binned_over_theory_TT": ratio},
        "variants": variants, "strong_scaling_1024_maps": strong, "pipeline_path": pipe.path,
        "device": capi.device_name(),
    }
    return line


def per_call_api(cx, args, P, dtype):
    """The reference's call sequence, one map at a time, with its signatures: get_map -> *taper -> power2d -> bin."""
    from orphics_b200 import maps
    capi = cx.capi
    shape, wcs, ps, window, binner = P["shape"], P["wcs"], P["ps"], P["window"], P["binner"]
    nmaps = 16
    mg1 = maps.MapGen(shape, wcs, ps, noise=args.noise, dtype=dtype, max_batch=1)
    fc1 = maps.FourierCalc(shape, wcs, dtype=dtype, max_batch=1)
    taper, w2 = maps.get_taper(shape, wcs)      # as the tutorials build it (a device-resident ndmap here)
    if dtype == np.float32:
        from orphics_b200 import enmap
        taper = enmap.to_device(np.asarray(taper).astype(np.float32), wcs=wcs)

    def one(seed):
        m = mg1.get_map(seed=int(seed))
        if window is not None:
            m = m * taper
        return binner.bin(fc1.power2d(m)[0])[1]
    one(SEED0)
    capi.synchronize()
    t0 = time.perf_counter()
    for i in range(nmaps):
        bp_last = np.asarray(one(SEED0 + i))
    dt1 = time.perf_counter() - t0
    info = getattr(maps, "PER_CALL_API_NOTE", None)
    return {"value": nmaps / dt1, "unit": "maps/s", "maps": nmaps,
            "note": info or ("drop-in calls one map at a time with the reference's signatures: MapGen.get_map -> * get_taper(...)[0] -> "
                             "FourierCalc.power2d -> bin2D.bin; the maps between the calls are device-resident ndmaps "
                             "(enmap.devmap), the bandpowers come back to the host every map"),
            "checksum": float(np.nansum(bp_last))}


# --------------------------------------------------------------------------- configs[2]: IQU
def run_config_iqu(cx, args):
    capi, mpi = cx.capi, cx.mpi
    B, K, npix = 16, 16, 2048
    P = build_pipeline(args, npix, True, B, np.float64, args.noise)
    pipe = P["pipe"]
    mode = capi.NOISE_MODES[args.noise]
    flags = pipe._flags(False, False, keep_maps=True)
    total = cx.ws * K * B
    _, tasks = mpi.mpi_distribute(total, cx.ws)
    seeds = capi.PinnedArray((K, B), np.int64)
    seeds.array[:] = (np.array(tasks[cx.rank], dtype=np.int64) + SEED0).reshape(K, B)
    out_pin = capi.PinnedArray((B, pipe.nspec, pipe.nbins), np.float64)
    for _ in range(3):
        pipe.run_raw(seeds.array[0], B, mode, flags, None)
    pipe.reset_stats()
    fin = lambda: pipe.allreduce(cx.comm)
    ms, launches = cx.timed(lambda k: pipe.run_raw(seeds.array[k], B, mode, flags, None), K, fin)
    pipe.reset_stats()
    dt = cx.wall(lambda k: pipe.run_raw(seeds.array[k], B, mode, flags, capi.ptr(out_pin.array)), K, fin)
    if cx.rank != 0:
        return None
    st_ms = profile_stages(capi, lambda: pipe.run_raw(seeds.array[0], B, mode, flags, None))
    alg = pipeline_alg_bytes(npix, 3, 8, 1, B)
    st_ms = {k: v for k, v in st_ms.items() if k in alg}
    value = total / (ms / 1e3)
    pb = (30 * 8 + 1) * npix * npix
    cpu = None
    if cx.ws == 1 and args.cpu_sample > 0:
        cpu = cpu_baseline_sample(npix, args.res, True, not args.no_window, 3)
        cpu["unit"] = "realisations/s"
    return {"workload": "configs[2]: IQU 2048x2048 (0.5') sims with the TEB rotation, taper, 6 binned auto/cross spectra per realisation",
            "value": value, "unit": "realisations/s", "ms_per_step": ms / K, "steps": K, "realisations_per_step_per_gpu": B,
            "n_gpus": cx.ws, "scaling": "weak", "dtype": "f64", "gpu_launches": int(launches), "pipeline_path": pipe.path,
            "e2e": {"value": total / dt, "unit": "realisations/s", "h2d_bytes_per_step": int(B * 8), "d2h_bytes_per_step": int(out_pin.array.nbytes)},
            "roofline": cx.roofline(st_ms, alg, "IQU2048_f64"),
            "roofline_pipeline": {"algorithmic_bytes_per_realisation": pb, "achieved": value / cx.ws * pb / 1e9, "unit": "GB/s",
                                  "frac": value / cx.ws * pb / 1e9 / cx.peak, "note": "(30s+1)N per realisation (SURVEY 8d)"},
            "stages": cx.stages(st_ms, alg), "cpu_baseline": cpu}


# --------------------------------------------------------------------------- configs[3] / configs[4]: quadratic estimator
def qe_alg_bytes(est, npix, s, nb):
    """Minimal HBM bytes of each stage of the hand-written estimator chain for nb realisations (s = bytes per real,
    half-plane complex array = s N bytes; the mean-field stack is complex128 whatever the arithmetic type)."""
    N = npix * npix
    nl = 3 if est == "TT" else 6
    nin = 1 if est == "TT" else 2
    return {"Q1 rows r2c": nin * 2 * s * N * nb, "Q2a cols fwd": nin * 2 * s * N * nb,
            "Q2b legs cols inv": (nin + nl) * s * N * nb, "Q3a rows c2r": 2 * nl * s * N * nb,
            "Q3b product rows r2c": (nl + 2) * s * N * nb, "Q4 cols fwd": 4 * s * N * nb,
            "finish div+meanfield": (2 * s + 2 * s + 16) * N * nb}


def run_config_qe(cx, args, est, npix, dtype_name, nb, total_real, label, scaling, cpu=True, shared=None):
    from orphics_b200 import maps, lensing, cosmology, stats
    capi, mpi = cx.capi, cx.mpi
    rdt = np.float32 if dtype_name == "f32" else np.float64
    s = 4 if dtype_name == "f32" else 8
    t_setup = time.perf_counter()
    shape, wcs = maps.rect_geometry(width_arcmin=npix * 0.5, px_res_arcmin=0.5)
    th = cosmology.default_theory()
    g = maps.Geometry.get(shape, wcs)
    if shared and shared.get("npix") == npix:
        qn, kw = shared["N"], shared["kw"]
    else:
        modl = g.modlmap()
        beam = maps.gauss_beam(modl, 1.5)                                   # beam 1.5', white noise 1 uK'
        n2d = np.zeros(shape) + (1.0 * np.pi / 180 / 60) ** 2
        tmask = maps.mask_kspace(shape, wcs, lmin=300, lmax=2000)
        kmask = maps.mask_kspace(shape, wcs, lmin=20, lmax=3500)
        kw = dict(noise2d=n2d, beam2d=beam, kmask=tmask, kmask_P=tmask, kmask_K=kmask, unlensed_equals_lensed=True)
        qn = None
        del modl
    q = lensing.qest(shape, wcs, th, pol=(est == "EB"), dtype=rdt, max_batch=nb, quadnorm=qn, **kw)
    if est not in q._plans:
        q._make_plan(est)
    setup_s = time.perf_counter() - t_setup
    if shared is not None:
        shared.update(npix=npix, N=q.N, kw=kw)
    N = npix * npix
    rng = np.random.RandomState(cx.rank)
    # inputs resident in HBM: observed-like real maps (T, or E and B)
    amp = 60.0 if est == "TT" else 3.0
    xin = capi.PinnedArray((nb, npix, npix), rdt)
    xin.array[:] = (rng.standard_normal((nb, npix, npix)) * amp).astype(rdt)
    X = capi.DeviceBuffer(xin.array.nbytes).upload(xin.array)
    yin = Y = None
    if est == "EB":
        yin = capi.PinnedArray((nb, npix, npix), rdt)
        yin.array[:] = (rng.standard_normal((nb, npix, npix)) * 1.0).astype(rdt)
        Y = capi.DeviceBuffer(yin.array.nbytes).upload(yin.array)
    out = capi.DeviceBuffer(nb * N * 2 * s)
    h = q._plans[est][0]
    import ctypes as C

    def step(k=0):
        capi.check(capi.lib.ox_qe_reconstruct(h, C.c_void_p(X.ptr), C.c_void_p(Y.ptr) if Y else None, capi.OX_DEVICE, nb, 0, 1, 1,
                                              C.c_void_p(out.ptr), capi.OX_DEVICE))
    if scaling == "strong":
        mine = len(mpi.mpi_distribute(total_real, cx.ws)[1][cx.rank])
    else:
        mine = total_real
    K = max(1, mine // nb)
    done = K * nb * cx.ws
    for _ in range(3):
        step()
    q.reset_meanfield(est)
    fin = lambda: q.allreduce_meanfield(est, cx.comm)       # ONE ncclAllReduce of [stack | count] (stats.py:1227-1228)
    ms, launches = cx.timed(step, K, fin)
    count = q.meanfield_count(est) if cx.rank == 0 else None
    ar_ms = None
    if cx.ws > 1:
        ar_ms, _ = cx.timed(lambda k: fin(), 1)
    # end to end through the public API with host buffers: pinned maps in -> kappa maps in pinned memory out
    kout = capi.PinnedArray((nb, npix, npix), rdt)
    Ke = min(K, 2)
    q.kappa_from_maps(est, xin.array, None if yin is None else yin.array, out=kout.array)
    dt = cx.wall(lambda k: q.kappa_from_maps(est, xin.array, None if yin is None else yin.array, accumulate_meanfield=True,
                                             out=kout.array), Ke)
    e2e_val = Ke * nb * cx.ws / dt
    if cx.rank != 0:
        return None
    st_ms = profile_stages(capi, step)
    alg = qe_alg_bytes(est, npix, s, nb)
    st_ms = {k: v for k, v in st_ms.items() if k in alg}
    value = done / (ms / 1e3)
    pb = (38 if est == "TT" else 76) * s * N
    res = {"workload": label, "value": value, "unit": "realisations/s", "ms_per_step": ms / K, "steps": K,
           "realisations_per_step_per_gpu": nb, "realisations": done, "n_gpus": cx.ws, "scaling": scaling, "dtype": dtype_name,
           "gpu_launches": int(launches), "path": q.path(est), "setup_seconds": setup_s,
           "meanfield_count_after_allreduce": count,
           "collective": {"what": "one ncclAllReduce(sum) of the packed float64 [mean-field stack | count] (ox_qe_meanfield_allreduce)",
                          "bytes": int(8 * (2 * (npix * (npix // 2 + 1)) + 1)), "ms": ar_ms},
           "e2e": {"value": e2e_val, "unit": "realisations/s", "h2d_bytes_per_step": int(xin.array.nbytes * (2 if yin is not None else 1)),
                   "d2h_bytes_per_step": int(kout.array.nbytes),
                   "note": "qest.kappa_from_maps: observed maps from pinned host memory -> kappa maps in pinned host memory "
                           "(+ mean-field accumulate on the device); PCIe-bound: the call pipelines the realisations (upload of r+1 and "
                           "download of r-1 on their own streams while r is reconstructed)"},
           "roofline": cx.roofline(st_ms, alg, f"QE_{est}{npix}_{dtype_name}"),
           "roofline_pipeline": {"algorithmic_bytes_per_realisation": pb, "achieved": value / cx.ws * pb / 1e9, "unit": "GB/s",
                                 "frac": value / cx.ws * pb / 1e9 / cx.peak,
                                 "note": ("38 s N" if est == "TT" else "76 s N") + " per realisation (SURVEY 8d: cuFFT-chain accounting, each 2-D FFT two passes)"},
           "stages": cx.stages(st_ms, alg), "cpu_baseline": None}
    if est == "TT" and not args.no_extras:
        res["sim_to_kappa_on_device"] = sim_to_kappa_chain(cx, args, q, shape, wcs, th, npix)
    if cpu and cx.ws == 1 and args.cpu_sample > 0:
        res["cpu_baseline"] = qe_cpu_baseline(q, est, npix, xin.array[0], None if yin is None else yin.array[0])
    for b in (X, Y, out):
        if b is not None:
            b.free()
    return res


def sim_to_kappa_chain(cx, args, q, shape, wcs, th, npix, nreal=6):
    """configs[3] end to end as tutorials/tt_verification.ipynb:597-617 runs it, one realisation at a time through the
    reference signatures, every map device-resident: FlatLensingSims.get_sim (unlensed GRF -> kappa GRF -> phi -> Taylens
    order 5 -> beam -> + noise GRF) -> FourierCalc.power2d -> qest.kappa_from_map (+ mean-field accumulate)."""
    from orphics_b200 import maps, lensing
    capi = cx.capi
    t0 = time.perf_counter()
    sims = lensing.FlatLensingSims(shape, wcs, th, 1.5, 1.0, pol=False, noise="philox")
    fc = maps.FourierCalc(shape, wcs)
    setup_s = time.perf_counter() - t0

    def one(i, parts=None):
        obs = sims.get_sim(seed_cmb=3 * i, seed_kappa=3 * i + 1, seed_noise=3 * i + 2, lens_order=5)
        if parts is not None:
            capi.synchronize()
            parts.append(time.perf_counter())
        _, kT, _ = fc.power2d(obs)
        rec = q.kappa_from_map("TT", kT, alreadyFTed=True, accumulate_meanfield=True)
        return rec
    one(0)
    capi.synchronize()
    t0 = time.perf_counter()
    for i in range(1, nreal + 1):
        rec = one(i)
    capi.synchronize()
    dt = time.perf_counter() - t0
    parts = [time.perf_counter()]
    one(99, parts)
    capi.synchronize()
    parts.append(time.perf_counter())
    chk = float(np.asarray(rec[:4, :4]).sum())
    return {"value": nreal / dt, "unit": "realisations/s", "realisations": nreal, "setup_seconds": setup_s,
            "ms_get_sim": 1e3 * (parts[1] - parts[0]), "ms_power2d_and_estimator": 1e3 * (parts[2] - parts[1]), "checksum": chk,
            "note": "per-call API, device-resident maps between the calls; Taylens order 5 = 14 half-plane inverse transforms per map"}


def qe_cpu_baseline(q, est, npix, x, y):
    """One realisation of the numpy oracle's kappa_from_map chain (filters handed in, set-up excluded) with the
    FFTs on all host cores."""
    import scipy.fft
    from oracle import qe_np, maps_np as omaps
    so, wo = omaps.rect_geometry(width_arcmin=npix * 0.5, px_res_arcmin=0.5)
    Nn = q.N
    tables = {"WXY_" + est: np.asarray(Nn.WXY(est)), "WY_" + est[1] * 2: np.asarray(Nn.WY(est[1] * 2))}
    qo = qe_np.qest_from_tables(so, wo, tables, {est: np.asarray(Nn.AL[est])}, None if Nn.fmaskK is None else np.asarray(Nn.fmaskK))
    cores = host_cores()
    x64 = np.asarray(x, dtype=np.float64)
    y64 = None if y is None else np.asarray(y, dtype=np.float64)
    with scipy.fft.set_workers(cores):
        t0 = time.perf_counter()
        if est == "TT":
            qo.kappa_from_map("TT", x64, returnFt=True)
        else:
            qo.kappa_from_map("EB", None, x64, y64, returnFt=True)
        dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "realisations/s", "cores": cores, "kind": "port",
            "sample": f"1 realisation of the {est} estimator at {npix}^2 through the numpy oracle (qe_np.kappa_from_map, returnFt), "
                      f"scipy.fft on {cores} threads, numpy elementwise passes on 1; filters and A_L handed in (set-up excluded)"}


def run_ours(args):
    cx = Ctx()
    line = run_headline(cx, args)
    want = [] if args.configs == "none" else (["2", "3", "4"] if args.configs == "all" else args.configs.split(","))
    configs = {}
    t0 = time.perf_counter()
    if "2" in want and not args.pol:
        r = run_config_iqu(cx, args)
        if r:
            configs["configs[2]"] = r
    if "3" in want:
        r = run_config_qe(cx, args, "TT", 4096, "f64", 8, 512, "configs[3]: TT Hu-Okamoto quadratic estimator on 4096x4096 (0.5') observed maps "
                          "(beam 1.5', noise 1 uK', l in (300,2000)), 512 realisations sharded over the GPUs, kappa_hat(l) + mean-field "
                          "accumulate, one mean-field all-reduce", "strong")
        if r:
            configs["configs[3]"] = r
    if "4" in want:
        shared = {}
        for dt_name in ("f64", "f32"):
            r = run_config_qe(cx, args, "EB", 8192, dt_name, 2, 16, f"configs[4]: EB quadratic estimator at 8192x8192 (0.5') {dt_name}, "
                              "16 realisations per GPU from real E/B maps, kappa_hat(l) + mean-field accumulate, one mean-field all-reduce",
                              "weak", cpu=(dt_name == "f64"), shared=shared)
            if r:
                if dt_name == "f32" and "configs[4] fp64" in configs:
                    base = configs["configs[4] fp64"]["cpu_baseline"]
                    r["cpu_baseline"] = dict(base, note="the reference has no float32 path: fp64 CPU figure") if base else None
                configs[f"configs[4] fp{dt_name[1:]}"] = r
    if cx.rank == 0:
        line["configs"] = configs
        line["configs_seconds"] = time.perf_counter() - t0
        print(json.dumps(line))
    cx.comm.free()
    if cx.ws > 1:
        cx.dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    import builtins
    _print = builtins.print

    def print_json(*a, **k):
        _print(*a, **k, file=real_stdout)
        real_stdout.flush()
    builtins.print = print_json
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
