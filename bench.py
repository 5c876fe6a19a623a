#!/usr/bin/env python
"""Benchmark of the north-star path: maps/sec through sim -> FFT -> power2d -> bin2D at
2048^2 fp64 (BASELINE.json metric; workload = configs[1], T-only GRF sims from the CAMB
lensed TT spectrum at 0.5 arcmin, 72 bandpowers), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, liborphx.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle) on host cores

A step = one batch of --batch maps per GPU through ox_pipeline_run.  On power-of-two maps the
pipeline is three hand-written kernels (orphics_b200/csrc/ox_fused.cu):
  K_A Philox noise x covsqrt -> inverse FFT along y        (writes the transposed half plane)
  K_B inverse FFT along x -> real map (stored) x taper -> forward FFT along x
  K_C forward FFT along y -> |k|^2 -> deterministic annular binning -> bandpowers
followed by the Statistics triple; other sizes use cuFFT passes with hand-written kernels
around them (ORPHX_PIPELINE=cufft forces that path).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EDGES = np.arange(100, 3000, 40.0)      # tutorials/demo-grf.ipynb:159
SEED0 = 1000                            # sim i uses seed 1000+i
METRIC = "maps/sec sim->FFT->power2d->bin2D at 2048^2 fp64"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128, help="timed steps (128 x 64 maps ~ 0.7 s on one B200)")
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
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--pol", action="store_true", help="IQU sims, 6 spectra (configs[2])")
    ap.add_argument("--no-window", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=24, help="maps timed for cpu_baseline: ~10 s on one host core (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


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


def cpu_baseline_sample(args, nmaps):
    """1 core, nmaps maps of the same workload (set-up excluded, 1 warm-up map)."""
    _oracle_setup(args.npix, args.res, args.pol, not args.no_window)
    _oracle_one(SEED0)
    t0 = time.perf_counter()
    for i in range(nmaps):
        _oracle_one(SEED0 + i)
    dt = time.perf_counter() - t0
    return {"value": nmaps / dt, "unit": "maps/s", "cores": 1, "kind": "port",
            "sample": f"{nmaps} maps of {args.npix}^2 {'IQU' if args.pol else 'T'} through the numpy oracle "
                      f"(MapGen.get_map -> taper -> power2d -> bin2D.bin), 1 process / 1 thread, set-up excluded"}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (numpy oracle; the reference
    itself cannot be imported: pixell is absent) on all host cores, one process per core
    with realisations split as the reference does under MPI (mpi.py:78-91)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    ncpu = os.cpu_count() or 1
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        pass
    mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    per_proc_gb = 1.5 * (args.npix / 2048.0) ** 2 * (3 if args.pol else 1)
    P = max(1, int(min(ncpu, 64, mem_gb * 0.5 / per_proc_gb)))
    ctx = mp.get_context("fork")
    window = not args.no_window
    with ctx.Pool(P, initializer=_oracle_setup, initargs=(args.npix, args.res, args.pol, window)) as pool:
        step_maps = P                                             # one map per core per step
        for w in range(args.warmup):
            pool.map(_oracle_one, [SEED0 + i for i in range(step_maps)], chunksize=1)
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_oracle_one, [SEED0 + k * step_maps + i for i in range(step_maps)], chunksize=1)
        dt = time.perf_counter() - t0
    value = args.steps * step_maps / dt
    sample = (f"{step_maps} maps per step ({P} processes x 1 map, 1 FFT thread each) of {args.npix}^2 "
              f"{'IQU' if args.pol else 'T'}; numpy-oracle restatement of MapGen.get_map -> taper -> power2d -> bin2D.bin")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, step_maps, 1),
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
            "parallelism": f"realisations sharded over {ngpu} GPU(s) (mpi_distribute rule), one all-reduce of the Statistics triple",
            "l2": "working set per step (batch x 67 MB of maps+Fourier planes) >> 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------- this repo
class _DevView:
    """__cuda_array_interface__ holder so torch can wrap a liborphx device buffer in place."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def run_ours(args):
    from orphics_b200 import _capi, maps, stats, mpi, cosmology
    rank, local, ws = mpi.init_process_group()
    dist = torch = None
    if ws > 1:
        import torch
        import torch.distributed as dist
    _capi.require_device()
    _capi.set_device(local)
    npix, B, K, W = args.npix, args.batch, args.steps, args.warmup
    pol = args.pol
    dtype = np.float32 if args.dtype == "f32" else np.float64
    shape, wcs = maps.rect_geometry(width_arcmin=npix * args.res, px_res_arcmin=args.res, pol=pol)
    th = cosmology.default_theory()
    g = maps.Geometry.get(shape, wcs)
    modl = g.modlmap()
    ps = cosmology.power_from_theory(np.arange(0, modl.max() + 1, 1.0), th, lensed=True, pol=pol)
    mg = maps.MapGen(shape, wcs, ps, noise=args.noise, dtype=dtype, max_batch=B)
    fc = maps.FourierCalc(shape, wcs, dtype=dtype, max_batch=B)
    binner = stats.bin2D(modl, EDGES, geometry=g)
    window = None if args.no_window else np.asarray(maps.get_taper(shape, wcs)[0])
    pipe = maps.SimPipeline(mg, fc, binner, window=window)
    mode = _capi.NOISE_MODES[args.noise]
    # the real-space maps are materialised in HBM (SURVEY 8d anti-short-circuit rule) unless asked not to
    flags = pipe._flags(False, False, keep_maps=not args.no_keep_maps)

    # shard: the job is ws*K*B realisations, contiguous blocks per rank (mpi.py:78-91)
    total = ws * K * B
    _, tasks = mpi.mpi_distribute(total, ws)
    my = np.array(tasks[rank], dtype=np.int64) + SEED0
    seeds_pin = _capi.PinnedArray((K, B), np.int64)
    seeds_pin.array[:] = my.reshape(K, B)
    warm_pin = _capi.PinnedArray((B,), np.int64)
    warm_pin.array[:] = np.arange(B) + 7
    out_pin = _capi.PinnedArray((B, pipe.nspec, pipe.nbins), np.float64)

    def barrier():
        _capi.synchronize()
        if ws > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stat_tensors = None
    if ws > 1:
        n_p, s_p, c_p, d = pipe.stats_pointers()
        stat_tensors = [torch.as_tensor(_DevView(n_p, (1,), "<i8"), device=f"cuda:{local}"),
                        torch.as_tensor(_DevView(s_p, (d,), "<f8"), device=f"cuda:{local}"),
                        torch.as_tensor(_DevView(c_p, (d, d), "<f8"), device=f"cuda:{local}")]

    def reduce_stats():
        if ws > 1:
            for t in stat_tensors:     # N, SUM, CROSS (stats.py:1215-1217) over NCCL, in place
                dist.all_reduce(t, op=dist.ReduceOp.SUM)

    # ---- device-resident throughput
    for _ in range(max(W, 3)):
        pipe.run_raw(warm_pin.array, B, mode, flags, None)
    reduce_stats()
    pipe.reset_stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timer = _capi.Timer()
    l0 = _capi.launch_count()
    timer.start()
    for k in range(K):
        pipe.run_raw(seeds_pin.array[k], B, mode, flags, None)
    reduce_stats()
    timer.stop()
    ms = timer.elapsed_ms()
    launches = _capi.launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if ws > 1:
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    N_stat, S, Cm = pipe.stats()
    value = total / (ms / 1e3)

    # ---- end to end through the Python API with host buffers: pinned seeds in, bandpowers out, every step
    e2e = None
    if not args.no_e2e:
        pipe.reset_stats()
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            pipe.run_raw(seeds_pin.array[k], B, mode, flags, _capi.ptr(out_pin.array))   # D2H + sync inside
        _capi.synchronize()
        dt = time.perf_counter() - t0
        if ws > 1:
            t = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total / dt, "unit": "maps/s", "h2d_bytes_per_step": int(B * 8),
               "d2h_bytes_per_step": int(out_pin.array.nbytes),
               "note": "SimPipeline.run_raw: seeds from pinned host memory -> bandpowers in pinned host memory each step; "
                       "the noise itself is drawn on the device (Philox), as the reference draws it in-process"}

    # ---- the reference's per-call API with host arrays on both sides of every call (rank 0, a few maps):
    # MapGen.get_map -> numpy map on the host -> x taper -> FourierCalc.power2d -> bin2D.bin
    percall = None
    if rank == 0 and not args.no_e2e and not pol:
        nmaps = 8
        mg1 = maps.MapGen(shape, wcs, ps, noise=args.noise, dtype=dtype, max_batch=1)
        fc1 = maps.FourierCalc(shape, wcs, dtype=dtype, max_batch=1)

        def one(seed):
            m = mg1.get_map(seed=int(seed))
            if window is not None:
                m = m * window
            return binner.bin(fc1.power2d(m)[0])[1]
        one(SEED0)
        _capi.synchronize()
        t0 = time.perf_counter()
        for i in range(nmaps):
            bp_last = one(SEED0 + i)
        dt1 = time.perf_counter() - t0
        mb = npix * npix * np.dtype(dtype).itemsize
        percall = {"value": nmaps / dt1, "unit": "maps/s", "maps": nmaps,
                   "h2d_bytes_per_map": int(2 * mb), "d2h_bytes_per_map": int(4 * mb),
                   "note": "drop-in calls one map at a time with pageable numpy arrays, as the reference's signatures "
                           "require: get_map returns the map (D2H), power2d takes it (H2D) and returns p2d and the "
                           "complex k-map (D2H), bin2D.bin takes p2d (H2D); PCIe- and host-numpy-bound by construction, "
                           "the batched SimPipeline call (e2e) is the product path"}
        del mg1, fc1

    if rank != 0:
        if ws > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: per-stage device times from CUDA events between the stages (rank 0)
    s = 4 if args.dtype == "f32" else 8
    Npx = npix * npix
    nc = 3 if pol else 1
    keep = 0 if args.no_keep_maps else 1
    if pipe.path == "fused":
        # three hand-written kernels; bytes = what each one must move (half planes are s*N bytes)
        names = {"sim_fill": "K_A sim+col_ifft", "cufft_inverse": "K_B row_c2r+taper+r2c", "power_bin": "K_C col_fft+power+bin",
                 "statistics": "statistics"}
        alg = {"K_A sim+col_ifft": nc * s * Npx, "K_B row_c2r+taper+r2c": nc * (2 + keep) * s * Npx,
               "K_C col_fft+power+bin": nc * s * Npx + Npx, "statistics": 0}
        ours_keys = ["K_A sim+col_ifft", "K_B row_c2r+taper+r2c", "K_C col_fft+power+bin"]
    else:
        names = {k2: k2 for k2 in pipe.STAGES}
        alg = {"sim_fill": nc * s * Npx, "cufft_inverse": nc * 4 * s * Npx, "window": nc * 2 * s * Npx if window is not None else 0,
               "cufft_forward": nc * 4 * s * Npx, "power_bin": nc * s * Npx + Npx, "statistics": 0}
        ours_keys = ["sim_fill", "window", "power_bin"]
    reps = 5
    acc = {k: 0.0 for k in alg}
    for r in range(reps):
        st = pipe.profile(seeds_pin.array[r % K], B, mode, flags)
        for k2, v in st.items():
            if k2 in names:
                acc[names[k2]] += v / reps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    stages = {}
    for k2 in alg:
        gbs = alg[k2] * B / (acc[k2] * 1e-3) / 1e9 if acc[k2] > 0 and alg[k2] else None
        stages[k2] = {"ms_per_launch": acc[k2], "algorithmic_bytes_per_map": alg[k2], "achieved_gbs": gbs,
                      "frac": gbs / peak if gbs else None}
    ours = {k2: stages[k2] for k2 in ours_keys}
    dom = max(ours, key=lambda k2: ours[k2]["ms_per_launch"])
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per map from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if npix == 2048 and args.dtype == "f64" and not pol and dom in tj["kernels"]:
            traffic = tj["kernels"][dom]["dram_bytes_per_map"] * B
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom] * B,
                "note": "per-kernel bytes are the fused kernel's own minimum traffic; the kernels are bound by shared-memory "
                        "bandwidth and FP64 issue (FFT butterflies, Box-Muller), not by HBM -- see roofline_pipeline for the "
                        "BASELINE metric's fraction"}
    pipe_bytes = (10 * s + 1) * Npx * nc if not pol else (30 * s + 1) * Npx
    pipe_gbs = value / ws * pipe_bytes / 1e9
    roofline_pipeline = {"bound": "hbm", "algorithmic_bytes_per_map": pipe_bytes, "achieved": pipe_gbs, "peak": peak,
                         "unit": "GB/s", "frac": pipe_gbs / peak,
                         "note": "(10s+1)N per T map (SURVEY 8d); the taper RMW pass (2sN) is real traffic that the formula does not credit"}

    variants = None
    if ws == 1 and not args.no_extras:
        def timed(vmode, vflags, steps=4):
            pipe.run_raw(warm_pin.array, B, vmode, vflags, None)
            _capi.synchronize()
            tm = _capi.Timer()
            tm.start()
            for k in range(steps):
                pipe.run_raw(seeds_pin.array[k % K], B, vmode, vflags, None)
            tm.stop()
            return steps * B / (tm.elapsed_ms() / 1e3)
        variants = {
            "philox_fullplane_noise_maps_per_s": timed(_capi.NOISE_PHILOX, flags),
            "philox_hermitian_noise_maps_per_s": timed(_capi.NOISE_PHILOX_HERMITIAN, flags),
            "without_storing_maps_maps_per_s": timed(mode, flags & ~_capi.FLAG_KEEP_MAPS),
            "note": "same pipeline, 4 steps each: the reference's full complex-plane noise (4N normals/map) vs the "
                    "Hermitian half-plane draw (N normals/map); and without materialising the real-space maps in HBM"}
        pipe.reset_stats()

    cpu = None
    if ws == 1 and args.cpu_sample > 0:
        cpu = cpu_baseline_sample(args, args.cpu_sample)

    ratio = None
    bp_mean = S[:pipe.nbins] / max(N_stat, 1)
    with np.errstate(all="ignore"):
        w2 = float(np.mean(window ** 2)) if window is not None else 1.0
        ratio = float(np.nanmean(bp_mean / w2 / th.lCl("TT", binner.centers)))

    line = {
        "metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": ws, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, B, ws),
        "clocks": clocks, "e2e": e2e, "e2e_per_call_api": percall, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_pipeline": roofline_pipeline, "stages": stages, "cpu_baseline": cpu,
        "check": {"stat_N": int(N_stat), "mean_binned_over_theory_TT": ratio},
        "variants": variants, "pipeline_path": pipe.path,
        "device": _capi.device_name(),
    }
    print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


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
