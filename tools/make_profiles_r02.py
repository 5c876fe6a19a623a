"""Regenerates the committed round-2 ncu evidence under profiles/ from the scratch captures in gpurun_out/
(tools/prof_run_r02.sh on the GPU box, plus the row-pass captures of tools/r02_run*.sh): per-kernel summaries, the launch
list of the bench command and the per-launch DRAM traffic bench.py quotes as roofline.traffic.
Usage: python tools/make_profiles_r02.py [output directory, default profiles/]"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")   # (on the GPU box: a directory under gpurun_out/)
os.makedirs(PR, exist_ok=True)
BENCH = "python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 --no-extras"


def summary(rep, pat=None):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep] + ([pat] if pat else [])
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    unit = dict(zip(hdr, units))
    return rows[2:], ix, unit


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def dram(rep, match):
    """(read, write, duration_us, name) of the first kernel of the report whose name contains `match`."""
    rows, ix, unit = raw_rows(rep)
    for r in rows:
        name = r[ix["Kernel Name"]]
        if match in name:
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], unit["dram__bytes_read.sum"])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], unit["dram__bytes_write.sum"])
            dur = float(r[ix["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit["gpu__time_duration.sum"]]
            return rd, wr, dur, name
    return None


traffic, notes = {}, {}


def add(key, stage, rep, match, what):
    p = os.path.join(GO, rep)
    if not os.path.exists(p):
        print("missing", rep)
        return
    d = dram(p, match)
    if d is None:
        print("no kernel matching", match, "in", rep)
        return
    traffic.setdefault(key, {})[stage] = d[0] + d[1]
    notes.setdefault(key, {})[stage] = {"dram_read": d[0], "dram_write": d[1], "duration_us_under_ncu": d[2], "kernel": d[3][:160], "capture": what}


# ---- the three pipeline kernels at batch 64 (as bench.py times them)
T = os.path.join(GO, "prof_r02_T.ncu-rep")
if os.path.exists(T):
    open(os.path.join(PR, "r02_ncu_fused_kernels_summary.txt"), "w").write(
        "ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'fused_(sim_col|row_tma|col_bin)_kernel' "
        f"--launch-skip 9 --launch-count 3 : {BENCH} --configs none\n"
        "(64 maps of 2048^2 fp64 per launch, the batch bench.py times; under ncu the kernels run serialised with cold caches)\n\n" + summary(T))
    for stage, m in (("K_A sim+col_ifft", "fused_sim_col_kernel"), ("K_B row_c2r+taper+r2c", "fused_row_tma_kernel"),
                     ("K_C col_fft+power+bin", "fused_col_bin_kernel")):
        add("T2048_f64", stage, "prof_r02_T.ncu-rep", m, "prof_r02_T: 64 maps per launch")
txt = []
for rep, stage, key, m, what in (("prof_r02_IQU.ncu-rep", "K_A sim+col_ifft", "IQU2048_f64", "fused_sim_col_kernel", "configs[2]: 16 IQU realisations per launch"),
                                 ("prof_r02_TT.ncu-rep", "Q3a rows c2r", "QE_TT4096_f64", "fused_row_tma_kernel", "configs[3]: 8 realisations of 4096^2 per launch"),
                                 ("prof_r02_EB64.ncu-rep", "Q3a rows c2r", "QE_EB8192_f64", "fused_row_kernel", "configs[4] fp64: 2 realisations of 8192^2 per launch"),
                                 ("prof_r02_EB32.ncu-rep", "Q2b legs cols inv", "QE_EB8192_f32", "fused_col_kernel", "configs[4] fp32: 2 realisations of 8192^2 per launch")):
    add(key, stage, rep, m, what)
    p = os.path.join(GO, rep)
    if os.path.exists(p):
        txt.append(f"---- {what}: dominant stage '{stage}'\n" + summary(p))
if txt:
    open(os.path.join(PR, "r02_ncu_config_kernels_summary.txt"), "w").write(
        "ncu --set full --clock-control none of the dominant kernel of each other BASELINE configuration, one launch at the batch size\n"
        f"bench.py uses ({BENCH} --configs N; tools/prof_run_r02.sh)\n\n" + "\n".join(txt))
if traffic:
    traffic["_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at the bench's batch size, ncu --set full "
                          "(profiles/r02_ncu_fused_kernels_summary.txt, r02_ncu_config_kernels_summary.txt); details below")
    traffic["_details"] = notes
    json.dump(traffic, open(os.path.join(PR, "r02_traffic.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in traffic.items() if not k.startswith("_")}, indent=1))

# ---- the row pass, step by step
steps = (("prof_r02d.ncu-rep", "fused_row_kernel", "one tile per CTA, LDGSTS tile loads (round 1 kernel + planes-fastest order + separable window)"),
         ("prof_r02c.ncu-rep", "fused_row_tma", "persistent TMA kernel, first version (slot bookkeeping in shared memory, tables through L1)"),
         ("prof_r02e.ncu-rep", "fused_row_tma", "static schedule + window profile / twiddles in shared memory"),
         ("prof_r02g.ncu-rep", "fused_row_tma", "+ parity-split twiddle table, compact second-stage table, flat-window skip"),
         ("prof_r02_planes.ncu-rep", "fused_row_tma", "same kernel, single tiles dealt planes-fastest, captured in the session of the next entry"),
         ("prof_r02_pairs.ncu-rep", "fused_row_tma", "+ PAIRS of adjacent row tiles per CTA (the two 64 B halves of a 128 B line meet in L2): final"),
         ("prof_r02f.ncu-rep", "fused_row_w32", "experiment: one warp per row, two radix-32 stages (254 registers, 8 warps/SM) -- rejected"))
out = ["ncu --set full --clock-control none -k regex:<row kernel> --launch-skip 3 --launch-count 1 : python bench.py --steps 1 --warmup 3 --no-e2e "
       "--cpu-sample 0 --batch 64 --no-extras --configs none", "(64 maps of 2048^2 fp64 per launch; algorithmic bytes 3 s N x 64 = 6.44 GB)", ""]
for rep, pat, what in steps:
    p = os.path.join(GO, rep)
    if os.path.exists(p):
        out.append(f"######## {what}")
        out.append(summary(p, pat))
if len(out) > 3:
    open(os.path.join(PR, "r02_ncu_row_pass.txt"), "w").write("\n".join(out))

# ---- launch lists
for csvname, outname, title, tail in (("launches_r02.csv", "r02_ncu_launches_bench_batch64.txt", f"{BENCH} --configs none", 5),
                                      ("launches_r02_qe.csv", "r02_qe_launches.txt", f"{BENCH} --configs 3   (--launch-skip 200 -c 120)", 9)):
    p = os.path.join(GO, csvname)
    if not os.path.exists(p):
        continue
    lrows = [r for r in csv.reader(open(p)) if len(r) > 10 and r[0].isdigit()]
    by = collections.OrderedDict()
    for r in lrows:
        by.setdefault(r[4], []).append(float(r[-1]) / 1e3)
    step = lrows[-tail:]
    tot = sum(float(r[-1]) for r in step) / 1e3
    o = [f"ncu --metrics gpu__time_duration.sum --clock-control none : {title}",
         "(cold-cache, serialised launches: compare SHARES with the CUDA-event stage times of the bench line)", ""]
    for r in step:
        o.append(f"{float(r[-1]) / 1e3:10.1f} us  {100 * float(r[-1]) / 1e3 / tot:5.1f}%  {r[4][:140]}")
    o += [f"{tot:10.1f} us  total of the last {tail} launches (one step)", "", "all launches:"]
    for k, v in by.items():
        o.append(f"{len(v):4d} x {sum(v) / len(v):10.1f} us  {k[:140]}")
    open(os.path.join(PR, outname), "w").write("\n".join(o) + "\n")
    print("\n".join(o[:14]))
