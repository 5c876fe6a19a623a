"""Regenerates the committed ncu evidence under profiles/ from the scratch captures in gpurun_out/
(tools/prof_run.sh on the GPU box): per-kernel summary of the --set full capture, the launch list of
the bench command and the per-map DRAM traffic bench.py quotes.  Usage: python tools/make_profiles.py r01"""
import collections, csv, io, json, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}c.ncu-rep")
BATCH = 16

summary = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "fused_"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_fused_kernels_summary.txt"), "w").write(
    f"ncu --set full --clock-control none --import-source on -k regex:fused_ --launch-skip 9 --launch-count 3 : "
    f"python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch {BATCH} --no-extras\n"
    f"({BATCH} maps of 2048^2 fp64 per launch; under ncu the kernels run serialised with cold caches)\n\n" + summary)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
names = {"fused_sim_col": "K_A sim+col_ifft", "fused_row": "K_B row_c2r+taper+r2c", "fused_col_bin": "K_C col_fft+power+bin"}
traffic = {"source": f"profiles/{tag}_ncu_fused_kernels_summary.txt (ncu --set full, bench.py --batch {BATCH}, 2048^2 fp64, maps materialised)",
           "kernels": {}}
unit = {h: u for h, u in zip(rows[0], rows[1])}


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


for r in rows[2:]:
    for key, label in names.items():
        if key + "_kernel" in r[ix["Kernel Name"]]:
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], unit["dram__bytes_read.sum"])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], unit["dram__bytes_write.sum"])
            dur = float(r[ix["gpu__time_duration.sum"]])
            du = unit["gpu__time_duration.sum"]
            dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(du, 1)
            traffic["kernels"][label] = {"dram_bytes_per_map": (rd + wr) / BATCH, "dram_read_per_map": rd / BATCH,
                                         "dram_write_per_map": wr / BATCH, "duration_us_per_map_under_ncu": dur_us / BATCH}
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)

# launch list
lrows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv"))) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
for r in lrows:
    by.setdefault(r[4], []).append(float(r[-1]) / 1e3)
step = [r for r in lrows[-6:]]
out = [f"ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 --batch {BATCH} --no-extras",
       f"(cold-cache, serialised launches: compare SHARES; {BATCH} maps of 2048^2 fp64 per step)", ""]
tot = sum(float(r[-1]) for r in step) / 1e3
for r in step:
    out.append(f"{float(r[-1]) / 1e3:10.1f} us  {100 * float(r[-1]) / 1e3 / tot:5.1f}%  {r[4][:120]}")
out += [f"{tot:10.1f} us  total of the last step", "", "all launches:"]
for k, v in by.items():
    out.append(f"{len(v):4d} x {sum(v) / len(v):10.1f} us  {k[:120]}")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_bench_batch{BATCH}.txt"), "w").write("\n".join(out) + "\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_bench_batch{BATCH}.txt")).read())
print(json.dumps(traffic, indent=1))
