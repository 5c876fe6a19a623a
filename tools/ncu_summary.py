"""Summarise an .ncu-rep (read here, no GPU needed): per kernel duration, DRAM bytes, pipe
utilisation, occupancy and the top stall reasons.  Usage: python tools/ncu_summary.py rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum"]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
keys += ["smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if pat and not pat.search(name):
        continue
    print("====", name[:100])
    for k in keys:
        if k in idx:
            print(f"  {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
    top = sorted(((float(r[idx[h]]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall if r[idx[h]]), reverse=True)[:6]
    print("  stalled warps per issued instruction (top reasons): " + ", ".join(f"{n} {v:.2f}" for v, n in top))
