"""Top SASS instructions by stall samples for one kernel of an .ncu-rep, grouped coarsely.
Usage: python tools/ncu_hot.py rep kernel_regex [n]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("kernel:", rows[0][1][:90], " total samples", tot, " instrs", len(body))
stallcols = [h for h in hdr if h.startswith("stall_")] 
byop = collections.Counter()
for r in body:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    byop[op.split(".")[0]] += int(r[ix["# Samples"]])
print("by opcode:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in byop.most_common(14)))
top = sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:n]
for r in top:
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stallcols if r[ix[h]].isdigit()), reverse=True)[:2]
    print(f"{100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {st}")
