python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 32 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T pipeline', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"
python bench.py --pol --batch 16 --steps 16 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('IQU', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"
for cfg in "4096 8 f64 TT" "2048 16 f64 TT" "4096 8 f32 TT" "8192 2 f64 TT"; do python tools/bench_qe.py $cfg 2>&1 | tail -1 | cut -c1-160; done
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 120 -c 40 --csv --log-file gpurun_out/qe_launches.csv python tools/bench_qe.py 4096 8 f64 TT > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/qe_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:10]: print(f"{float(r[-1])/1e3:9.1f} us  {r[4][:110]}")
PY
