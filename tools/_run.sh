for cfg in "1024 32 f64 EB" "2048 16 f64 EB" "4096 4 f64 EB" "1024 32 f32 EB" "2048 16 f32 EB" "4096 8 f32 EB" "1024 32 f64 TT" "2048 16 f64 TT" "4096 8 f64 TT" "2048 16 f32 TT" "4096 8 f32 TT"; do python tools/bench_qe.py $cfg 2>/dev/null | tail -1; done > gpurun_out/qe_sweep.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/qe_sweep.jsonl'):
    d=json.loads(l); print(d['metric'][:40], round(d['value'],1), round(d['roofline_frac'],3), d['path'])
PY
