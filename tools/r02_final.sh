#!/bin/bash
# round-2 final validation on one GPU: whole GPU suite, smoke, default bench, reference arm, ncu evidence.  Bounded steps.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02_final_tests.log 2>&1
tail -6 gpurun_out/r02_final_tests.log
( time timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_final_smoke.log 2>&1
tail -4 gpurun_out/r02_final_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
tail -3 gpurun_out/r02_final_bench.err
python - <<PY
import json
try:
    e=json.loads(open('gpurun_out/r02_final_bench.json').read().strip().splitlines()[0])
    print(round(e['value']), e['ms_per_step'], e['roofline']['frac'], e['roofline'].get('traffic_over_algorithmic'), e['roofline_pipeline']['frac'], round(e['e2e']['value']), round(e['e2e_per_call_api']['value']), e['clocks']['sm_mhz'], e['clocks']['reasons'])
    for k,c in e['configs'].items(): print(k, round(c['value'],1), c['roofline']['kernel'], round(c['roofline']['frac'],3), c['roofline'].get('traffic_over_algorithmic'), round(c['e2e']['value'],1), (c['cpu_baseline'] or {}).get('value'))
except Exception as ex: print('failed', ex)
PY
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err
cut -c1-200 gpurun_out/r02_final_ref.json
timeout 900 bash tools/prof_run_r02.sh > gpurun_out/r02_final_prof.log 2>&1
tail -8 gpurun_out/r02_final_prof.log
