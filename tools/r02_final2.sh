#!/bin/bash
# final state check: whole GPU suite, smoke, default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02_final2_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r02_final2_tests.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1
( time timeout 600 python bench.py ) > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err
python - <<PY
import json
try:
    e=json.loads(open('gpurun_out/r02_final2_bench.json').read().strip().splitlines()[0])
    print(round(e['value']), e['ms_per_step'], e['roofline']['frac'], e['roofline'].get('traffic_over_algorithmic'), e['roofline_pipeline']['frac'], round(e['e2e']['value']), round(e['e2e_per_call_api']['value']), e['clocks']['sm_mhz'], e['clocks']['reasons'])
    for k,c in e['configs'].items(): print(k, round(c['value'],1), c['roofline']['kernel'], round(c['roofline']['frac'],3), c['roofline'].get('traffic_over_algorithmic'), round(c['e2e']['value'],1))
except Exception as ex: print('failed', ex)
PY
