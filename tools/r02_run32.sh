#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests32.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/r02_tests32.log | head
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1
timeout 300 python bench.py --steps 16 --warmup 3 --configs 3 --cpu-sample 4 > gpurun_out/r02_bench32.json 2> gpurun_out/r02_bench32.err; tail -2 gpurun_out/r02_bench32.err
python - <<PY
import json
try:
    e=json.loads(open('gpurun_out/r02_bench32.json').read().strip().splitlines()[0])
    print(round(e['value']), round(e['e2e_per_call_api']['value']), e['cpu_baseline']['value'], {k:(round(c['value'],1), round(c['e2e']['value'],1), (c['cpu_baseline'] or {}).get('value'), (c.get('sim_to_kappa_on_device') or {}).get('value')) for k,c in e['configs'].items()})
except Exception as ex: print('failed', ex)
PY
