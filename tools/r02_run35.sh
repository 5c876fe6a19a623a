#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_devmap.py tests/test_gpu_refbody.py tests/test_gpu_sim_power.py tests/test_gpu_callers.py -m gpu -x -q ) > gpurun_out/r02_tests35.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r02_tests35.log | head -5
timeout 300 python bench.py --steps 16 --warmup 3 --configs none --cpu-sample 0 > gpurun_out/r02_bench35.json 2> gpurun_out/r02_bench35.err
python - <<PY
import json
try:
    e=json.loads(open('gpurun_out/r02_bench35.json').read().strip().splitlines()[0]); print(round(e['value']), e['e2e_per_call_api'])
except Exception as ex: print('failed', ex)
PY
