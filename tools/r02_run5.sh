#!/bin/bash
# round-2 GPU call 5: device-resident maps behind the reference signatures; whole GPU suite; per-call API rate
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_devmap.py -m gpu -x -q ) > gpurun_out/r02_tests5a.log 2>&1
tail -25 gpurun_out/r02_tests5a.log
( time timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_devmap.py ) > gpurun_out/r02_tests5.log 2>&1
tail -25 gpurun_out/r02_tests5.log
timeout 600 python bench.py --steps 32 --warmup 3 --configs none --cpu-sample 0 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench5.json')); print(round(e['value']), e['e2e'], e['e2e_per_call_api'])
except Exception as ex: print('failed', ex)
PY
tail -5 gpurun_out/r02_bench5.err
