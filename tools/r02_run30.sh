#!/bin/bash
# round-2 GPU call 30: pipelined host path of ox_qe_reconstruct
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_qe.py tests/test_gpu_atsize.py tests/test_gpu_callers.py -m gpu -x -q ) > gpurun_out/r02_tests30.log 2>&1
grep -E "passed|failed|Error|assert" gpurun_out/r02_tests30.log | head
for pl in 1 0; do
ORPHX_QE_PIPELINE=$pl timeout 300 python bench.py --steps 8 --warmup 3 --configs 3,4 --no-extras --cpu-sample 0 > gpurun_out/r02_bench30_$pl.json 2> gpurun_out/r02_bench30_$pl.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench30_$pl.json'))
    for k,c in d['configs'].items(): print('pipeline=$pl', k, round(c['value'],1), 'e2e', round(c['e2e']['value'],1))
except Exception as ex: print('failed', ex)
PY
done
