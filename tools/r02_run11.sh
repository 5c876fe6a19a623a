#!/bin/bash
# round-2 GPU call 11: fused pipeline at 8192^2; smoke; whole GPU suite timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "8192" ) > gpurun_out/r02_tests11a.log 2>&1
tail -25 gpurun_out/r02_tests11a.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke11.log 2>&1
tail -5 gpurun_out/r02_smoke11.log
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests11.log 2>&1
tail -8 gpurun_out/r02_tests11.log
timeout 300 python bench.py --steps 4 --warmup 3 --npix 8192 --batch 4 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench11_8192.json 2> gpurun_out/r02_bench11_8192.err
python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench11_8192.json')); print('8192', e['pipeline_path'], round(e['value'],1), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()}, e['roofline_pipeline']['frac'])
except Exception as ex: print('failed', ex)
PY
tail -3 gpurun_out/r02_bench11_8192.err
