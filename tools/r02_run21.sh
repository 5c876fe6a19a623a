#!/bin/bash
# round-2 GPU call 21: fill counters in front of the mbarrier waits (ABA-safe); 2-row TMA tiles again.  Every step is bounded.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row or fused_pipeline" ) > gpurun_out/r02_tests21a.log 2>&1
tail -4 gpurun_out/r02_tests21a.log
for cfg in "512 4096 8" "4096 4096 2" "4096 4096 8"; do
timeout 200 python tools/dbg_r2.py $cfg 2>&1 | tail -3
done
( time timeout 600 python -m pytest tests/test_gpu_atsize.py -m gpu -x -q -k "2048 or 4096" ) > gpurun_out/r02_tests21.log 2>&1
tail -4 gpurun_out/r02_tests21.log
for r2 in 1 0 1 0; do
ORPHX_KB_R2=$r2 timeout 200 python bench.py --steps 32 --warmup 3 --configs 3 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench21_$r2.json 2> gpurun_out/r02_bench21_$r2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench21_$r2.json')); e=d['configs']['configs[3]']; print('R2=$r2 headline', round(d['value']), {k:round(v['ms_per_launch'],3) for k,v in d['stages'].items()}, 'TT', round(e['value'],1), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('failed', ex)
PY
done
