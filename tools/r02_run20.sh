#!/bin/bash
# round-2 GPU call 20: ABA-safe mbarrier waits; 2-row TMA tiles again
mkdir -p gpurun_out
for cfg in "512 4096 8" "2048 4096 2" "4096 4096 2" "4096 4096 8"; do
timeout 300 python tools/dbg_r2.py $cfg 2>&1 | tail -3
done
( time timeout 900 python -m pytest tests/test_gpu_sim_power.py tests/test_gpu_atsize.py -m gpu -x -q -k "tma_row or fused_pipeline or 2048 or 4096" ) > gpurun_out/r02_tests20.log 2>&1
tail -5 gpurun_out/r02_tests20.log
for r2 in 1 0 1 0; do
ORPHX_KB_R2=$r2 timeout 600 python bench.py --steps 32 --warmup 3 --configs 3 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench20_$r2.json 2> gpurun_out/r02_bench20_$r2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench20_$r2.json')); e=d['configs']['configs[3]']; print('R2=$r2 headline', round(d['value']), {k:round(v['ms_per_launch'],3) for k,v in d['stages'].items()}, 'TT', round(e['value'],1), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('failed', ex)
PY
done
