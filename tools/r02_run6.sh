#!/bin/bash
# round-2 GPU call 6: warp-per-row radix-32 row pass
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row or fused_pipeline or fused_and_cufft" ) > gpurun_out/r02_tests6.log 2>&1
tail -12 gpurun_out/r02_tests6.log
for v in "w32 1" "tma 1" "w32 0"; do
  set -- $v
  ORPHX_KB=$1 ORPHX_WINDOW_SEPARABLE=$2 timeout 300 python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench6_$1_$2.json 2> gpurun_out/r02_bench6_$1_$2.err
  python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench6_$1_$2.json')); print('$1 sepwin=$2', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('$1 $2 failed', ex)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_row_w32 --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02f -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02f.log 2>&1
ls -la gpurun_out | tail -3
