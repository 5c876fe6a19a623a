#!/bin/bash
# round-2 GPU call 28: four groups of 2-row tiles over six slots (ORPHX_KB=tma4) against the default two groups of 4-row tiles
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row" ) 2>&1 | grep -E "passed|failed|Error|assert" | head
for kb in tma4 tma tma4 tma; do
  ORPHX_KB=$kb timeout 200 python bench.py --steps 64 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench28_$kb.json 2> gpurun_out/r02_bench28_$kb.err
  python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench28_$kb.json')); print('$kb', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()}, e['clocks']['sm_mhz'])
except Exception as ex: print('$kb failed', ex)
PY
done
ORPHX_KB=tma4 timeout 300 ncu --set full --clock-control none -k regex:fused_row_tma --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_tma4 -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02_tma4.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_tma4.ncu-rep 2>/dev/null
