#!/bin/bash
# round-2 GPU call 17: paired tile order of the TMA row pass
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sim_power.py tests/test_gpu_qe.py -m gpu -x -q -k "tma_row or fused_pipeline or fused_and_cufft_paths_agree_at_2048 or fused_chain" ) > gpurun_out/r02_tests17.log 2>&1
tail -6 gpurun_out/r02_tests17.log
for order in pairs planes pairs planes; do
  ORPHX_KB_TILE_ORDER=$order timeout 300 python bench.py --steps 64 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench17_$order.json 2> gpurun_out/r02_bench17_$order.err
  python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench17_$order.json')); print('$order', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()}, e['clocks']['sm_mhz'])
except Exception as ex: print('$order failed', ex)
PY
done
ORPHX_KB_TILE_ORDER=pairs ORPHX_WINDOW_SEPARABLE=0 timeout 300 python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench17_pairs_genwin.json 2>/dev/null
python - <<PY
import json
e=json.load(open('gpurun_out/r02_bench17_pairs_genwin.json')); print('pairs general window', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
PY
for order in pairs planes; do
ORPHX_KB_TILE_ORDER=$order timeout 600 ncu --set full --clock-control none -k regex:fused_row_tma --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_$order -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02_$order.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_$order.ncu-rep | head -8
done
