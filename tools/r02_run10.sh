#!/bin/bash
# round-2 GPU call 10: row pass tile order x TMA L2 promotion
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row or fused_pipeline or fused_and_cufft" ) > gpurun_out/r02_tests10.log 2>&1
tail -5 gpurun_out/r02_tests10.log
for order in planes rows; do for promo in 0 128 256; do
  ORPHX_KB_TILE_ORDER=$order ORPHX_KB_L2PROMO=$promo timeout 300 python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench10_${order}_$promo.json 2> gpurun_out/r02_bench10_${order}_$promo.err
  python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench10_${order}_$promo.json')); print('$order promo=$promo', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('$order $promo failed', ex)
PY
done; done
ORPHX_KB_TILE_ORDER=rows timeout 600 python bench.py --steps 8 --warmup 3 --configs 2 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench10_cfg.json 2> gpurun_out/r02_bench10_cfg.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench10_cfg.json'))['configs']
    for k,e in d.items(): print(k, round(e['value'],1), {s:round(v['ms_per_launch'],3) for s,v in e.get('stages',{}).items()})
except Exception as ex: print('cfg failed', ex)
PY
