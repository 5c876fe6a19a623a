#!/bin/bash
# round-2 GPU call 3: TMA row pass after the register fixes, separable window A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row or fused_pipeline or fused_and_cufft" ) > gpurun_out/r02_tests3.log 2>&1
tail -6 gpurun_out/r02_tests3.log
( time timeout 900 python -m pytest tests/test_gpu_atsize.py tests/test_gpu_qe.py -m gpu -x -q -k "not 8192" ) > gpurun_out/r02_tests3b.log 2>&1
tail -6 gpurun_out/r02_tests3b.log
for v in "tma 1" "legacy 1" "tma 0" "legacy 0"; do
  set -- $v
  ORPHX_KB=$1 ORPHX_WINDOW_SEPARABLE=$2 timeout 300 python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench3_$1_$2.json 2> gpurun_out/r02_bench3_$1_$2.err
  python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench3_$1_$2.json')); print('$1 sepwin=$2', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('$1 $2 failed', ex)
PY
done
for kb in tma legacy; do
ORPHX_KB=$kb timeout 600 python bench.py --steps 8 --warmup 3 --configs 3 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench3_qe_$kb.json 2> gpurun_out/r02_bench3_qe_$kb.err
done
python - <<PY
import json
for kb in ('tma','legacy'):
    try:
        e=json.load(open(f'gpurun_out/r02_bench3_qe_{kb}.json'))['configs']['configs[3]']; print(kb, round(e['value'],1), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
    except Exception as ex: print(kb,'failed', ex)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_row_tma --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02c -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02c.log 2>&1
ORPHX_KB=legacy timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_row_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02d -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02d.log 2>&1
ls -la gpurun_out | tail -5
