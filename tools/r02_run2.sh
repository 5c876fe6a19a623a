#!/bin/bash
# round-2 GPU call 2: TMA row pass -- correctness (bit-identity with the legacy kernel, at-size oracle parity), timing A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row or fused_pipeline or fused_and_cufft" ) > gpurun_out/r02_tests2.log 2>&1
tail -15 gpurun_out/r02_tests2.log
( time timeout 900 python -m pytest tests/test_gpu_atsize.py tests/test_gpu_qe.py -m gpu -x -q -k "not 8192" ) > gpurun_out/r02_tests2b.log 2>&1
tail -15 gpurun_out/r02_tests2b.log
for kb in tma legacy; do
  ORPHX_KB=$kb timeout 300 python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench2_$kb.json 2> gpurun_out/r02_bench2_$kb.err
  python - <<PY
import json
try:
    e=json.load(open('gpurun_out/r02_bench2_$kb.json')); print('$kb', round(e['value']), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('$kb failed', ex)
PY
done
ORPHX_KB=tma timeout 600 python bench.py --steps 8 --warmup 3 --configs 3 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench2_qe_tma.json 2> gpurun_out/r02_bench2_qe_tma.err
ORPHX_KB=legacy timeout 600 python bench.py --steps 8 --warmup 3 --configs 3 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench2_qe_legacy.json 2> gpurun_out/r02_bench2_qe_legacy.err
python - <<PY
import json
for kb in ('tma','legacy'):
    try:
        e=json.load(open(f'gpurun_out/r02_bench2_qe_{kb}.json'))['configs']['configs[3]']; print(kb, round(e['value'],1), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
    except Exception as ex: print(kb,'failed', ex)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_row_tma --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02b -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02b.log 2>&1
ls -la gpurun_out | tail -8
