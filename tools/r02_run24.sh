#!/bin/bash
# round-2 GPU call 24: K_A for IQU with realisations fastest (covsqrt column shared by the resident CTAs)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sim_power.py tests/test_gpu_atsize.py -m gpu -x -q -k "fused_pipeline or fused_and_cufft_paths_agree_at_2048 or pipeline_2048 or tma_row" ) > gpurun_out/r02_tests24.log 2>&1
grep -E "passed|failed" gpurun_out/r02_tests24.log
for i in 1 2; do
timeout 300 python bench.py --steps 16 --warmup 3 --configs 2 --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench24.json 2> gpurun_out/r02_bench24.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench24.json')); e=d['configs']['configs[2]']; print('IQU', round(e['value'],1), {k:round(v['ms_per_launch'],3) for k,v in e['stages'].items()})
except Exception as ex: print('failed', ex)
PY
done
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"fused_sim_col_kernel<double, \(int\)2048, \(int\)3" --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02_IQU2 -f python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 --no-extras --configs 2 > gpurun_out/ncu_r02_IQU2.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_IQU2.ncu-rep 2>/dev/null | head -9
