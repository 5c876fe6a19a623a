#!/bin/bash
# round-2 GPU call 1: tests, smoke, bench with the configs block, A/B of the row-pass CTA order, fp32 diagnostic, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
free -g >> gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02_tests1.log 2>&1
tail -30 gpurun_out/r02_tests1.log
( time python __graft_entry__.py smoke ) > gpurun_out/r02_smoke1.log 2>&1; tail -3 gpurun_out/r02_smoke1.log
( time python bench.py --steps 32 --warmup 3 ) > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -5 gpurun_out/r02_bench1.err
python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench1_planes.json 2>> gpurun_out/r02_bench1.err
ORPHX_KB_ORDER=rows python bench.py --steps 32 --warmup 3 --configs none --no-extras --cpu-sample 0 --no-e2e > gpurun_out/r02_bench1_rows.json 2>> gpurun_out/r02_bench1.err
python tools/diag_fp32_qe.py 512 > gpurun_out/r02_diag_fp32.txt 2>&1; cat gpurun_out/r02_diag_fp32.txt
ncu --set full --clock-control none --import-source on -k regex:fused_row --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r02a -f python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --batch 64 --no-extras --configs none > gpurun_out/ncu_r02a.log 2>&1
ls -la gpurun_out | tail -15
