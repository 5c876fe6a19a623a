#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sim_power.py tests/test_gpu_dist.py -m gpu -x -q -k "options or single_rank" ) > gpurun_out/r02_tests16.log 2>&1
tail -15 gpurun_out/r02_tests16.log
