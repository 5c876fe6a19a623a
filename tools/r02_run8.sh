#!/bin/bash
# round-2 GPU call 8: device set-up of the estimator; full GPU suite; set-up timings
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_qe.py -m gpu -x -q -k "setup or filters" ) > gpurun_out/r02_tests8a.log 2>&1
tail -25 gpurun_out/r02_tests8a.log
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests8.log 2>&1
tail -15 gpurun_out/r02_tests8.log
timeout 900 python tools/bench_qe_setup.py 2048 4096 > gpurun_out/r02_qe_setup.log 2>&1
tail -4 gpurun_out/r02_qe_setup.log
