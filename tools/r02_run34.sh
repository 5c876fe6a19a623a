#!/bin/bash
# memcheck of the round's new kernels on small cases (devmap arithmetic, lensing, estimator set-up, TMA row pass)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_devmap.py tests/test_gpu_refbody.py "tests/test_gpu_qe.py::test_device_setup_matches_host_setup" "tests/test_gpu_qe.py::test_flat_lensing_sims_match_oracle" -m gpu -x -q > gpurun_out/r02_memcheck.log 2>&1
echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/r02_memcheck.log | head
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sim_power.py -m gpu -x -q -k "tma_row" > gpurun_out/r02_memcheck_tma.log 2>&1
echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/r02_memcheck_tma.log | head
