#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_atsize.py -m gpu -x -q -k "stress" ) 2>&1 | grep -E "passed|failed|Error|real"
