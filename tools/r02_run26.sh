#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/stress_row_pass.py 32 64 2>&1 | tail -2
( time timeout 600 python -m pytest tests/test_gpu_devmap.py -m gpu -x -q ) 2>&1 | grep -E "passed|failed|Error" 
