python -m pytest tests/test_gpu_callers.py -x -q -k ilc 2>&1 | tail -15
