python -m pytest tests/test_gpu_qe.py -x -q -k fp32 2>&1 | tail -12
