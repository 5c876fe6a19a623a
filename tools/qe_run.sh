python tools/bench_callers.py 2048 4 2>&1 | tail -1
