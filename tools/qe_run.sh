python -m pytest tests/test_gpu_qe.py tests/test_gpu_callers.py -x -q 2>&1 | tail -12
for cfg in "4096 4 f64 EB" "8192 1 f64 EB" "8192 2 f32 EB" "4096 8 f64 TT"; do python tools/bench_qe.py $cfg 2>&1 | tail -1 | cut -c1-420; done
ORPHX_QE=cufft python tools/bench_qe.py 4096 4 f64 EB 2>&1 | tail -1 | cut -c1-420
ORPHX_QE=cufft python tools/bench_qe.py 8192 1 f64 EB 2>&1 | tail -1 | cut -c1-420
