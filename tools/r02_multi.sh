#!/bin/bash
# round-2 multi-GPU call: tools/r02_multi.sh N  -- the 2-rank NCCL test (N >= 2), the bench at N ranks (weak, with the configs block:
# IQU, strong-scaled TT estimator + mean-field all-reduce, EB 8192^2).  Bounded steps.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q ) > gpurun_out/r02_tests_n2.log 2>&1
  tail -5 gpurun_out/r02_tests_n2.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 64 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
    print('bench n=$N', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['check'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['clocks'].get('per_rank_samples'), d['collective'], d['strong_scaling_1024_maps'])
    for k,c in d['configs'].items(): print(k, round(c['value'],1), c.get('collective'), c.get('meanfield_count_after_allreduce'))
except Exception as ex: print('failed', ex)
PY
tail -3 gpurun_out/r02_bench_n$N.err
