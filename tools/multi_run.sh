N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 32 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('bench n=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['check'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_qe.py 4096 8 f64 TT 512 2> gpurun_out/qe_n$N.err | tail -1 | tee gpurun_out/qe_n$N.json
tail -3 gpurun_out/qe_n$N.err
