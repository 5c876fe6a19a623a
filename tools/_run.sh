python -m pytest tests/test_gpu_sim_power.py -x -q 2>&1 | tail -4
python bench.py --dtype f32 --steps 32 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T fp32', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()}, d['check'])"
python bench.py --steps 32 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T fp64', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()}, d['check'])"
