#!/bin/bash
# round-2 GPU call 9: new option tests, tightened set-up parity, full default bench with the configs block
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sim_power.py tests/test_gpu_qe.py tests/test_gpu_refbody.py -m gpu -x -q -k "options or covsqrt or flat_lensing or refbody or lensing" ) > gpurun_out/r02_tests9.log 2>&1
tail -25 gpurun_out/r02_tests9.log
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err
tail -5 gpurun_out/r02_bench9.err
python - <<PY
import json
try:
    e=json.loads(open('gpurun_out/r02_bench9.json').read().strip().splitlines()[0])
    print(round(e['value']), e['ms_per_step'], e['roofline']['frac'], e['roofline_pipeline']['frac'], e['e2e']['value'], e['e2e_per_call_api']['value'], e['strong_scaling_1024_maps'])
    for k,c in e['configs'].items(): print(k, round(c['value'],1), c['roofline']['frac'], c['e2e']['value'], c.get('setup_seconds'), c.get('sim_to_kappa_on_device'), (c['cpu_baseline'] or {}).get('value'))
    print(e['configs_seconds'])
except Exception as ex: print('failed', ex)
PY
