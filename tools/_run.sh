python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['roofline_pipeline']['frac'], d['roofline']['frac'], d['e2e']['value'], d['e2e_per_call_api']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks']); print({k:v['ms_per_launch'] for k,v in d['stages'].items()}); print(d['variants'])"
python bench.py --pol --batch 16 --steps 32 --no-extras --no-e2e --cpu-sample 0 2>/dev/null > gpurun_out/bench_pol_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_pol_final.json')); print('IQU', d['value'], d['roofline_pipeline']['frac'], {k:v['ms_per_launch'] for k,v in d['stages'].items()})"
bash tools/prof_run.sh > /dev/null 2>&1; ls gpurun_out | head -20
