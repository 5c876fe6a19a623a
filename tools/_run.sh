for lib in liborphx.so liborphx_ka3b2.so liborphx_ka3seq.so; do ORPHX_LIB=$PWD/orphics_b200/_lib/$lib python bench.py --pol --batch 16 --steps 32 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib IQU', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"; done
for lib in liborphx.so liborphx_seg16.so; do echo $lib; ORPHX_LIB=$PWD/orphics_b200/_lib/$lib python tools/bench_qe.py 8192 2 f64 TT 2>&1 | tail -1 | cut -c1-330; done
