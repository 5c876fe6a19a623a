for p in 1 0; do echo "L2_PERSIST=$p"; ORPHX_L2_PERSIST=$p python bench.py --steps 32 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T pipeline', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"
ORPHX_L2_PERSIST=$p python bench.py --pol --batch 16 --steps 16 --no-extras --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('IQU', round(d['value'],1), {k.split()[0]:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
