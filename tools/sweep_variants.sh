#!/bin/bash
# Runs bench.py over every orphics_b200/_lib/liborphx*.so (tuning variants built by tools/build_variant.sh)
# and prints the per-kernel times: tools/sweep_variants.sh [bench args]
cd "$(dirname "$0")/.."
for lib in orphics_b200/_lib/liborphx*.so; do
  ORPHX_LIB=$PWD/$lib python bench.py --steps 16 --warmup 3 --no-extras --no-e2e --cpu-sample 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s %9.1f maps/s  ' % ('$(basename $lib)', d['value']) + '  '.join('%s %.3f' % (k.split()[0], v['ms_per_launch']) for k,v in d['stages'].items()))"
done
