#!/bin/bash
TAG=${1:-r02w}
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt
for v in 0 1 0 1; do
( MFAS_CHAIN_CLUSTER=$v timeout 300 python bench.py --workload mmimdb64 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('MFAS_CHAIN_CLUSTER=$v mmimdb64 value %.0f frac %.3f ms %.1f | e2e %.0f' % (r['value'], r['roofline']['frac'], r['ms_per_step'], r['e2e']['value']))
" ) >> $O/${TAG}_ab.txt 2>&1
done
tail -4 $O/${TAG}_pytest.txt; cat $O/${TAG}_ab.txt
