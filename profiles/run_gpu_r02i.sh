#!/bin/bash
TAG=${1:-r02i}
O=gpurun_out
mkdir -p $O
for m in 0 1 2 4 7 0 7; do
( MFAS_PDL=$m timeout 300 python profiles/small_step_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print('PDL=$m', r['case'], 'train %.1f eval %.1f' % (r['train_step_us'], r['eval_step128_us']), r['kernels_us'])
" ) >> $O/${TAG}_pdl.txt
done
cat $O/${TAG}_pdl.txt
