#!/bin/bash
TAG=${1:-r02l}
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt
for t in 0 1 2 0 1 2; do
( MFAS_FWD_TMA=$t timeout 300 python profiles/small_step_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print('TMA=$t', r['case'], 'train %.1f eval %.1f' % (r['train_step_us'], r['eval_step128_us']), r['kernels_us'])
" ) >> $O/${TAG}_tma.txt
done
tail -3 $O/${TAG}_pytest.txt; cat $O/${TAG}_tma.txt
