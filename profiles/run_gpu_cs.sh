#!/bin/bash
# k_chain_small: GPU tests, the chain's phase timeline at the search shape, per-kernel step times.  Usage: run_gpu_cs.sh TAG
TAG=$1; O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt
( env TL_L=2 TL_M=32 TL_H=16 timeout 300 python tests/cuda/chain_timeline.py 2>&1 | tail -12 ) > $O/${TAG}_timeline.txt
( timeout 300 python profiles/small_step_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['case'], 'train %.1f eval %.1f frac %.3f' % (r['train_step_us'], r['eval_step128_us'], r['frac']), r['kernels_us'])
" ) > $O/${TAG}_steps.txt
tail -4 $O/${TAG}_pytest.txt; cat $O/${TAG}_timeline.txt $O/${TAG}_steps.txt
