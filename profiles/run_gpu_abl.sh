#!/bin/bash
# A/B over MFAS_L2_HINTS values on the small-inner_repr step (bits 16 / 32 / 64: timing-only ablations of k_tc_fwd_small, results wrong;
# bit 128: L2 prefetch of the weight rows off).  Usage: run_gpu_abl.sh TAG "v1 v2 ..."
TAG=$1; VALS=${2:-"1 129"}; O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt
rm -f $O/${TAG}_abl.txt
for v in $VALS; do
( env MFAS_L2_HINTS=$v timeout 300 python profiles/small_step_bench.py 2>&1 | grep "search\|mmimdb" | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('hints=$v', r['case'], 'train %.1f' % r['train_step_us'], r['kernels_us'])
" ) >> $O/${TAG}_abl.txt
done
tail -4 $O/${TAG}_pytest.txt; cat $O/${TAG}_abl.txt
