#!/bin/bash
# A/B of one environment switch over profiles/small_step_bench.py (+ the GPU tests first).  Usage: run_gpu_ab.sh TAG VAR "v1 v2 ..." [pytest -k expr]
TAG=$1; VAR=$2; VALS=$3; KEXPR=$4
O=gpurun_out
mkdir -p $O
if [ -n "$KEXPR" ]; then ( timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt
else ( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt; fi
rm -f $O/${TAG}_ab.txt
for v in $VALS; do
( env $VAR=$v timeout 300 python profiles/small_step_bench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print('$VAR=$v', r['case'], 'train %.1f eval %.1f frac %.3f' % (r['train_step_us'], r['eval_step128_us'], r['frac']), r['kernels_us'])
" ) >> $O/${TAG}_ab.txt
done
tail -4 $O/${TAG}_pytest.txt; cat $O/${TAG}_ab.txt
