#!/bin/bash
# Quick iteration call: GPU parity tests + device-resident bench (+ optional ncu --set full of kernels matching $2).
# Usage: gpurun --timeout 900 -- 'bash profiles/run_gpu_quick.sh TAG [kernel-regex]'
TAG=${1:-q}
KRE=$2
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest.txt
( timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err )
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' --cache-control none -s 700 -c 900 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --epochs 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
if [ -n "$KRE" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 40 -c 2 -f -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --epochs 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
fi
tail -8 $O/${TAG}_pytest.txt; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
