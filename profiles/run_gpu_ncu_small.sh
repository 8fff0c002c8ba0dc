#!/bin/bash
# ncu --set full (with source counters) of one train step of search256: k_tc_fwd_small, k_chain_small, k_tc_bwd_small.  The report comes back.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --epochs 1 --no-e2e --no-cpu-baseline --no-extras --workload search256"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_fwd_small|k_chain_small|k_tc_bwd_small' -s 30 -c 3 -f -o $O/${TAG}_prof_small $B > $O/${TAG}_ncu_small.log 2>&1
python profiles/ncu_extract.py $O/${TAG}_prof_small.ncu-rep > $O/${TAG}_ncu_small.txt 2>&1
ls -la $O | grep ${TAG}; tail -5 $O/${TAG}_ncu_small.log; head -60 $O/${TAG}_ncu_small.txt
