#!/bin/bash
# ncu captures of the hot kernels (one GPU): launch list of one cfg2 epoch, ncu --set full of two cfg2 train steps and of two
# search256 train steps; the per-kernel summaries are extracted on the box (ncu is there) into gpurun_out/.
# Usage: gpurun --timeout 2400 -- 'bash profiles/run_gpu_ncu.sh TAG'
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --epochs 1 --no-e2e --no-cpu-baseline --no-extras"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'^k_' -s 700 -c 900 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_launch.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'k_tc_fwd_ws|k_chain_all|k_tc_bwd_ws' -s 30 -c 6 -f -o $O/${TAG}_prof_cfg2 $B > $O/${TAG}_ncu_full_cfg2.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'k_tc_fwd_small|k_chain_small|k_tc_bwd_small' -s 30 -c 6 -f -o $O/${TAG}_prof_search256 $B --workload search256 > $O/${TAG}_ncu_full_search256.log 2>&1
python profiles/ncu_extract.py $O/${TAG}_prof_cfg2.ncu-rep > $O/${TAG}_ncu_full_cfg2.txt 2>&1
python profiles/ncu_extract.py $O/${TAG}_prof_search256.ncu-rep > $O/${TAG}_ncu_full_search256.txt 2>&1
python profiles/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.txt 2>&1
rm -f $O/${TAG}_prof_search256.ncu-rep      # keep one report (<= 64 MiB comes back)
ls -la $O | grep ${TAG}; head -40 $O/${TAG}_launches.txt; grep -A4 "^---" $O/${TAG}_ncu_full_cfg2.txt | head -40
