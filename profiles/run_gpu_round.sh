#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and an ncu --set full capture of the
# streaming kernels.  Usage (from the repo root): gpurun --timeout 1700 -- 'bash profiles/run_gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest.txt
( timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err )
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err )
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'^k_' -s 700 -c 900 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --epochs 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_tc_fwd_ws|k_chain_all|k_tc_bwd_ws' -s 30 -c 6 -f -o $O/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --epochs 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
tail -5 $O/${TAG}_pytest.txt; cat $O/${TAG}_bench.json; cat $O/${TAG}_bench_reference.json; tail -3 $O/${TAG}_bench.err
