#!/bin/bash
# Round-2 call A: every GPU test with the tightened trajectory bounds (errors reported, no -x), the driver-contract bench line,
# the configs[4] depth sweep, compute-sanitizer memcheck / racecheck over profiles/sanitize_target.py.
# Usage: gpurun --timeout 2400 -- 'bash profiles/run_gpu_r02a.sh r02a'
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
rm -f $O/traj_errors.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -q -rf 2>&1 | tail -120 ) > $O/${TAG}_pytest.txt
cp $O/traj_errors.txt $O/${TAG}_traj_errors.txt 2>/dev/null
( timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err )
( timeout 600 python profiles/depth_sweep.py > $O/${TAG}_depth_sweep.jsonl 2> $O/${TAG}_depth_sweep.err )
( timeout 500 compute-sanitizer --tool memcheck --print-limit 30 python profiles/sanitize_target.py 2>&1 | tail -60 ) > $O/${TAG}_memcheck.txt
( timeout 500 compute-sanitizer --tool racecheck --print-limit 30 python profiles/sanitize_target.py 2>&1 | tail -60 ) > $O/${TAG}_racecheck.txt
tail -30 $O/${TAG}_pytest.txt; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; tail -5 $O/${TAG}_memcheck.txt; tail -5 $O/${TAG}_racecheck.txt
