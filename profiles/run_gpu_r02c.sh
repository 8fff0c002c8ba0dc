#!/bin/bash
# Round-2 call C: GPU tests (no -x, failures listed), bench line, depth sweep, MM-IMDB configs[3] timing.
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
rm -f $O/traj_errors.txt
( timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -150 ) > $O/${TAG}_pytest.txt
cp $O/traj_errors.txt $O/${TAG}_traj_errors.txt 2>/dev/null
( timeout 600 python bench.py --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err )
( timeout 600 python profiles/depth_sweep.py > $O/${TAG}_depth_sweep.jsonl 2> $O/${TAG}_depth_sweep.err )
( timeout 600 python profiles/mmimdb_bench.py > $O/${TAG}_mmimdb.json 2> $O/${TAG}_mmimdb.err )
tail -40 $O/${TAG}_pytest.txt; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; cat $O/${TAG}_depth_sweep.jsonl | cut -c1-400; tail -3 $O/${TAG}_depth_sweep.err; cat $O/${TAG}_mmimdb.json; tail -3 $O/${TAG}_mmimdb.err
