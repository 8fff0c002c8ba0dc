#!/bin/bash
# One short call: the whole GPU suite (no -x: list every failure) + the MM-IMDB configs[3] timing.
# Usage: gpurun --timeout 200 -- 'bash profiles/run_gpu_mmimdb.sh TAG'
TAG=${1:-m}
O=gpurun_out
mkdir -p $O
( timeout 170 python -m pytest tests -m gpu -q 2>&1 | tail -120 ) > $O/${TAG}_pytest.txt
( timeout 60 python profiles/mmimdb_bench.py > $O/${TAG}_mmimdb.json 2> $O/${TAG}_mmimdb.err )
( timeout 40 python profiles/pool_bench.py > $O/${TAG}_pool.json 2> $O/${TAG}_pool.err )
tail -40 $O/${TAG}_pytest.txt; cat $O/${TAG}_pool.json; tail -3 $O/${TAG}_pool.err; cat $O/${TAG}_mmimdb.json; tail -5 $O/${TAG}_mmimdb.err
