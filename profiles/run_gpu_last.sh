#!/bin/bash
# Last short call of the round: the cache-builder tests with the 4-rows-per-warp pooling kernel, its roofline, the MM-IMDB timing.
TAG=${1:-l}
O=gpurun_out
mkdir -p $O
( timeout 60 python -m pytest tests/test_cache_builder.py -m gpu -q 2>&1 | tail -40 ) > $O/${TAG}_pytest_pool.txt
( timeout 30 python profiles/pool_bench.py > $O/${TAG}_pool.json 2> $O/${TAG}_pool.err )
( timeout 40 python profiles/mmimdb_bench.py > $O/${TAG}_mmimdb.json 2> $O/${TAG}_mmimdb.err )
tail -15 $O/${TAG}_pytest_pool.txt; cat $O/${TAG}_pool.json; tail -3 $O/${TAG}_pool.err; cat $O/${TAG}_mmimdb.json; tail -5 $O/${TAG}_mmimdb.err
