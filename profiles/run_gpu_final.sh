#!/bin/bash
# End-of-round check on one GPU: smoke(), the GPU tests, the reference arm, the default bench line.
TAG=${1:-r02final}
O=gpurun_out
mkdir -p $O
rm -f $O/traj_errors.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1 )
( timeout 1200 python -m pytest tests -m gpu -q -rf 2>&1 | tail -40 ) > $O/${TAG}_pytest.txt
cp $O/traj_errors.txt $O/${TAG}_traj_errors.txt 2>/dev/null
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err )
( timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err )
cat $O/${TAG}_smoke.txt; tail -3 $O/${TAG}_pytest.txt; cat $O/${TAG}_bench_reference.json | cut -c1-600; cat $O/${TAG}_bench.json | cut -c1-1500; tail -3 $O/${TAG}_bench.err
