#!/bin/bash
TAG=${1:-r02h}
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/${TAG}_pytest.txt
( MFAS_PDL=0 timeout 300 python profiles/small_step_bench.py > $O/${TAG}_small_pdl0.txt 2>&1 )
( MFAS_PDL=1 timeout 300 python profiles/small_step_bench.py > $O/${TAG}_small_pdl1.txt 2>&1 )
( timeout 300 python profiles/e2e_phases.py 2>&1 | grep "===" > $O/${TAG}_phases.txt )
tail -3 $O/${TAG}_pytest.txt; cat $O/${TAG}_small_pdl0.txt | cut -c1-200; echo; cat $O/${TAG}_small_pdl1.txt | cut -c1-200; cat $O/${TAG}_phases.txt
