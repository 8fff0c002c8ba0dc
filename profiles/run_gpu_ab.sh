#!/bin/bash
# A/B call: GPU parity tests, smoke, then the device-resident bench under the env switches given as arguments
# Usage: gpurun --timeout 900 -- 'bash profiles/run_gpu_ab.sh TAG "" "MFAS_FWD_XR=1" "MFAS_HEAD=ffma"'
TAG=$1; shift
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest.txt
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -5 ) > $O/${TAG}_smoke.txt
i=0
for ENVS in "$@"; do
  ( env $ENVS timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/${TAG}_bench_$i.json 2> $O/${TAG}_bench_$i.err )
  echo "== $i [$ENVS]"; python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$i.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["frac"], [(k["kernel"], round(k["ms_per_launch"], 4)) for k in d["roofline"].get("kernels", [])])
except Exception as e:
    print("bench failed:", e); print(open("$O/${TAG}_bench_$i.err").read()[-1500:])
PY
  i=$((i+1))
done
( timeout 200 python tests/cuda/chain_timeline.py 2>&1 | tail -14 ) | tee $O/${TAG}_chain_timeline.txt
tail -6 $O/${TAG}_pytest.txt; cat $O/${TAG}_smoke.txt
