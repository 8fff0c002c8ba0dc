#!/bin/bash
# GPU parity tests + smoke + the inner_repr=16 search workloads.   Usage: gpurun --timeout 900 -- 'bash profiles/run_gpu_small.sh TAG'
TAG=${1:-s}
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > $O/${TAG}_pytest.txt
tail -30 $O/${TAG}_pytest.txt
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 )
for W in search32 search256; do
  timeout 300 python bench.py --workload $W > $O/${TAG}_bench_$W.json 2> $O/${TAG}_bench_$W.err
done
python - <<PY
import json
for n in ("search32", "search256"):
    try:
        d = json.load(open("$O/${TAG}_bench_%s.json" % n))
        print(n, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), [round(x) for x in d["e2e"]["ms_per_call"]], "frac", round(d["roofline"]["frac"],3), "ms/step", round(d["ms_per_step"],1), "launches", d["gpu_launches"], d["roofline"]["kernel"][:60])
    except Exception as e:
        print(n, "failed", e); print(open("$O/${TAG}_bench_%s.err" % n).read()[-2000:])
PY
