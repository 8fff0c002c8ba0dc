#!/bin/bash
# Lean A/B: the device-resident bench only, once per env-switch string.  Usage: gpurun -- 'bash profiles/run_gpu_ab_lean.sh TAG "" "MFAS_X=1" ...'
TAG=$1; shift
O=gpurun_out
mkdir -p $O
i=0
for ENVS in "$@"; do
  ( env $ENVS timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/${TAG}_bench_$i.json 2> $O/${TAG}_bench_$i.err )
  echo "== $i [$ENVS]"; python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$i.json"))
    print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["roofline"]["frac"], 4), [(k["kernel"], round(k["ms_per_launch"], 4)) for k in d["roofline"].get("kernels", [])], "eval", round(d["roofline"]["eval_step"]["ms"], 4))
except Exception as e:
    print("bench failed:", e); print(open("$O/${TAG}_bench_$i.err").read()[-1500:])
PY
  i=$((i+1))
done
