#!/bin/bash
# racecheck alone with the hazard limit raised; summary = unique (kernel, access pair by source line).  Usage: run_gpu_racecheck.sh TAG
TAG=${1:-race}
O=gpurun_out
mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --cudart static -w \
     -DMFAS_WAIT_SPINS=0xFFFFFFF0u -DMFAS_WAIT_CYCLES=400000000000000LL -o /tmp/_mfas_san.so mfas_b200/csrc/mfas_abi.cu mfas_b200/csrc/host_init.cpp > $O/${TAG}_build.txt 2>&1
export MFAS_LIB_PATH=/tmp/_mfas_san.so
export NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=100000
( timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 100000 python profiles/sanitize_target.py 2>&1 ) > /tmp/race_full.txt
python - <<'PY' > $O/${TAG}_racecheck_summary.txt
import re, collections
t = open('/tmp/race_full.txt').read().splitlines()
pairs = collections.Counter()
for i, l in enumerate(t):
    m = re.search(r"(Error|Warning): Race reported between (\w+) access at (.*?)\+0x[0-9a-f]+ in (\S+)", l)
    if m and i + 1 < len(t):
        m2 = re.search(r"and (\w+) access at (.*?)\+0x[0-9a-f]+ in (\S+)", t[i + 1])
        if m2:
            def short(s): return re.sub(r"\(.*", "", s).replace("void ", "").replace("mfas::", "")
            pairs[(m.group(1), m.group(2), short(m.group(3)), m.group(4), m2.group(1), short(m2.group(2)), m2.group(3))] += 1
for k, n in sorted(pairs.items(), key=lambda x: -x[1]):
    print(n, *k)
print([l for l in t if 'RACECHECK SUMMARY' in l or l.startswith('ok ')])
PY
cat $O/${TAG}_racecheck_summary.txt | cut -c1-260
