#!/bin/bash
# compute-sanitizer over profiles/sanitize_target.py (every kernel family, tiny shapes).  The bounded mbarrier waits of the
# tensor-core pipelines expire under the tools' slowdown, so the SAME sources are rebuilt with the bounds raised
# (-DMFAS_WAIT_SPINS / -DMFAS_WAIT_CYCLES) into /tmp and loaded through MFAS_LIB_PATH.
# Usage: gpurun --timeout 2400 -- 'bash profiles/run_gpu_sanitize.sh TAG'
TAG=${1:-san}
O=gpurun_out
mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --cudart static \
     -DMFAS_WAIT_SPINS=0xFFFFFFF0u -DMFAS_WAIT_CYCLES=400000000000000LL -o /tmp/_mfas_san.so mfas_b200/csrc/mfas_abi.cu mfas_b200/csrc/host_init.cpp > $O/${TAG}_build.txt 2>&1
export MFAS_LIB_PATH=/tmp/_mfas_san.so
( timeout 700 compute-sanitizer --tool memcheck --print-limit 40 python profiles/sanitize_target.py 2>&1 | tail -80 ) > $O/${TAG}_memcheck.txt
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 200 python profiles/sanitize_target.py 2>&1 | tail -400 ) > $O/${TAG}_racecheck.txt
( timeout 500 compute-sanitizer --tool synccheck --print-limit 40 python profiles/sanitize_target.py 2>&1 | tail -60 ) > $O/${TAG}_synccheck.txt
tail -12 $O/${TAG}_memcheck.txt; grep -c "Race reported\|ERROR\|WARN" $O/${TAG}_racecheck.txt; tail -5 $O/${TAG}_racecheck.txt; tail -5 $O/${TAG}_synccheck.txt
