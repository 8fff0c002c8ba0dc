#!/bin/bash
# bench.py at N GPUs (N = $2, default 1): the driver's command line.  Usage: gpurun [--gpus N] -- 'bash profiles/run_gpu_bench.sh TAG N [extra flags]'
TAG=${1:-bench}; N=${2:-1}; shift; shift
O=gpurun_out
mkdir -p $O
if [ "$N" = "1" ]; then
  ( timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 "$@" > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err )
else
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" > $O/${TAG}_n${N}.json 2> $O/${TAG}_n${N}.err )
fi
cat $O/${TAG}_n${N}.json | cut -c1-6000; tail -5 $O/${TAG}_n${N}.err
