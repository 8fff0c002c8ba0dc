#!/bin/bash
# Generic call: run a command, keep stdout/stderr under gpurun_out/<TAG>.txt.  Usage: gpurun -- 'bash profiles/run_gpu_cmd.sh TAG cmd...'
TAG=$1; shift
mkdir -p gpurun_out
( "$@" ) > gpurun_out/${TAG}.txt 2>&1
tail -60 gpurun_out/${TAG}.txt
