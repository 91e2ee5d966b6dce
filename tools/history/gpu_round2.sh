#!/bin/bash
# tests + smoke + bench, then a full-set ncu capture of selected kernels at the bench batch size.  $1 tag  $2 regex  $3 skip  $4 count
bash tools/gpu_round.sh
TAG=${1:-x}; KRE=${2:-k_threshold4}; SKIP=${3:-3}; CNT=${4:-1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out | tail -3
