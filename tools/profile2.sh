#!/bin/bash
# cheaper profile: launch list of one step + full set of selected kernels.  $1 tag, $2 kernel regex, $3 skip, $4 count
TAG=${1:-r01}; KRE=${2:-k_quadfit}; SKIP=${3:-12}; CNT=${4:-4}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG \
  python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out | tail -4
