#!/bin/bash
# ncu evidence for one bench configuration.  $1 = tag for output names (e.g. r01a)
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list: every kernel of warmup+1 steps with its device time (cold cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# (2) full sets for the heavy kernels at a smaller batch (ncu replays each kernel ~40x)
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'k_quadfit|k_ccl_tile|k_ccl_border|k_ccl_flatten|k_ccl_mark|k_cluster_pass|k_cluster_select|k_threshold4|k_preprocess|k_decode|k_reconcile|k_pose' \
  -s 51 -c 18 -o gpurun_out/prof_$TAG python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out | tail -8
