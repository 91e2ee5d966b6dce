#!/bin/bash
# ncu --set full (with source) of the kernels matching $2 at batch 32, after the GPU suite.  $1 = tag, $2 = kernel regex, $3 = skip, $4 = count
TAG=${1:-prof}; KRE=${2:-k_}; SKIP=${3:-0}; CNT=${4:-12}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
B200AT_TUNE_CONFIGS="${5:-;}" timeout 600 python tools/gpu_tune.py --device-only > gpurun_out/${TAG}_tune.jsonl 2> gpurun_out/${TAG}_tune.err; tail -2 gpurun_out/${TAG}_tune.err
python tools/tune_report.py gpurun_out/${TAG}_tune.jsonl > gpurun_out/${TAG}_tune.md; cat gpurun_out/${TAG}_tune.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches.csv; cat gpurun_out/${TAG}_launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof \
  python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out | grep ${TAG}_ | tail -8
