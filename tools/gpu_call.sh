#!/bin/bash
# One GPU call of the round: GPU suite in both quad-fit modes, knob sweep, launch list and ncu --set full of the kernels named in $2.
# $1 = tag for the output names, $2 = ncu kernel regex ('' = skip the full-set capture), $3 = launches to skip, $4 = launches to capture
TAG=${1:-r03}; KRE=${2:-}; SKIP=${3:-0}; CNT=${4:-12}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
B200AT_TUNE="qf_exact=1" timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu_exact.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu_exact.log
timeout 400 python tools/gpu_tune.py > gpurun_out/${TAG}_tune.jsonl 2> gpurun_out/${TAG}_tune.err; tail -2 gpurun_out/${TAG}_tune.err
python tools/tune_report.py gpurun_out/${TAG}_tune.jsonl > gpurun_out/${TAG}_tune.md; cat gpurun_out/${TAG}_tune.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
if [ -n "$KRE" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof \
    python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_prof.log 2>&1
fi
ls -la gpurun_out | tail -6
