#!/bin/bash
# A/B call: GPU suite at the defaults, device-resident sweep of $2 (';'-separated B200AT_TUNE strings), ncu launch list of one bench step.
# $1 = tag, $2 = configs, $3 = 1 to also run the suite with qf_exact=1
TAG=${1:-ab}; CFG=${2:-;}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
if [ "$3" = "1" ]; then
  B200AT_TUNE=qf_exact=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu_qf_exact.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu_qf_exact.log
fi
B200AT_TUNE_CONFIGS="$CFG" timeout 600 python tools/gpu_tune.py --device-only > gpurun_out/${TAG}_tune.jsonl 2> gpurun_out/${TAG}_tune.err; tail -2 gpurun_out/${TAG}_tune.err
python tools/tune_report.py gpurun_out/${TAG}_tune.jsonl > gpurun_out/${TAG}_tune.md; cat gpurun_out/${TAG}_tune.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches.csv; cat gpurun_out/${TAG}_launches.csv
