#!/bin/bash
# First GPU call of the next session (~2 min of box time): measure the opt-in variants and the level-2 host schedule with parity
# check, run the GPU suite with them on, then the evidence for the current defaults.  The earlier calls are under tools/history/.
mkdir -p gpurun_out
timeout 300 python tools/gpu_tune.py > gpurun_out/next_tune.jsonl 2> gpurun_out/next_tune.err
tail -2 gpurun_out/next_tune.err
python tools/tune_report.py gpurun_out/next_tune.jsonl > gpurun_out/next_tune.md
B200AT_TUNE="cluster_eager=4,ccl_sweep=4,ccl_flat=1,qf_mc=3,qf_sort=1,decode_pair=1" B200AT_HOST_PIPE=2 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest_gpu_optin.log 2>&1
tail -3 gpurun_out/next_pytest_gpu_optin.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest_gpu.log 2>&1
tail -3 gpurun_out/next_pytest_gpu.log
B200AT_HOST_TRACE=1 B200AT_HOST_PIPE=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/next_bench_pipe2.json 2> gpurun_out/next_trace_pipe2.txt
cut -c1-300 gpurun_out/next_bench_pipe2.json
