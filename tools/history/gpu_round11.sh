#!/bin/bash
# fifth (last) call: second DMA queue for the host path
mkdir -p gpurun_out
timeout 60 python tools/gpu_tune.py --host-only > gpurun_out/r06_tune.jsonl 2> gpurun_out/r06_tune.err
tail -2 gpurun_out/r06_tune.err
B200AT_HOST_COPY_STREAMS=2 timeout 50 python -m pytest tests -m gpu -x -q > gpurun_out/r06_pytest_gpu_copy2.log 2>&1
tail -2 gpurun_out/r06_pytest_gpu_copy2.log
B200AT_HOST_COPY_STREAMS=2 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r06_bench_copy2.json 2> gpurun_out/r06_bench_copy2.err
cut -c1-300 gpurun_out/r06_bench_copy2.json
