#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config C4 --batch 128 --steps 5 --warmup 3 2>/dev/null > gpurun_out/bench_C4.json; cut -c1-150 gpurun_out/bench_C4.json
timeout 600 python bench.py --config C5 --batch 128 --steps 5 --warmup 3 2>/dev/null > gpurun_out/bench_C5.json; cut -c1-150 gpurun_out/bench_C5.json
timeout 900 python bench.py --config C3 --batch 128 --distinct 8 --steps 3 --warmup 3 2>gpurun_out/bench_C3.err > gpurun_out/bench_C3.json; cut -c1-150 gpurun_out/bench_C3.json; tail -2 gpurun_out/bench_C3.err
timeout 600 python bench.py --config C2 --encoding mono8 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_C2_mono8.json; cut -c1-150 gpurun_out/bench_C2_mono8.json
