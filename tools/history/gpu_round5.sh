#!/bin/bash
bash tools/gpu_round.sh
timeout 600 python bench.py --config C4 --batch 128 --steps 5 --warmup 3 2>gpurun_out/bench_C4.err > gpurun_out/bench_C4.json; cut -c1-150 gpurun_out/bench_C4.json; tail -2 gpurun_out/bench_C4.err
