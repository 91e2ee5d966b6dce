#!/bin/bash
mkdir -p gpurun_out
timeout 120 compute-sanitizer --tool memcheck tools/probe/tma_probe 2>&1 | tail -25 | tee gpurun_out/tma_probe.log
bash tools/gpu_round.sh
