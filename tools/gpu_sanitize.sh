#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -15 | tee gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -25 | tee gpurun_out/sanitize_racecheck.log
