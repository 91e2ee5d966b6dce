#!/bin/bash
# last seconds of the round's GPU budget: two opt-in variants + a timeline of the host path
mkdir -p gpurun_out
B200AT_HOST_TRACE=1 timeout 40 python tools/gpu_tune.py > gpurun_out/r07_tune.jsonl 2> gpurun_out/r07_trace.txt
tail -3 gpurun_out/r07_tune.jsonl | cut -c1-200
