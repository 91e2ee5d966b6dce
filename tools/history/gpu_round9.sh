#!/bin/bash
# third call: sort-network variant, pipelined sparse host path; GPU suite with everything on
mkdir -p gpurun_out
ALL="ccl_sweep=1,cluster_eager=2,decode_split=1,qf_mc=1,qf_keys23=1,qf_net=1"
timeout 300 python tools/gpu_tune.py > gpurun_out/r04_tune.jsonl 2> gpurun_out/r04_tune.err
tail -2 gpurun_out/r04_tune.err
B200AT_TUNE=$ALL B200AT_SPARSE_H2D=1 B200AT_HOST_PIPE=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r04_pytest_gpu_allon.log 2>&1
tail -3 gpurun_out/r04_pytest_gpu_allon.log
ls -la gpurun_out | tail -4
