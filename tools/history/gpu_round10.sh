#!/bin/bash
# fourth call: new library defaults (sweep CCL, 4-px cluster passes, split decode, quad-fit MC + keys23; sparse + pipelined host path)
mkdir -p gpurun_out
OFF="thr_early=0,ccl_sweep=0,cluster_eager=0,decode_split=0,qf_mc=0,qf_keys23=0"
timeout 300 python tools/gpu_tune.py > gpurun_out/r05_tune.jsonl 2> gpurun_out/r05_tune.err
tail -2 gpurun_out/r05_tune.err
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r05_pytest_gpu.log 2>&1
tail -3 gpurun_out/r05_pytest_gpu.log
B200AT_TUNE=$OFF B200AT_SPARSE_H2D=0 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r05_pytest_gpu_oldpath.log 2>&1
tail -2 gpurun_out/r05_pytest_gpu_oldpath.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r05_bench_n1.json 2> gpurun_out/r05_bench_n1.err
cut -c1-400 gpurun_out/r05_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r05_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r05_bench_under_ncu.log 2>&1
timeout 200 python __graft_entry__.py smoke > gpurun_out/r05_smoke.log 2>&1
tail -1 gpurun_out/r05_smoke.log
ls -la gpurun_out | tail -8
