#!/bin/bash
# second call of the round: re-measure after the sweep-kernel / fetch-kernel changes, isolate host-path costs, source-level profiles
mkdir -p gpurun_out
ALL="cluster_eager=2,decode_split=1,qf_mc=1,qf_keys23=1"
timeout 300 python tools/gpu_tune.py > gpurun_out/r03_tune.jsonl 2> gpurun_out/r03_tune.err
tail -2 gpurun_out/r03_tune.err
B200AT_SPARSE_NOFETCH=1 timeout 200 python tools/gpu_tune.py --host-only --tag nofetch > gpurun_out/r03_tune_nofetch.jsonl 2> gpurun_out/r03_tune_nofetch.err
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r03_pytest_gpu_default.log 2>&1
tail -3 gpurun_out/r03_pytest_gpu_default.log
B200AT_TUNE=$ALL,ccl_sweep=1 B200AT_SPARSE_H2D=1 B200AT_HOST_STREAMS=2 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r03_pytest_gpu_allon.log 2>&1
tail -3 gpurun_out/r03_pytest_gpu_allon.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_quadfit" -s 24 -c 8 -o gpurun_out/r03_prof_quadfit \
  python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r03_prof_quadfit.log 2>&1
B200AT_TUNE=$ALL,ccl_sweep=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_ccl_tile_sweep|k_ccl_flatten|k_ccl_border|k_cluster_pass4|k_refine|k_decode_bits" -s 21 -c 7 -o gpurun_out/r03_prof_allon \
  python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r03_prof_allon.log 2>&1
ls -la gpurun_out | tail -8
