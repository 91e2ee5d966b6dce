#!/bin/bash
# One GPU call, ordered by value of the evidence (a cut-off call keeps what was written so far):
#  1. knob sweep with parity check (tools/gpu_tune.py)      2. GPU test suite, default knobs
#  3. GPU test suite with every knob on + sparse staging    4. bench line (default), bench line (all on)
#  5. ncu launch list (all on)                              6. smoke, reference arm
mkdir -p gpurun_out
ALL="thr_early=1,ccl_sweep=1,cluster_eager=2,decode_split=1,qf_mc=1"
date +%s > gpurun_out/r02_t0
timeout 400 python tools/gpu_tune.py > gpurun_out/r02_tune.jsonl 2> gpurun_out/r02_tune.err
tail -2 gpurun_out/r02_tune.err
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_default.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_default.log
B200AT_TUNE=$ALL B200AT_SPARSE_H2D=1 B200AT_HOST_STREAMS=2 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_allon.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_allon.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
cut -c1-300 gpurun_out/r02_bench_default.json
B200AT_TUNE=$ALL timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_allon.json 2> gpurun_out/r02_bench_allon.err
cut -c1-300 gpurun_out/r02_bench_allon.json
B200AT_TUNE=$ALL timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_allon.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
tail -1 gpurun_out/r02_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null
cut -c1-200 gpurun_out/r02_bench_ref.json
date +%s > gpurun_out/r02_t1
ls -la gpurun_out | tail -15
