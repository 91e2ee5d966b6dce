#!/bin/bash
# final validation + evidence for profiles/: tests, smoke, bench (N=1), reference arm, launch list, full-set capture of the roofline kernel
mkdir -p gpurun_out
bash tools/history/gpu_round.sh
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_threshold4|k_preprocess" -s 6 -c 2 -o gpurun_out/prof_final \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_final.log 2>&1
ls -la gpurun_out | tail -5
