#!/bin/bash
# Evidence for profiles/ at the current defaults: GPU suite (both quad-fit modes), smoke, bench (N=1) + reference arm, ncu launch list,
# ncu --set full of the roofline kernel at the bench batch (source of roofline.traffic) and of the irregular kernels at batch 32.
# $1 = tag
TAG=${1:-r04}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
B200AT_TUNE=qf_exact=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu_qf_exact.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu_qf_exact.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-200 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref_n1.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_ref_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches.csv; tail -3 gpurun_out/${TAG}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_threshold4|k_preprocess|k_tile_thresh" -s 9 -c 3 -o gpurun_out/${TAG}_prof_dense \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_qf_|k_ccl|k_cluster|k_refine|k_decode|k_reconcile|k_pose" -s 69 -c 23 -o gpurun_out/${TAG}_prof_irregular \
  python bench.py --batch 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | grep ${TAG}_ | tail -12
