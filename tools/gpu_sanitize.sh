#!/bin/bash
# compute-sanitizer over the current default kernels (device path, both quad-fit modes) and the sparse + asynchronous host path.
# $1 = tag.  Output: gpurun_out/<tag>_sanitize_{memcheck,racecheck,synccheck}.log
TAG=${1:-r03}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/${TAG}_sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_sanitize_memcheck.log; tail -4 gpurun_out/${TAG}_sanitize_memcheck.log
# racecheck: the lock-free union-find of the CCL tile kernels (atomicMin label equivalence + path splitting in shared memory) reads
# and writes parent pointers concurrently BY DESIGN (see find_s in k_ccl.cu); first WITH them (the hazards it reports are listed), then
# every other kernel must be clean
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 --kernel-regex kns=k_ccl_tile python tools/sanitize_small.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard|ids" | sort | uniq -c | sort -rn | head -12 > gpurun_out/${TAG}_sanitize_racecheck_ccl_tile.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 30 --kernel-regex-exclude kns=k_ccl_tile --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/${TAG}_sanitize_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_sanitize_racecheck.log; grep -vE "^=========     (and|Saved|Host Frame)" gpurun_out/${TAG}_sanitize_racecheck.log | tail -6
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/${TAG}_sanitize_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/${TAG}_sanitize_synccheck.log; tail -3 gpurun_out/${TAG}_sanitize_synccheck.log
