#!/bin/bash
# parity first, then quad-fit variants (stage times), then the full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
run() { env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stages_ms_per_step']; print('$*', 'value', round(d['value']), 'ccl %.2f clu %.2f qf %.2f dec %.2f' % (s['ccl'], s['cluster'], s['quadfit'], s['decode']))"; }
run X=base | tee gpurun_out/tune.log
run B200AT_QF_KEYS23=1 | tee -a gpurun_out/tune.log
run B200AT_QF_SCALE=2 | tee -a gpurun_out/tune.log
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
