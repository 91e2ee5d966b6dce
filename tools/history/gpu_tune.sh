#!/bin/bash
run() { env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stages_ms_per_step']; print('$*', 'value', round(d['value']), 'thr %.4f ccl %.2f clu %.2f qf %.2f dec %.2f' % (s['threshold'], s['ccl'], s['cluster'], s['quadfit'], s['decode']), 'roof %.3f' % d['roofline']['frac'])"; }
run X=base
run B200AT_QF_GLOBAL=1
run B200AT_PIPELINE=1 B200AT_QF_SCALE=0.7
run B200AT_PIPELINE=1 B200AT_QF_SCALE=0.5
run B200AT_PIPELINE=1 B200AT_QF_GLOBAL=1 B200AT_QF_SCALE=0.7
