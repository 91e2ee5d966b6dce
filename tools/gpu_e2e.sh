#!/bin/bash
mkdir -p gpurun_out
for S in 16 24 48; do
  B200AT_HOST_SUB=$S timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('S=$S value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
B200AT_PIPELINE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pipelined value', round(d['value']))"
