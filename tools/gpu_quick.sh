#!/bin/bash
# quick A/B call: knob sweep of tools/gpu_tune.py (device path only), optional pytest subset.  $1 = tag, $2 = pytest -k expression ('' = skip)
TAG=${1:-q}; KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log; fi
timeout 400 python tools/gpu_tune.py --device-only > gpurun_out/${TAG}_tune.jsonl 2> gpurun_out/${TAG}_tune.err; tail -2 gpurun_out/${TAG}_tune.err
python tools/tune_report.py gpurun_out/${TAG}_tune.jsonl > gpurun_out/${TAG}_tune.md; cat gpurun_out/${TAG}_tune.md
