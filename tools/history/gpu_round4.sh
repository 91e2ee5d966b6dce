#!/bin/bash
mkdir -p gpurun_out
if timeout 60 tools/probe/tma_probe 5 > gpurun_out/tma_probe.log 2>&1; then export B200AT_USE_TMA=1; echo "TMA probe OK -> TMA enabled"; else echo "TMA probe FAILED -> TMA disabled"; fi
cat gpurun_out/tma_probe.log
bash tools/gpu_round.sh
