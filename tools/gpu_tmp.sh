mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r04u_pytest_gpu.log 2>&1; tail -3 gpurun_out/r04u_pytest_gpu.log
bash tools/gpu_latency.sh r04u | tail -32
B200AT_TUNE_CONFIGS=";" timeout 600 python tools/gpu_tune.py --device-only > gpurun_out/r04u_tune.jsonl 2>/dev/null; python tools/tune_report.py gpurun_out/r04u_tune.jsonl | head -5
