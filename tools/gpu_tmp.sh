mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r04x_pytest_gpu.log 2>&1; tail -3 gpurun_out/r04x_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r04x_bench.json 2> gpurun_out/r04x_bench.err; tail -2 gpurun_out/r04x_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r04x_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_per_step'], d['latency_720p'], d['gpu_launches'])"
