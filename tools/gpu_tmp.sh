mkdir -p gpurun_out
timeout 300 python tools/dbg_quads.py 2>&1 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r04j_pytest_gpu.log 2>&1; tail -3 gpurun_out/r04j_pytest_gpu.log
