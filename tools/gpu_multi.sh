#!/bin/bash
# multi-GPU bench exactly as the driver launches it.  $1 = N
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>>gpurun_out/bench_n$N.err | tee gpurun_out/bench_ref_n$N.json
