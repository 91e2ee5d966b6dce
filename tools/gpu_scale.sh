#!/bin/bash
# multi-GPU bench lines exactly as the driver launches them.  $1 = N, then any number of "CONFIG:BATCH:DISTINCT" triples
N=${1:-2}; shift
mkdir -p gpurun_out
bash tools/gpu_topo.sh > /dev/null 2>&1; cp gpurun_out/topo.txt gpurun_out/topo_n$N.txt
PORT=29511
for spec in "$@"; do
  IFS=: read CFG B D <<< "$spec"
  PORT=$((PORT+1))
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 10 --warmup 3 --config $CFG --batch $B --distinct $D --no-cpu-baseline \
    2> gpurun_out/scale_${CFG}_n$N.err > gpurun_out/scale_${CFG}_n$N.json
  tail -2 gpurun_out/scale_${CFG}_n$N.err; cut -c1-300 gpurun_out/scale_${CFG}_n$N.json
done
