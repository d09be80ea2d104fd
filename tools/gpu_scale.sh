#!/bin/bash
# configs[3] (batch 256 sharded by image, strong scaling) and configs[4] (1024x1024 DDIM-100 tiled, batch 64) on N GPUs of one box
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --global-batch 256 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/scale_strong256_n$N.json 2> gpurun_out/scale_strong256_n$N.err
if [ "$N" = "8" ]; then
  run --workload tiled1024 --global-batch 64 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/scale_tiled1024_n$N.json 2> gpurun_out/scale_tiled1024_n$N.err
fi
cut -c1-700 gpurun_out/scale_*_n$N.json; tail -2 gpurun_out/scale_*_n$N.err
