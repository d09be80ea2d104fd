#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload tiled1024 --global-batch $((8*N)) --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/scale_tiled1024_n$N.json 2> gpurun_out/scale_tiled1024_n$N.err
cut -c1-900 gpurun_out/scale_tiled1024_n$N.json; grep -i "libxrd\|timed out" gpurun_out/scale_tiled1024_n$N.err | head -5
