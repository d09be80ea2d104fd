#!/bin/bash
# builds the standalone probes next to their sources (binaries travel to the GPU box with the snapshot)
set -e
cd "$(dirname "$0")"
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ../../include -I ../../medical-image-denoising-using-diffusion_b200/csrc -o "${f%.cu}" "$f" -cudart static
done
