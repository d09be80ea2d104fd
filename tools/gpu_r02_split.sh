#!/bin/bash
# Split-key attention for single images: the whole GPU suite (attention cases at 1024 / 4096 tokens, hybrid at 512x512 against the
# oracle), then the served shape with the split on and off.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r02k_gpu_pytest.log
cat $O/r02k_gpu_pytest.log
python tools/profile_step.py 1 512 8 ddim > $O/r02k_times.txt 2>&1
XRD_ATT_SPLIT=0 python tools/profile_step.py 1 512 8 ddim >> $O/r02k_times.txt 2>&1
python tools/profile_step.py 1 512 8 hybrid >> $O/r02k_times.txt 2>&1
cat $O/r02k_times.txt
