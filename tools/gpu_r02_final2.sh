#!/bin/bash
# Refresh of the final evidence after the small-batch changes (conv3w output split, one-Q-tile attention, GroupNorm spread):
# GPU suite, bench line, step times.  The batch-16 kernels are the ones of gpu_r02_final.sh.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02h_gpu_pytest.log
python bench.py --steps 4 --warmup 3 > $O/r02h_bench.json 2> $O/r02h_bench.err
for w in naf router hybrid ddim; do python tools/profile_step.py 16 512 2 $w; done > $O/r02h_step_times.txt 2>&1
for w in naf router hybrid ddim; do python tools/profile_step.py 1 512 8 $w; done >> $O/r02h_step_times.txt 2>&1
for b in 2 4; do python tools/profile_step.py $b 512 8 hybrid; done >> $O/r02h_step_times.txt 2>&1
python tools/profile_step.py 16 512 50 hybrid >> $O/r02h_step_times.txt 2>&1
cat $O/r02h_gpu_pytest.log $O/r02h_step_times.txt; cut -c1-400 $O/r02h_bench.json
