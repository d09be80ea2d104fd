#!/bin/bash
# Verification of the last host-side changes (per-device shared-memory limits / SM counts, MicroBatcher close): GPU suite + bench line.
set -u
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02j_gpu_pytest.log
python bench.py --steps 4 --warmup 3 > $O/r02j_bench.json 2> $O/r02j_bench.err
cat $O/r02j_gpu_pytest.log; cut -c1-300 $O/r02j_bench.json; tail -3 $O/r02j_bench.err
