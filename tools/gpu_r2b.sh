#!/bin/bash
# round-2 second GPU visit: the N-stacked conv kernel -- correctness (each check in its own process), timings, then the suite
set -u
mkdir -p gpurun_out
XRD_CHECK_TIMEOUT=200 timeout 900 python tools/gpu_diag.py conv3s_fp16 conv3s_cat_fp16 conv3s_stats_fp16 conv3s_cat_stats_fp16 conv3s_gn_fp16 conv3s_gn_cat_fp16 conv3s_bf16 > gpurun_out/r2b_diag.log 2>&1
cp gpurun_out/diag.json gpurun_out/r2b_diag.json
timeout 600 python tools/conv3s_time.py > gpurun_out/r2b_time.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cat gpurun_out/r2b_diag.log | cut -c1-600; cat gpurun_out/r2b_time.log; cat gpurun_out/r2b_pytest.log; cut -c1-300 gpurun_out/r2b_bench.json
