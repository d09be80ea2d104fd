#!/bin/bash
# conv3s rework (warpgroup roles + setmaxnreg, decade schedule, chunked register epilogue): correctness, timings, ncu
set -u
mkdir -p gpurun_out
T=${TAG:-r2j}
XRD_CHECK_TIMEOUT=200 timeout 600 python tools/gpu_diag.py conv3s_fp16 conv3s_cat_fp16 conv3s_stats_fp16 conv3s_cat_stats_fp16 conv3s_gn_fp16 conv3s_gn_cat_fp16 conv3s_bf16 > gpurun_out/${T}_diag.log 2>&1
timeout 600 python tools/conv3s_time.py > gpurun_out/${T}_time.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "blob or yardstick or benched_configuration or stacked" 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
ncu --set full --clock-control none --import-source on -k regex:"k_conv3s" --launch-count 2 -o /tmp/c3s_rep -f python tools/ncu_targets_c3s.py > gpurun_out/${T}_ncu.log 2>&1
ncu -i /tmp/c3s_rep.ncu-rep --page source --csv > gpurun_out/${T}_src.csv 2>> gpurun_out/${T}_ncu.log
python tools/ncu_export.py /tmp/c3s_rep.ncu-rep gpurun_out/${T}_c3s_raw.csv >> gpurun_out/${T}_ncu.log 2>&1
cut -c1-500 gpurun_out/${T}_diag.log; cut -c1-420 gpurun_out/${T}_time.log; cat gpurun_out/${T}_pytest.log
