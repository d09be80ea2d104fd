#!/bin/bash
# depthwise v3 + k_sca grid: NAFNet parity, then NAFNet kernel times with and without v3
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "nafnet or hybrid_512 or hybrid_final or bf16_operands" 2>&1 | tail -3 > gpurun_out/r4d_tests.log
cat gpurun_out/r4d_tests.log
for v in 1 0; do
  XRD_DW16_V3=$v ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4d_launches_naf_v3_$v.csv python tools/profile_step.py 16 512 2 naf 2>&1 | tail -1
  python tools/ncu_summary.py gpurun_out/r4d_launches_naf_v3_$v.csv 600 | head -14
done
XRD_DW16_V3=1 python tools/profile_step.py 16 512 2 naf
XRD_DW16_V3=0 python tools/profile_step.py 16 512 2 naf
