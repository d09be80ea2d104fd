#!/bin/bash
# first conv / out_conv: parity of the sampler after the out_conv change, then one --set full capture of each with the details page
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sampler or teacher_forced or hybrid_512 or f16_nafnet or fp32_nafnet" 2>&1 | tail -3 > gpurun_out/r4a_tests.log
cat gpurun_out/r4a_tests.log
for K in k_first_conv_mma k_conv_cout1_v2; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip 2 --launch-count 1 -o /tmp/rep_$K -f python tools/profile_step.py 16 512 2 ddim > gpurun_out/r4a_ncu_$K.log 2>&1
  ncu -i /tmp/rep_$K.ncu-rep --page details > gpurun_out/r4a_details_$K.txt 2>&1
  ncu -i /tmp/rep_$K.ncu-rep --page source --csv > gpurun_out/r4a_source_$K.csv 2>&1
done
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Executed Ipc Active|Registers Per|Achieved Occupancy|Theoretical Occupancy" gpurun_out/r4a_details_*.txt | head -40
