#!/bin/bash
# parity after the first-conv / out_conv / staged depthwise changes, then their timings in situ
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "first_conv or nafnet or sampler or teacher_forced or hybrid_512 or hybrid_final or bf16_operands or full_size_properties_config2" 2>&1 | tail -3 > gpurun_out/r4b_tests.log
cat gpurun_out/r4b_tests.log
(python tools/insitu_profile.py 16 512 2 | head -16
 XRD_DW16_STAGED=0 python tools/profile_step.py 16 512 2 naf
 python tools/profile_step.py 16 512 2 naf
 python tools/profile_step.py 1 512 8 naf
 python tools/profile_step.py 16 512 50 hybrid
 python tools/profile_step.py 1 512 8 hybrid) > gpurun_out/r4b_times.log 2>&1
cut -c1-150 gpurun_out/r4b_times.log
