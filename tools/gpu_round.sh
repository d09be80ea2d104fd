#!/bin/bash
# One GPU visit: parity tests, NAFNet/router timings, launch list, and a --set full capture exported to CSV on the box
# (the .ncu-rep itself stays in /tmp: gpurun_out/ is limited to 64 MiB).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
for w in naf router hybrid ddim; do python tools/profile_step.py 16 512 2 $w; done > gpurun_out/step_times.log 2>&1
# the served shape (RUN:72-73,107): batch 1, 512x512, inference_steps 8 -> 9 evaluations
for w in naf hybrid ddim; do python tools/profile_step.py 1 512 8 $w; done >> gpurun_out/step_times.log 2>&1
if [ "${NCU_NAF:-1}" = "1" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"k_dwconv|k_layernorm|k_conv_simt|k_simple_gate|k_scale_nc|k_conv1|k_sca" \
      --launch-count ${NCU_COUNT:-30} -o /tmp/naf_rep -f python tools/ncu_targets_naf.py > gpurun_out/ncu_naf.log 2>&1
  python tools/ncu_export.py /tmp/naf_rep.ncu-rep gpurun_out/naf_kernels.csv >> gpurun_out/ncu_naf.log 2>&1
fi
cat gpurun_out/pytest_gpu.log gpurun_out/step_times.log
