#!/bin/bash
# Batch-1 (served shape) evidence of the final tree: smoke(), in-situ launch profile, ncu launch list of the sampler.
set -u
O=gpurun_out; mkdir -p $O
python __graft_entry__.py smoke > $O/r02i_smoke.log 2>&1
python tools/insitu_profile.py 1 512 8 > $O/r02i_insitu_b1.txt 2>&1
python tools/profile_step.py 1 512 8 ddim > $O/r02i_b1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02i_launches_ddim_b1.csv python tools/profile_step.py 1 512 8 ddim > $O/r02i_ncu.log 2>&1
tail -3 $O/r02i_smoke.log; cat $O/r02i_b1.log; tail -30 $O/r02i_insitu_b1.txt
