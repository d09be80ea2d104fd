#!/bin/bash
# round-2 checkpoint visit: whole GPU suite, smoke, bench (with CPU baseline), launch list of the sampler loop
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2i_smoke.log
timeout 600 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
for w in naf router hybrid ddim; do python tools/profile_step.py 16 512 2 $w; done > gpurun_out/r2i_step_times.log 2>&1
for w in naf hybrid ddim; do python tools/profile_step.py 1 512 8 $w; done >> gpurun_out/r2i_step_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_ddim.csv python tools/profile_step.py 16 512 2 ddim > /dev/null 2>&1
python tools/eval_breakdown.py gpurun_out/r2i_launches_ddim.csv 30 > gpurun_out/r2i_breakdown.log 2>&1
tail -8 gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_smoke.log; cut -c1-400 gpurun_out/r2i_bench.json; cat gpurun_out/r2i_step_times.log; cat gpurun_out/r2i_breakdown.log
