#!/bin/bash
# ncu --set full of the fused fleet-step kernel (tc_mlp_kernel<fp16, fused>), 1,048,576 robots (last launches of scripts/step_time.py)
mkdir -p gpurun_out
CMD="python scripts/step_time.py"
timeout 300 $CMD > gpurun_out/plain_step.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_mlp_kernel --launch-skip 44 -c 2 -f -o gpurun_out/prof_step_r2 $CMD > gpurun_out/ncu_step.log 2>&1
echo "step capture exit $?"; tail -6 gpurun_out/plain_step.log; tail -3 gpurun_out/ncu_step.log
