#!/bin/bash
# ncu --set full of the wide policy's per-layer launches on one 75,776-row pass (third pass of scripts/wide_one.py)
mkdir -p gpurun_out
CMD="python scripts/wide_one.py 75776"
timeout 300 $CMD > gpurun_out/plain_wide.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wide_gemm|obs_to_blocked" -s 10 -c 5 -f -o gpurun_out/prof_wide_r2 $CMD > gpurun_out/ncu_wide.log 2>&1
echo "wide capture exit $?"; tail -2 gpurun_out/plain_wide.log
