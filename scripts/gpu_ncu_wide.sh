#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/wide_one.py"
timeout 300 $CMD > gpurun_out/plain_wide.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_gemm -s 8 -c 4 -f -o gpurun_out/prof_wide $CMD > gpurun_out/ncu_wide.log 2>&1
echo "wide capture exit $?"
tail -3 gpurun_out/ncu_wide.log
