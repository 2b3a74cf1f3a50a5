#!/bin/bash
# ncu capture of the four GEMM launches of one pass of the wide policy; GO2P_LIB selects an experimental build
mkdir -p gpurun_out
CMD="python scripts/wide_one.py"
GO2P_DEBUG=1 timeout 300 $CMD > gpurun_out/plain_wide.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_gemm -s 8 -c 4 -f -o gpurun_out/prof_wide${TAG} $CMD > gpurun_out/ncu_wide.log 2>&1
echo "wide capture exit $?"
tail -6 gpurun_out/plain_wide.log
