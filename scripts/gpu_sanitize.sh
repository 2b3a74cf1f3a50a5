#!/bin/bash
# compute-sanitizer on the mbarrier / TMEM / bulk-copy kernels at small sizes.  ONE tool per gpurun call
# (B200_PROFILING.md): scripts/gpu_sanitize.sh memcheck | racecheck | synccheck | initcheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 240 python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_case.py > gpurun_out/sanitize_$TOOL.log 2>&1
echo "compute-sanitizer --tool $TOOL exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$" gpurun_out/sanitize_$TOOL.log | tail -12
