#!/bin/bash
# one ncu --set full capture of tc_mlp_kernel (fp16, 1,048,576 rows) after the same command ran clean without ncu
mkdir -p gpurun_out
NAME=${1:-prof_tc_r2}
CMD="python scripts/tc_time.py 1048576"
timeout 300 $CMD > gpurun_out/plain_$NAME.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_mlp -s 6 -c 1 -f -o gpurun_out/$NAME $CMD > gpurun_out/ncu_$NAME.log 2>&1
echo "tc capture exit $?"; tail -3 gpurun_out/plain_$NAME.log
