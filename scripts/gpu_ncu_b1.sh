#!/bin/bash
mkdir -p gpurun_out
B1="python scripts/b1_profile.py 2000"
timeout 300 $B1 > gpurun_out/plain_b1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfdriven -s 1 -c 1 -f -o gpurun_out/prof_b1 $B1 > gpurun_out/ncu_b1.log 2>&1
echo "b1 capture exit $?"; tail -1 gpurun_out/plain_b1.log
