#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-b1 --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain_bench2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_mlp -s 3 -c 1 -f -o gpurun_out/prof_tc2 $CMD > gpurun_out/ncu_tc2.log 2>&1
echo "tc capture exit $?"
