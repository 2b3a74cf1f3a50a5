#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full capture of the batched tcgen05 kernel,
# full capture of the bounded batch-1 twin.  Each ncu run directly follows a plain exit-0 run of the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-b1 --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 300 $CMD > gpurun_out/plain_bench2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_mlp -s 3 -c 2 -f -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
echo "tc capture exit $?"
B1="python scripts/b1_profile.py 2000"
timeout 300 $B1 > gpurun_out/plain_b1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfdriven -s 1 -c 1 -f -o gpurun_out/prof_b1 $B1 > gpurun_out/ncu_b1.log 2>&1
echo "b1 capture exit $?"
tail -2 gpurun_out/plain_b1.log
