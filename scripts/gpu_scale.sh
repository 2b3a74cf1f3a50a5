#!/bin/bash
# The driver's scaling commands at the N given on the command line (run under `gpurun --gpus N`): our arm, then the reference arm.
mkdir -p gpurun_out
for N in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_n$N.log 2> gpurun_out/scale_n$N.err
  echo "N=$N exit $?"; grep -o '{"metric".*' gpurun_out/scale_n$N.log | tail -1 > gpurun_out/scale_n$N.json; head -c 300 gpurun_out/scale_n$N.json; echo
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_ref_n$N.log 2> gpurun_out/scale_ref_n$N.err
  echo "ref N=$N exit $?"; tail -1 gpurun_out/scale_ref_n$N.log | head -c 200; echo
done
