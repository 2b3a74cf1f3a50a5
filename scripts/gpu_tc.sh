#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_tc python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensor_core or ragged or clamp_mask or null_button or host_buffer or full_size or coexists" -s
TAILN=2 run bench python bench.py --steps 20 --warmup 5 --no-cpu --no-b1
