#!/bin/bash
# One gpurun call: GPU test groups in separate processes (a faulting kernel poisons only its own group),
# smoke, then a short bench.  Everything is wrapped in `timeout` so a hang can never outlive the call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run t_b1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "act_ or cpp_class or fused_step or resident" -s
run t_fp32 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32 or assembly or controller_step or matmul_add or empty or wide or unsupported"
run t_tc python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensor_core or (ragged and not PREC_FP32) or clamp_mask or null_button or host_buffer or full_size or other_narrow" -s
run smoke python __graft_entry__.py --smoke
TAILN=3 run bench python bench.py --steps 20 --warmup 5
cat gpurun_out/bench.log | tail -1 > gpurun_out/bench.json
