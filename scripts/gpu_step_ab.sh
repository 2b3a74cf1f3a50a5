#!/bin/bash
# fused fleet-step variants: parity tests of the step paths, then timings (lib/exp_<name>.so)
for n in "$@"; do
  export GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/exp_$n.so
  echo "=== $n"
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu -x -k "controller_step or step_batch or fleet or motor_cmd or saturation" 2>&1 | tail -2
  timeout 300 python scripts/step_time.py 2>&1 | grep -E "infer clamp|step_batch"
done
