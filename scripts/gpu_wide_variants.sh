#!/bin/bash
for n in "$@"; do
  export GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/exp_$n.so
  echo "=== $n"
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "wide" 2>&1 | tail -1
  timeout 300 python scripts/wide_time.py 2>&1 | grep "fp16\|bf16" | grep "B=  18944\|B= 606208"
done
