#!/bin/bash
# A/B of experimental builds of the batched kernel: lib/exp_<name>.so (built by hand with -D flags)
mkdir -p gpurun_out
for n in "$@"; do
  export GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/exp_$n.so
  echo "=== $n"
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "${KSEL:-tensor_core_vs_golden}" 2>&1 | grep "TC prec=2\|passed\|failed" 
  for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-b1 --e2e-steps 1 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4e ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"; done
done
