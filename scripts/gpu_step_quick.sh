for n in g0 g1 g0 g1; do
  export GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/exp_$n.so
  echo "=== $n"
  if [ $n = g1 ] && [ -z "$DONE" ]; then DONE=1; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "controller_step" 2>&1 | tail -1; fi
  timeout 300 python scripts/step_time.py 2>&1 | grep -E "step_batch"
done
