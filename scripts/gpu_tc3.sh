#!/bin/bash
# parity tests of the batched tensor-core path, then timing of the library builds named on the command line
# (lib/libgo2policy_<name>.so; the default build is always timed)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "${KSEL:-tensor_core or ragged or clamp_mask or null_button or host_buffer or full_size or coexists or narrow or spelling or controller_step}" > gpurun_out/t_tc3.log 2>&1
echo "tests exit $?"; tail -n ${TAILN:-15} gpurun_out/t_tc3.log
timeout 300 python scripts/tc_time.py 2>&1 | tail -3
for n in "$@"; do
  GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/libgo2policy_$n.so timeout 300 python scripts/tc_time.py 2>&1 | tail -3
done
