#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run t_wide python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k wide
run wide_time python scripts/wide_time.py
