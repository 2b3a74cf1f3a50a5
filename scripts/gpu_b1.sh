#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run t_b1 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "act_ or cpp_class or fused_step or resident or selfdriven" -s
TAILN=3 run b1prof python scripts/b1_profile.py 2000
TAILN=1 run bench python bench.py --steps 5 --warmup 3 --no-cpu
