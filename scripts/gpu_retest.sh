#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run t_all python -m pytest tests -q -m gpu -s
