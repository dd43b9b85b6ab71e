#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01g}
timeout 900 python -m pytest tests -m gpu -x -q -k "q1" > gpurun_out/pytest_q1_$T.log 2>&1; echo "pytest q1 exit $?"; tail -15 gpurun_out/pytest_q1_$T.log
timeout 900 python tools/c5_check.py 2>&1 | tee gpurun_out/c5_$T.log
