#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01j}
timeout 600 python tools/q1_variants.py 2>&1 | tee gpurun_out/q1_variants_$T.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
timeout 600 python tools/c5_check.py 2>&1 | tee gpurun_out/c5_$T.log
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_$T.json
