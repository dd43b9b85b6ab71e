#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01n}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_$T.json
timeout 600 python tools/big_check.py c3 2>&1 | tee gpurun_out/c3_$T.log
timeout 600 python tools/big_check.py c4 2>&1 | tee gpurun_out/c4_$T.log
