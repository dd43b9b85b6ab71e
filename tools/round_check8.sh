#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01p}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
timeout 600 python tools/gpu_fuzz.py 777 1500 2>&1 | tee gpurun_out/gpu_fuzz_$T.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_$T.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_ref_$T.json
