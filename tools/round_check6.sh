#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01l}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -c 1 -f -o gpurun_out/prof_walk_$T python tools/prof_one.py 100000000 > gpurun_out/prof_walk.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$T.csv python tools/prof_one.py 100000000 > gpurun_out/prof_one.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_$T.json
