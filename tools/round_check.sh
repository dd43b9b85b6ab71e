#!/bin/bash
# One GPU-box pass: parity tests, full-size C4, tuning variants, bench (both arms), ncu launch list + full capture.
mkdir -p gpurun_out
T=${1:-r01f}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
timeout 600 python tools/big_check.py c4 2>&1 | tee gpurun_out/c4_$T.log
timeout 600 python tools/variant_time.py libbrotlienc_b200.so libvar_minb6.so libvar_minb8.so 2>&1 | tee gpurun_out/variants_$T.log
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_$T.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_ref_$T.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$T.csv python tools/prof_one.py 100000000 > gpurun_out/prof_one.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -c 1 -f -o gpurun_out/prof_walk_$T python tools/prof_one.py 100000000 > gpurun_out/prof_walk.log 2>&1
ls -la gpurun_out | tail -12
