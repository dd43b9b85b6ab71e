#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01i}
timeout 600 python tools/variant_time.py libvar_a.so libbrotlienc_b200.so libvar_c.so 2>&1 | tee gpurun_out/variants_$T.log
timeout 600 python tools/q1_variants.py 2>&1 | tee gpurun_out/q1_variants_$T.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_q1_parse -c 1 -f -o gpurun_out/prof_q1_parse_$T python tools/q1_variants.py --one > gpurun_out/prof_q1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_q1_$T.csv python tools/q1_variants.py --one > gpurun_out/prof_q1b.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
ls -la gpurun_out | tail -5
