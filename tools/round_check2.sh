#!/bin/bash
mkdir -p gpurun_out
T=${1:-r01h}
timeout 600 python tools/c5_check.py 2>&1 | tee gpurun_out/c5_$T.log
timeout 600 python tools/variant_time.py libbrotlienc_b200.so libvar_a.so libvar_b.so libvar_c.so 2>&1 | tee gpurun_out/variants_$T.log
cp brotli_b200/libvar_b.so brotli_b200/libbrotlienc_b200.so
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_$T.log
