#!/bin/bash
# One GPU-box pass (run under gpurun).  usage: tools/gpu_pass.sh <tag> <step> [<step> ...]
#   steps: pytest | q234 | fuzz | smoke | c3 | c4 | c5 | bench | benchref | launches | ncu:<kernel regex> | sh:<command>
mkdir -p gpurun_out
T=$1; shift
for step in "$@"; do
  case "$step" in
    pytest)   timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_$T.log ;;
    q234)     timeout 1200 python -m pytest tests/test_zz_gpu_q234.py -m gpu -x -q > gpurun_out/pytest_q234_$T.log 2>&1; echo "q234 exit $?"; tail -4 gpurun_out/pytest_q234_$T.log
              timeout 400 python bench.py --q234-child > gpurun_out/bench_q234_$T.json 2> gpurun_out/bench_q234_err_$T.log; cat gpurun_out/bench_q234_$T.json ;;
    fuzz)     timeout 600 python tools/gpu_fuzz.py 777 1500 2>&1 | tee gpurun_out/gpu_fuzz_$T.log | tail -5 ;;
    smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    c3|c4)    timeout 900 python tools/big_check.py $step 2>&1 | tee gpurun_out/${step}_$T.log ;;
    c5)       timeout 900 python tools/c5_check.py 2>&1 | tee gpurun_out/c5_$T.log ;;
    bench)    timeout 1500 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_err_$T.log; cat gpurun_out/bench_$T.json; tail -3 gpurun_out/bench_err_$T.log ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2>> gpurun_out/bench_err_$T.log; cat gpurun_out/bench_ref_$T.json ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$T.csv python tools/prof_one.py 100000000 > gpurun_out/prof_one_$T.log 2>&1; tail -2 gpurun_out/prof_one_$T.log ;;
    ncu:*)    K=${step#ncu:}; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o gpurun_out/prof_${K}_$T python tools/prof_one.py 100000000 > gpurun_out/prof_ncu_$T.log 2>&1; tail -2 gpurun_out/prof_ncu_$T.log ;;
    sh:*)     bash -c "${step#sh:}" ;;
  esac
done
ls -la gpurun_out | tail -8
