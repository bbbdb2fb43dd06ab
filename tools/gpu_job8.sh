#!/bin/bash
mkdir -p gpurun_out/prof
EB_PTIME_ALGOS=4,5 python tools/ptime.py 8192 32 2>&1 | grep -v Warn
EB_PTIME_ALGOS=4 python tools/ptime.py 4096 16 2>&1 | grep -v Warn
timeout 300 python -m pytest tests/test_gpu_c3_golden.py tests/test_gpu_phik.py -m gpu -x -q 2>&1 | tail -3
T="tests/test_gpu_control.py::test_every_basis_count_path tests/test_gpu_phik.py tests/test_gpu_control.py::test_long_horizon_many_rounds tests/test_gpu_control.py::test_opt_traj"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 python -m pytest $T -m gpu -x -q > gpurun_out/prof/sanitizer_racecheck2.log 2>&1
echo "racecheck rc=$?"; grep -E "hazard|RACECHECK SUMMARY|passed|failed" gpurun_out/prof/sanitizer_racecheck2.log | cut -c1-220 | sort | uniq -c | sort -rn | head -30
