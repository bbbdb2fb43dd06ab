#!/bin/bash
# solve_kernel2 vs the round-1 kernel: parity tests, then kernel timings (EB_SOLVE_V1=1 selects the old kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_control.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -15
echo "--- v1"; EB_SOLVE_V1=1 python tools/ktime.py c5 c4 2>&1 | grep -v Warn
echo "--- v2"; python tools/ktime.py c5 c4 2>&1 | grep -v Warn
for lib in variants/lib_*.so; do [ -e "$lib" ] && EB_LIB_PATH=$PWD/$lib python tools/ktime.py c5 c4 2>&1 | grep -v Warn; done
true
