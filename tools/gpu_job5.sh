#!/bin/bash
mkdir -p gpurun_out
EB_PTIME_ALGOS=4,5 python tools/ptime.py 8192 32 2>&1 | grep -v Warn
for lib in variants/lib_*.so; do echo $lib; EB_LIB_PATH=$PWD/$lib EB_PTIME_ALGOS=4 python tools/ptime.py 8192 32 2>&1 | grep -v Warn; done
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_c3_golden.py tests/test_gpu_models.py tests/test_gpu_map_target.py tests/test_cpp_adapter.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_control.py -m gpu -x -q -k "wide or out_of_range or too_long or nan_guard" 2>&1 | tail -15
