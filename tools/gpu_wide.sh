#!/bin/bash
# num_basis > 32: parity tests first, then the whole GPU suite, then a timing of the CTA-per-instance kernel
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_control.py tests/test_gpu_phik.py -x -q -m gpu -k "wide or phik_matches or rejects" 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/wide_tests.log
( time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/gpu_tests.log
timeout 300 python tools/wide_timing.py 2>&1 | tee gpurun_out/wide_timing.log
