#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_phik.py -m gpu -x -q 2>&1 | tail -5
python tools/ptime.py 8192 32 2>&1 | grep -v Warn | tee gpurun_out/ptime.txt
python tools/ptime.py 4096 16 2>&1 | grep -v Warn | tee -a gpurun_out/ptime.txt
python tools/ptime.py 1030 20 770 2>&1 | grep -v Warn | tee -a gpurun_out/ptime.txt
