#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/ptime.py 8192 32 2>&1 | grep -v Warn | tee gpurun_out/ptime.txt
