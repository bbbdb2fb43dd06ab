#!/bin/bash
# 2-GPU pass: peer-gather tests and the default bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_scale.sh 2
