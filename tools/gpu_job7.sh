#!/bin/bash
mkdir -p gpurun_out
echo "--- c2: v1 (default) vs v2-small"
python tools/ktime.py c2 2>&1 | grep -v Warn
EB_SOLVE_V2_SMALL=1 python tools/ktime.py c2 2>&1 | grep -v Warn
EB_SOLVE_V2_SMALL=1 EB_SOLVE_WIDE=0 python tools/ktime.py c2 2>&1 | grep -v Warn
EB_KTIME_BATCH=262144 python tools/ktime.py c2 2>&1 | grep -v Warn
EB_KTIME_BATCH=262144 EB_SOLVE_V2_SMALL=1 python tools/ktime.py c2 2>&1 | grep -v Warn
EB_SOLVE_V2_SMALL=1 timeout 600 python -m pytest tests/test_gpu_control.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_peer_gather.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload dwa --steps 20 --warmup 3 2>/dev/null | tail -1 | python tools/benchsum.py -
python bench.py --workload collide --steps 20 --warmup 3 2>/dev/null | tail -1 | python tools/benchsum.py -
