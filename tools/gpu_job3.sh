#!/bin/bash
mkdir -p gpurun_out
echo "--- c2 wide (default) vs EB_SOLVE_WIDE=0"
python tools/ktime.py c2 2>&1 | grep -v Warn
EB_SOLVE_WIDE=0 python tools/ktime.py c2 2>&1 | grep -v Warn
echo "--- ncu phik tma"
ncu --set full --clock-control none --import-source on -k regex:phik_tma_kernel -s 3 -c 1 -f -o gpurun_out/phik_tma_c3 python tools/ptime.py 8192 32 > gpurun_out/ncu_phik_tma.log 2>&1
tail -3 gpurun_out/ncu_phik_tma.log
