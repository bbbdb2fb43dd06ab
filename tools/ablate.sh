#!/bin/bash
# kernel-only timing of every variants/lib_*.so (tools/ktime.py) on $EB_BENCH (default c2 c5 c4)
mkdir -p gpurun_out
: > gpurun_out/ktime.txt
for lib in variants/lib_*.so; do
  EB_LIB_PATH=$PWD/$lib python tools/ktime.py ${EB_BENCH:-c2 c5 c4} 2>&1 | grep -v Warning | tee -a gpurun_out/ktime.txt
done
