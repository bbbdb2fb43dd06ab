#!/bin/bash
# phi_k TMA kernel: time against the number of rows (fixed cost vs streaming rate)
mkdir -p gpurun_out
{
for lib in "" $EB_PHIK3_LIBS; do
  echo "== lib ${lib:-default}"
  for ny in 512 2048 8192 16384; do
    EB_LIB_PATH=$lib EB_PTIME_ALGOS=${EB_PTIME_ALGOS:-4} python tools/ptime.py 8192 32 $ny
  done
done
} > gpurun_out/ptime_rows.txt 2>&1
cat gpurun_out/ptime_rows.txt
