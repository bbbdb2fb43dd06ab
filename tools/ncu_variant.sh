#!/bin/bash
# ncu --set full of the solve kernel for the listed variants on one workload: tools/ncu_variant.sh c5 a0 a95
mkdir -p gpurun_out
w=$1; shift
for v in "$@"; do
  EB_LIB_PATH=$PWD/variants/lib_$v.so ncu --set full --clock-control none --import-source on \
    --metrics sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__inst_executed_pipe_fp64.sum \
    -k regex:solve_kernel -s 6 -c 1 -f -o gpurun_out/solve_${w}_$v python tools/ktime.py $w > gpurun_out/ncu_${w}_$v.log 2>&1
  tail -2 gpurun_out/ncu_${w}_$v.log
done
