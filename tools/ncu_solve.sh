#!/bin/bash
# ncu --set full of the solve kernel of the current library: tools/ncu_solve.sh <tag> <workload> [env...]
mkdir -p gpurun_out
tag=$1; w=$2
ncu --set full --clock-control none --import-source on \
  --metrics sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__inst_executed_pipe_fp64.sum \
  -k regex:solve_kernel -s 6 -c 1 -f -o gpurun_out/solve_${w}_$tag python tools/ktime.py $w > gpurun_out/ncu_${w}_$tag.log 2>&1
tail -2 gpurun_out/ncu_${w}_$tag.log
