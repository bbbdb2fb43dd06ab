#!/bin/bash
# parity diagnostics + bench of every variants/lib_*.so on the workloads in $EB_BENCH (default c2 c5 c4)
mkdir -p gpurun_out
for lib in variants/lib_*.so; do
  v=$(basename $lib .so)
  [ -n "$EB_DIAG" ] && EB_LIB_PATH=$PWD/$lib python tools/diag_parity.py 2>&1 | grep "^model" | sed "s/^/$v: /"
  for w in ${EB_BENCH:-c2 c5 c4}; do
    EB_LIB_PATH=$PWD/$lib python bench.py --workload $w --steps ${EB_STEPS:-20} --warmup 3 > gpurun_out/bench_${v}_$w.json 2> gpurun_out/bench_${v}_$w.err
    echo -n "$v: "; python tools/benchline.py < gpurun_out/bench_${v}_$w.json || tail -3 gpurun_out/bench_${v}_$w.err
  done
done
