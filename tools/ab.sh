#!/bin/bash
# A/B of library variants on one workload: tools/ab.sh <workload> <steps> <variant|default> ...   (two interleaved passes)
W=$1; S=$2; shift 2
mkdir -p gpurun_out
for pass in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then unset EB_LIB_PATH; else export EB_LIB_PATH=variants/lib_$v.so; fi
    python bench.py --workload $W --steps $S --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v pass $pass: ms/step %.5f  e2e ms %.5f  frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
  done
done
