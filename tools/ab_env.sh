#!/bin/bash
# A/B of an environment switch on one workload: tools/ab_env.sh <workload> <steps> VAR v1 v2 ...   (two interleaved passes)
W=$1; S=$2; VAR=$3; shift 3
for pass in 1 2; do
  for v in "$@"; do
    env $VAR=$v python bench.py --workload $W --steps $S --warmup 5 --loop-steps 300 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$VAR=$v pass $pass: ms/step %.5f  e2e ms %.5f  frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))"
  done
done
