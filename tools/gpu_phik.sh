#!/bin/bash
# phi_k tile kernel: tests (optional) and sweeps of the L2 prefetch window / piece, folded (algo 2) and unfolded (algo 3)
[ -n "$EB_TESTS" ] && python -m pytest tests/test_gpu_phik.py -m gpu -q -x 2>&1 | tail -5
for a in ${EB_ALGOS:-2 3}; do for ah in ${EB_AHEADS:--1}; do for pl in ${EB_PFLENS:--1}; do
  echo -n "algo $a ahead $ah pflen $pl: "; EB_PHIK_ALGO=$a EB_PHIK_AHEAD=$ah EB_PHIK_PFLEN=$pl python bench.py --workload c3 --steps 20 --warmup 5 | python tools/benchline.py
done; done; done
