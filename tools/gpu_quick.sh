#!/bin/bash
# quick pass: gpu tests (optionally a subset), phase timing, benches
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${EB_TESTS:-} 2>&1 | tail -5
for w in ${EB_PHASES:-c2 c5 c4}; do python tools/phase_timing.py $w 2>&1 | tail -9; done > gpurun_out/phases.txt
cat gpurun_out/phases.txt
for w in ${EB_BENCH:-c2 c5 c4}; do
  python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python tools/benchline.py < gpurun_out/bench_$w.json || tail -5 gpurun_out/bench_$w.err
done
