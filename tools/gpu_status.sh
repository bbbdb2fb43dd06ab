#!/bin/bash
# full status pass on the GPU box: gpu tests, every bench workload, launch list, ncu captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python tools/benchline.py < gpurun_out/bench_default.json
for w in c5 c4 c3 c2big collide dwa; do
  python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python tools/benchline.py < gpurun_out/bench_$w.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch_c2.log 2>&1
for w in c2 c5 c4; do
  ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 4 -c 1 -f -o gpurun_out/solve_$w python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:phik_dmma -s 3 -c 1 -f -o gpurun_out/phik_c3 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/ncu_c3.log 2>&1

ncu --set full --clock-control none --import-source on -k regex:"dwa_control|inflate_scatter|validate_control" -c 3 -f -o gpurun_out/avoid python bench.py --workload dwa --steps 2 --warmup 3 > gpurun_out/ncu_dwa.log 2>&1
ls gpurun_out | head -60
