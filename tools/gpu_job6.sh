#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err; tail -2 gpurun_out/bench_all_n1.err
python tools/benchsum.py gpurun_out/bench_all_n1.json
python bench.py --workload dwa --steps 20 --warmup 3 2>/dev/null | tail -1 | python tools/benchsum.py -
echo "--- empty kernel (launch cost) wide / 4-warp"
EB_LIB_PATH=$PWD/variants/lib_e128.so python tools/ktime.py c2 2>&1 | grep -v Warn
EB_SOLVE_WIDE=0 EB_LIB_PATH=$PWD/variants/lib_e128.so python tools/ktime.py c2 2>&1 | grep -v Warn
echo "--- edge study"
python tools/edge_study.py 5000000 2>&1 | grep -v Warn | tee gpurun_out/edge_study.txt
echo "--- ncu c2 wide"
bash tools/ncu_solve.sh wide c2
