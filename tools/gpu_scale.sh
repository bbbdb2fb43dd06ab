#!/bin/bash
# the driver's scaling run, builder-side: bench.py at N GPUs of one box (N = $1), reference arm first at N = 1
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( time $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_all_n$N.json 2> gpurun_out/bench_all_n$N.err ) 2>&1 | tail -3
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_all_n$N.err | tail -5 | cut -c1-300
python tools/benchsum.py gpurun_out/bench_all_n$N.json
