#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer_gather.py -m gpu -x -q 2>&1 | tail -5
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err; tail -2 gpurun_out/bench_all_n1.err
python tools/benchsum.py gpurun_out/bench_all_n1.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_all_n2.json 2> gpurun_out/bench_all_n2.err
tail -3 gpurun_out/bench_all_n2.err | cut -c1-300
python tools/benchsum.py gpurun_out/bench_all_n2.json
