#!/bin/bash
# 2-GPU pass: peer-gather tests, the default bench at N = 2 (with peer stress), loop tuning runs
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_all_n2.json 2> gpurun_out/bench_all_n2.err ) 2>&1 | tail -3
tail -5 gpurun_out/bench_all_n2.err
python tools/benchsum.py gpurun_out/bench_all_n2.json
for thr in 4096 32768; do
  echo "--- loop 16384 total (8192 / GPU), fuse threshold $thr"
  EB_C5_TOTAL=16384 EB_GATHER_FUSE_MIN_BATCH=$thr $TR bench.py --gpus 2 --workload c5loop --loop-steps 300 2>/dev/null | tail -1 | python tools/benchsum.py -
done
echo "--- loop 16384 total on ONE GPU (8192 instances), for the per-tick overhead"
EB_C5_TOTAL=8192 python bench.py --workload c5loop --loop-steps 300 2>/dev/null | tail -1 | python tools/benchsum.py -
