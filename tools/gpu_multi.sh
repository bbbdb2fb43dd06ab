#!/bin/bash
# N-GPU pass (N = $1, default 2): the driver's launch line for every sharded workload
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
for w in ${EB_BENCH:-c2 c5 c3}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err
  python tools/benchline.py < gpurun_out/bench_${w}_n$N.json || tail -5 gpurun_out/bench_${w}_n$N.err
done
[ -z "$EB_SKIP_REF" ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --impl reference --steps 2 --warmup 1 | cut -c1-300
[ -n "$EB_TESTS" ] && python -m pytest $EB_TESTS -m gpu -q 2>&1 | tail -4
