#!/bin/bash
# round-2 first pass: sanity tests, mixed-pipe probe, ncu metric names, base vs bank-fix variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
tools/microbench/mix_probe > gpurun_out/mix_probe.txt 2>&1; cat gpurun_out/mix_probe.txt
ncu --query-metrics 2>/dev/null | grep -i -E "dmma|fp64|pipe_fmaheavy|pipe_tensor" | head -60 > gpurun_out/ncu_fp64_metrics.txt; wc -l gpurun_out/ncu_fp64_metrics.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
EB_BENCH="c2 c5 c4" EB_STEPS=20 bash tools/gpu_variants.sh
