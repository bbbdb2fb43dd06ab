#!/bin/bash
# compute-sanitizer memcheck + racecheck over the num_basis > 32 paths (solve_kernel_big, the wide phi_k pair)
mkdir -p gpurun_out
T="tests/test_gpu_control.py::test_wide_basis_counts[33] tests/test_gpu_control.py::test_wide_basis_counts[64] tests/test_gpu_control.py::test_wide_basis_persistent_ctas_and_sampled_memory tests/test_gpu_phik.py::test_phik_wide_raw_block_and_tile_algos_refused"
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 99 --kernel-regex kns=solve_kernel_big --kernel-regex kns=phik_stage --kernel-regex kns=phik_finalize_wide \
    python -m pytest $T -m gpu -x -q > gpurun_out/sanitizer_wide_$tool.log 2>&1
  echo "$tool rc=$?"; tail -5 gpurun_out/sanitizer_wide_$tool.log
done
