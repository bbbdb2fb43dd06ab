#!/bin/bash
# compute-sanitizer passes over the hot path (run on a GPU box: gpurun -- 'bash tools/sanitize.sh'):
#   memcheck + racecheck of the fused solve kernel (both kernel generations, wide and 4-warp launches), the phi_k TMA
#   tile kernel and the collision kernels, through the small GPU tests; with 2 GPUs also the peer-gather tests
#   (racecheck sees the shared-memory hazards of the warp-synchronous table exchange; the cross-GPU flag protocol is
#   exercised by bench.py's peer_stress and tests/test_gpu_peer_gather.py, which a sanitizer cannot observe).
mkdir -p gpurun_out
T="tests/test_gpu_control.py::test_batched_warm_state tests/test_gpu_control.py::test_replay_memory_branches tests/test_gpu_control.py::test_every_basis_count_path tests/test_gpu_phik.py tests/test_gpu_collision.py"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --kernel-regex kns=solve_kernel --kernel-regex kns=phik \
    python -m pytest $T -m gpu -x -q > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
