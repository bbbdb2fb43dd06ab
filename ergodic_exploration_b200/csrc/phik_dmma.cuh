// phik_dmma.cuh -- DMMA/TMA tile kernel for the phi_k contraction (placeholder
// until the tile kernel lands: reports "unsupported" so the shape-agnostic
// kernels in phik_kernels.cuh are used).
#pragma once
#include "common.cuh"
namespace eb
{
inline bool phik_dmma_supported(int, int) { return false; }
inline int phik_dmma_launch(const double*, int, int, const double*, const double*, double*, int, cudaStream_t)
{
  return -1;
}
}  // namespace eb
