// common.cuh -- shared device helpers for the sm_100a ergodic-control kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace eb
{
constexpr double kPi = 3.14159265358979323846;  // numerics.hpp:58
constexpr unsigned kFull = 0xffffffffu;

// One FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).
// Fragment ownership (PTX ISA, mma.m8n8k4 .f64): with g = lane >> 2 and
// q = lane & 3, a = A[g][q], b = B[q][g], {d0, d1} = D[g][2q], D[g][2q + 1].
// On sm_100a this assembles to DMMA.8x8x4.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// inclusive prefix sum over the lanes of a warp (lane 0 first)
__device__ __forceinline__ double warp_scan_incl(double v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const double n = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// inclusive suffix sum over the lanes of a warp (lane 31 first)
__device__ __forceinline__ double warp_scan_incl_rev(double v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const double n = __shfl_down_sync(kFull, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// numerics.hpp:77-89
__device__ __forceinline__ double normalize_angle_pi(double rad)
{
  const double q = floor((rad + kPi) / (2.0 * kPi));
  rad = (rad + kPi) - q * 2.0 * kPi;
  if (rad < 0.0) rad += 2.0 * kPi;
  return rad - kPi;
}

// counter-based generator for the on-device replay sampler (splitmix64 finaliser)
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ double clampd(double v, double lo, double hi)
{
  // std::clamp(v, lo, hi): (v < lo) ? lo : (hi < v) ? hi : v
  return v < lo ? lo : (hi < v ? hi : v);
}
}  // namespace eb
