// map_target.cuh -- target density derived from an occupancy grid, sm_100a.
//
// The reference defines the information measure of one cell, entropy(p) (numerics.hpp:164-179), over the cell
// probability GridMap::getCell returns, int8 / 100 (grid.cpp:177-184; -1 = unknown), and names entropy /
// mutual-information targets as the intended use (README.md:20,74) without wiring them up.  Here:
//     Phi[i][j] = entropy(int8[i][j] / 100)            one value per occupancy cell, x fastest
//     phi_k     = (C_y^T Phi C_x) / sum(Phi)           Target::fill's normalisation + Basis::spatialCoeff
// with the cell CENTRES ((j + 1/2) res, (i + 1/2) res) as sample points in the Fourier frame [0, lx] x [0, ly],
// lx = xsize * res -- the mass of a cell sits at its centre, and the centres are mirror-symmetric about lx / 2, so
// the folded phi_k tile kernel applies (phik_dmma.cuh).
// An int8 cell has 256 possible values: entropy is tabulated once per device (entropy_lut_kernel, 256 threads,
// the reference's expression with the device's log) and the density kernel is a pure byte -> double lookup:
// 1 B read + 8 B written per cell, HBM-bound.
#pragma once

#include "common.cuh"

namespace eb
{
// numerics.hpp:164-179 on p = v / 100 (grid.cpp:183); lut[(unsigned char)v]
__global__ void entropy_lut_kernel(double* __restrict__ lut)
{
  const int b = threadIdx.x;  // 0..255 = the byte pattern of the int8 cell
  const int v = b < 128 ? b : b - 256;
  const double p = (double)v / 100.0;
  double e;
  if (fabs(0.0 - p) < 1.0e-12 || fabs(1.0 - p) < 1.0e-12)  // almost_equal, numerics.hpp:67-70
    e = 1e-3;
  else if (p < 0.0)
    e = 0.7;
  else
    e = -p * log(p) - (1.0 - p) * log(1.0 - p);
  lut[b] = e;
}

// Phi = lut[cell].  A warp converts 512 consecutive cells per iteration: lane l takes the cell pairs l, l + 32, ..
// (2-byte loads, 64 contiguous bytes per instruction) and writes them as double2 -- 512 contiguous bytes per store
// instruction, every 128-byte line written by one instruction.  1 B read + 8 B written per cell: HBM-bound.
__global__ void __launch_bounds__(256) entropy_density_kernel(const signed char* __restrict__ cells, long long n,
                                                              const double* __restrict__ lut_g, double* __restrict__ phi)
{
  __shared__ double lut[256];
  lut[threadIdx.x] = lut_g[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool aligned = (reinterpret_cast<uintptr_t>(cells) & 1) == 0 && (reinterpret_cast<uintptr_t>(phi) & 15) == 0;
  const long long nblk = aligned ? (n >> 9) : 0;  // full 512-cell blocks
  const unsigned short* c2 = reinterpret_cast<const unsigned short*>(cells);
  double2* o2 = reinterpret_cast<double2*>(phi);
  for (long long b = warp; b < nblk; b += nwarps)
  {
    const long long p0 = (b << 8) + lane;  // pair index
    unsigned short w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = __ldg(c2 + p0 + 32 * k);
#pragma unroll
    for (int k = 0; k < 8; k++) o2[p0 + 32 * k] = make_double2(lut[w[k] & 255u], lut[w[k] >> 8]);
  }
  // tail (and unaligned buffers): one cell per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (nblk << 9) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    phi[i] = lut[(unsigned char)cells[i]];
}

// random 1-byte loads over an L2-resident buffer: the access shape of the collision probes (roofline denominator)
__global__ void __launch_bounds__(256) l2_gather_probe(const unsigned char* __restrict__ buf, unsigned int mask, int iters,
                                                       unsigned int* __restrict__ sink)
{
  unsigned int s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u, acc = 0;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
      s = s * 1664525u + 1013904223u;
      acc += __ldg(buf + ((s >> 4) & mask));
    }
  }
  if (acc == 0xffffffffu) *sink = acc;
}
}  // namespace eb
