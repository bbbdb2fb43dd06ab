// phik_dmma.cuh -- phi_raw = C_y^T Phi C_x with FP64 tensor-core tiles (sm_100a).
//
// Replaces the hot loop of Basis::spatialCoeff (basis.cpp:122-133) for large
// dense densities (config C3: 8192 x 8192 grid, 32 x 32 basis).  Algorithmic
// work: 8*nx*ny bytes (Phi read once) and 2*nx*ny*nb + 2*ny*nb^2 flops.
//
// Decomposition.  A work unit is 64 rows x (span * 128) columns of Phi.  A
// persistent CTA (16 consumer warps + 1 producer warp, one CTA per SM) walks its
// units; inside a unit consumer warp w owns rows 8*(w%8)..+8 and the (w/8)-th
// 64-column half of every chunk:
//     T[8 rows][32 kx] += Phi[8 rows][4 cols] * C_x[4 cols][32 kx]      (DMMA m8n8k4 x 4)
//  * Phi is streamed straight from HBM into the A fragments: each lane keeps a
//    register double buffer of four 32-byte loads per chunk (64 KB in flight
//    per SM, issued a whole chunk ahead); a quad reads 128 contiguous bytes per
//    row, every sector is used exactly once.
//  * the C_x chunk (128 columns x 32 bases, 36 KB with the conflict-free row
//    pitch of 36 doubles) is staged in shared memory by the TMA bulk-copy engine
//    (cp.async.bulk -> UBLKCP) into a three-stage ring with full/empty
//    mbarriers, issued by a dedicated producer warp, so the consumer warps never
//    meet at a CTA barrier inside a unit; the chunk rows are stored pre-permuted
//    so the B-fragment loads are bank-conflict free.  The producer warp also
//    runs cp.async.bulk.prefetch.L2 over the Phi rows two chunk iterations
//    ahead, so the register window is refilled at L2 latency, not HBM latency.
//  * at the end of a unit the two column halves' tiles are summed in a
//    64 x 32 shared-memory stage and the CTA folds it into its running 32 x 32
//    partial with C_y:  P[ky][kx] += C_y[64 rows][ky]^T * T, one 8 x 8 output
//    tile per warp (16 DMMAs each, ~3 % extra work at span = 8).
// Every CTA writes one 32 x 32 partial; phik_finalize sums them in a fixed
// order (deterministic) and normalises by P[0][0] = sum(Phi).
#pragma once

#include <cstdlib>

#include "common.cuh"

namespace eb
{
constexpr int kPdRows = 64;         // rows per unit (8 row groups of 8)
constexpr int kPdChunk = 128;       // columns per C_x chunk
constexpr int kPdPitch = 36;        // doubles per chunk row: 36*8 B = 288 = 32 (mod 128) -> conflict-free
constexpr int kPdWarps = 16;        // consumer warps: 8 row groups x 2 column halves of every chunk
constexpr int kPdThreads = (kPdWarps + 1) * 32;  // + 1 producer warp (TMA issue, L2 prefetch of Phi)
constexpr int kPdStages = 3;        // C_x ring depth
constexpr int kPdAhead = 2;         // Phi is prefetched into L2 this many chunk iterations ahead
constexpr int kPdChunkBytes = kPdChunk * kPdPitch * 8;  // 36864
constexpr int kPdStagePitch = 33;
constexpr int kPdSmemBytes = kPdStages * kPdChunkBytes + kPdRows * kPdStagePitch * 8 + 128;

inline bool phik_dmma_supported(int nx, int ny) { return nx % 4 == 0 && nx >= kPdChunk && ny >= 1; }

// C_x table re-laid for the tile kernel: rows padded to a multiple of 128,
// pitch 36, and within every group of 16 columns column (4q + s) is stored at
// row (4s + q) so that lanes (q, g) of a B-fragment load hit distinct banks.
__global__ void phik_permute_cx(const double* __restrict__ cx, int nx, int rows_padded, double* __restrict__ out)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_padded * kPdPitch) return;
  const int j = idx / kPdPitch, k = idx % kPdPitch;
  const int G = j >> 4, w = j & 15, q = w >> 2, s = w & 3;
  const int row = 16 * G + 4 * s + q;
  out[(size_t)row * kPdPitch + k] = (j < nx && k < 32) ? cx[(size_t)j * 32 + k] : 0.0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA prefetch of a contiguous global range into L2 (no shared-memory destination)
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void consumer_barrier()
{
  asm volatile("bar.sync 1, %0;" ::"n"(kPdWarps * 32) : "memory");
}

// 32 contiguous bytes of Phi, streamed (read once: no L1 allocation)
__device__ __forceinline__ void ldg_stream4(const double* p, bool pred, double (&v)[4])
{
  if (pred)
  {
    // one 256-bit load (sm_100+): a quad covers a full 128-byte line per row
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
  }
  else
    v[0] = v[1] = v[2] = v[3] = 0.0;
}

struct PhikDmmaParams
{
  const double* phi;   // [ny][nx]
  const double* cxp;   // permuted C_x, [nchunks*128][36]
  const double* cy;    // C_y, [ny][32]
  double* parts;       // [gridDim.x][1024]
  int nx, ny, nchunks, span, nspans, nunits, ahead;
};

// iteration cursor: unit id + chunk index inside the unit, advanced without divisions
struct PdCursor
{
  int unit, k, rb, cs;
  __device__ __forceinline__ void init(const PhikDmmaParams& p, int first_unit)
  {
    unit = first_unit;
    k = 0;
    rb = unit / p.nspans;
    cs = unit - rb * p.nspans;
  }
  __device__ __forceinline__ void advance(const PhikDmmaParams& p, int stride)
  {
    if (++k == p.span)
    {
      k = 0;
      unit += stride;
      rb = unit / p.nspans;
      cs = unit - rb * p.nspans;
    }
  }
  __device__ __forceinline__ int chunk(const PhikDmmaParams& p) const { return cs * p.span + k; }
};

__global__ void __launch_bounds__(kPdThreads, 1) phik_dmma_kernel(const PhikDmmaParams p)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* const cxs0 = reinterpret_cast<double*>(smem_raw);
  double* const tstage = reinterpret_cast<double*>(smem_raw + kPdStages * kPdChunkBytes);  // [64][33]
  uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + kPdStages * kPdChunkBytes + kPdRows * kPdStagePitch * 8);
  uint64_t* const empty = full + kPdStages;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < kPdStages; s++)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kPdWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // this CTA's units are blockIdx.x, blockIdx.x + gridDim.x, ...; each is `span` chunk iterations
  const int my_units = (p.nunits > (int)blockIdx.x) ? (p.nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_it = my_units * p.span;

  if (warp == kPdWarps)
  {
    // ===== producer warp =====
    PdCursor cx_cur, pf_cur;
    cx_cur.init(p, (int)blockIdx.x);
    pf_cur.init(p, (int)blockIdx.x);
    int pf_it = 0;
    auto prefetch_phi = [&]() {  // 64 rows x 1 KB of the chunk at pf_cur, two rows per lane
      const int col0 = pf_cur.chunk(p) * kPdChunk;
      if (col0 < p.nx)
      {
        const uint32_t bytes = (uint32_t)min(kPdChunk, p.nx - col0) * 8u;
#pragma unroll
        for (int r = 0; r < 2; r++)
        {
          const int row = pf_cur.rb * kPdRows + lane + 32 * r;
          if (row < p.ny) tma_prefetch_l2(p.phi + (size_t)row * p.nx + col0, bytes);
        }
      }
      pf_cur.advance(p, (int)gridDim.x);
      pf_it++;
    };
    while (pf_it < min(p.ahead, total_it)) prefetch_phi();
    for (int it = 0; it < total_it; it++)
    {
      const int s = it % kPdStages;
      if (it >= kPdStages) mbar_wait(&empty[s], ((it / kPdStages) - 1) & 1);
      if (lane == 0)
      {
        mbar_expect_tx(&full[s], kPdChunkBytes);
        tma_bulk_g2s(cxs0 + s * (kPdChunkBytes / 8), p.cxp + (size_t)cx_cur.chunk(p) * kPdChunk * kPdPitch,
                     kPdChunkBytes, &full[s]);
      }
      cx_cur.advance(p, (int)gridDim.x);
      if (p.ahead > 0 && pf_it < total_it) prefetch_phi();
    }
    return;
  }

  // ===== consumer warps =====
  const int g = lane >> 2, q = lane & 3;
  const int rg = warp & 7, half = warp >> 3;

  double P[2] = { 0.0, 0.0 };  // this warp's 8x8 tile of the CTA's 32x32 partial
  double T[4][2];
#pragma unroll
  for (int t = 0; t < 4; t++) T[t][0] = T[t][1] = 0.0;

  PdCursor ld_cur;  // the chunk whose Phi is being LOADED (one iteration ahead of the compute)
  ld_cur.init(p, (int)blockIdx.x);
  int cur_rb = ld_cur.rb;

  // Phi fragments are double-buffered in registers: the 32-byte loads of chunk
  // it + 1 are all issued BEFORE the DMMAs of chunk it (ptxas otherwise sinks
  // them behind the last use of a shared window and exposes the full latency).
  double va[4][4], vb[4][4];
  if (total_it > 0)
  {
    const int row = ld_cur.rb * kPdRows + rg * 8 + g, col0 = ld_cur.chunk(p) * kPdChunk + half * 64;
    const double* src = p.phi + (size_t)row * p.nx + col0 + 4 * q;
#pragma unroll
    for (int grp = 0; grp < 4; grp++)
      ldg_stream4(src + 16 * grp, row < p.ny && col0 + 16 * grp + 4 * q < p.nx, va[grp]);
    ld_cur.advance(p, (int)gridDim.x);
  }

  int k_in_unit = 0;
  auto iteration = [&](const int it, double (&vc)[4][4], double (&vn)[4][4]) {
    const int s = it % kPdStages;
    const bool has_next = it + 1 < total_it;
    const int nrow = ld_cur.rb * kPdRows + rg * 8 + g, ncol0 = ld_cur.chunk(p) * kPdChunk + half * 64;
    const int next_rb = ld_cur.rb;
    const double* nsrc = p.phi + (size_t)nrow * p.nx + ncol0 + 4 * q;
    const bool nrow_ok = has_next && nrow < p.ny;
    if (has_next) ld_cur.advance(p, (int)gridDim.x);
#pragma unroll
    for (int grp = 0; grp < 4; grp++)
      ldg_stream4(nsrc + 16 * grp, nrow_ok && ncol0 + 16 * grp + 4 * q < p.nx, vn[grp]);

    mbar_wait(&full[s], (it / kPdStages) & 1);
    const double* bbase = cxs0 + s * (kPdChunkBytes / 8) + (half * 64 + q) * kPdPitch + g;
#pragma unroll
    for (int grp = 0; grp < 4; grp++)
#pragma unroll
      for (int st = 0; st < 4; st++)
      {
        const double a = vc[grp][st];
        const double* brow = bbase + (16 * grp + 4 * st) * kPdPitch;
#pragma unroll
        for (int t = 0; t < 4; t++) dmma884(T[t][0], T[t][1], a, brow[8 * t]);
      }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the C_x stage

    if (++k_in_unit == p.span)
    {
      // Unit finished.  Sum the two column halves' tiles in the stage
      // (C layout: row 8*rg + g, cols 8t + 2q + e), then fold with C_y.
      k_in_unit = 0;
      double* trow = tstage + (rg * 8 + g) * kPdStagePitch + 2 * q;
      if (half == 0)
      {
#pragma unroll
        for (int t = 0; t < 4; t++)
        {
          trow[8 * t + 0] = T[t][0];
          trow[8 * t + 1] = T[t][1];
        }
      }
      consumer_barrier();
      if (half == 1)
      {
#pragma unroll
        for (int t = 0; t < 4; t++)
        {
          trow[8 * t + 0] += T[t][0];
          trow[8 * t + 1] += T[t][1];
        }
      }
#pragma unroll
      for (int t = 0; t < 4; t++) T[t][0] = T[t][1] = 0.0;
      consumer_barrier();
      // warp w owns output tile (m, t) = (w / 4, w % 4):
      // P[8m + g][8t + 2q + e] += sum_rows C_y[row][8m + g] * T[row][8t + ..]
      const int m = warp >> 2, t = warp & 3;
      const int r0 = cur_rb * kPdRows;
#pragma unroll 4
      for (int h = 0; h < kPdRows / 4; h++)
      {
        const int r = r0 + 4 * h + q;
        const double a = r < p.ny ? __ldg(p.cy + (size_t)r * 32 + 8 * m + g) : 0.0;
        const double b = tstage[(4 * h + q) * kPdStagePitch + 8 * t + g];
        dmma884(P[0], P[1], a, b);
      }
      cur_rb = next_rb;
      consumer_barrier();  // the T stage may be overwritten by the next unit
    }
  };
  for (int it = 0; it < total_it; it += 2)
  {
    iteration(it, va, vb);
    if (it + 1 < total_it) iteration(it + 1, vb, va);
  }

  {
    const int m = warp >> 2, t = warp & 3;
    double* out = p.parts + (size_t)blockIdx.x * 1024 + (8 * m + g) * 32 + 8 * t + 2 * q;
    out[0] = P[0];
    out[1] = P[1];
  }
}

// Picks the column span per unit so that the units fill the grid evenly while
// the per-unit epilogue stays small; returns the number of partial blocks
// written (= grid size) or -1 on a launch failure.
inline int phik_dmma_launch(const double* phi, int nx, int ny, const double* cxp, const double* cy, double* parts,
                            int max_parts, cudaStream_t stream)
{
  PhikDmmaParams p{};
  p.phi = phi;
  p.cxp = cxp;
  p.cy = cy;
  p.parts = parts;
  p.nx = nx;
  p.ny = ny;
  p.nchunks = (nx + kPdChunk - 1) / kPdChunk;
  {
    static const int ahead = getenv("EB_PHIK_AHEAD") ? atoi(getenv("EB_PHIK_AHEAD")) : kPdAhead;
    p.ahead = ahead;
  }
  const int row_blocks = (ny + kPdRows - 1) / kPdRows;
  const int grid_max = max_parts;
  double best = -1.0;
  for (int span = 1; span <= p.nchunks; span++)
  {
    const int nspans = (p.nchunks + span - 1) / span;
    const long long units = (long long)row_blocks * nspans;
    const int grid = (int)std::min<long long>(units, grid_max);
    const long long waves = (units + grid - 1) / grid;
    const double balance = (double)units / (double)(waves * grid);      // tail efficiency
    const double padding = (double)p.nchunks / (double)(nspans * span);  // wasted chunk slots
    const double epilogue = 1.0 / (1.0 + 32.0 / (span * (double)kPdChunk) + 0.02 / span);
    const double score = balance * padding * epilogue * ((double)grid / grid_max);
    if (score > best)
    {
      best = score;
      p.span = span;
      p.nspans = nspans;
      p.nunits = (int)units;
    }
  }
  const int grid = std::min(p.nunits, grid_max);
  static bool configured = false;
  if (!configured)
  {
    if (cudaFuncSetAttribute(phik_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPdSmemBytes) != cudaSuccess)
      return -1;
    configured = true;
  }
  phik_dmma_kernel<<<grid, kPdThreads, kPdSmemBytes, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return grid;
}
}  // namespace eb
