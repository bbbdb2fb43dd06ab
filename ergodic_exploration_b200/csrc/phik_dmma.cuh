// phik_dmma.cuh -- phi_raw = C_y^T Phi C_x with FP64 tensor-core tiles (sm_100a).
//
// Replaces the hot loop of Basis::spatialCoeff (basis.cpp:122-133) for large
// dense densities (config C3: 8192 x 8192 grid, 32 x 32 basis).  Algorithmic
// work: 8*nx*ny bytes (Phi read once) and 2*nx*ny*nb + 2*ny*nb^2 flops.
//
// Decomposition.  The grid is cut into 64-row bands and every band into chunks of
// 128 columns; the (band, chunk) iterations are numbered band-major and each
// persistent CTA (16 consumer warps + 1 producer warp, one CTA per SM) takes a
// contiguous range of them.  Consumer warp w owns rows 8*(w%8)..+8 of the band and
// the (w/8)-th 64-column half of every chunk:
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
//    so the B-fragment loads are bank-conflict free.  (The producer warp can
//    also run cp.async.bulk.prefetch.L2 over the band in front of the consumers;
//    with the band-major streaming order it only competes with the demand loads
//    -- measured 0.13 ms without vs 0.18-0.21 ms with, 8192^2 folded -- so the
//    window defaults to 0 and stays as a tuning knob, EB_PHIK_AHEAD / _PFLEN.)
//  * at the end of a band (or of the CTA's range) the two column halves' tiles are summed in a
//    64 x 32 shared-memory stage and the CTA folds it into its running 32 x 32
//    partial with C_y:  P[ky][kx] += C_y[64 rows][ky]^T * T, one 8 x 8 output
//    tile per warp (16 DMMAs each; once or twice per CTA on a large grid).
// Every CTA writes one 32 x 32 partial; phik_finalize sums them in a fixed
// order (deterministic) and normalises by P[0][0] = sum(Phi).
//
// Mirror fold (FOLD = true).  On the configTarget grid x_j = j * res with
// lx = (nx - 1) * res, so cos(k pi x_{nx-1-j} / lx) = (-1)^k cos(k pi x_j / lx):
// the even orders see only Phi[j] + Phi[nx-1-j] and the odd orders only
// Phi[j] - Phi[nx-1-j].  A lane loads 32 bytes from the left half of the row and
// the mirrored 32 bytes from the right half, forms the sum and the difference
// (2 DADD per column pair) and runs HALF the DMMAs: sum x 16 even orders,
// difference x 16 odd orders.  This exploits a symmetry of the BASIS on this
// grid, not any structure of the density.  The plan takes the fold only when the
// measured table symmetry max|C_x[j][k] - (-1)^k C_x[nx-1-j][k]| (accumulated
// grid coordinates are not exactly symmetric) is <= 1e-10, which bounds the
// coefficient error by the same number -- 10x inside the 1e-9 contract; other
// grids (odd nx, lx != (nx-1) res) use FOLD = false.  With the fold the kernel
// is HBM-bound at nb = 32 (4 flop/B against a ridge of ~5.7).
#pragma once

#include <cstdlib>

#include "common.cuh"

namespace eb
{
constexpr int kPdRows = 64;         // rows per unit (8 row groups of 8)
constexpr int kPdChunk = 128;       // columns per C_x chunk (64 column pairs with the mirror fold, PdGeom)
constexpr int kPdPitch = 36;        // doubles per chunk row: 36*8 B = 288 = 32 (mod 128) -> conflict-free
constexpr int kPdWarps = 16;        // consumer warps: 8 row groups x 2 column halves of every chunk
constexpr int kPdThreads = (kPdWarps + 1) * 32;  // + 1 producer warp (TMA issue, L2 prefetch of Phi)
constexpr int kPdStages = 3;        // C_x ring depth
constexpr int kPdAhead = 0;         // L2 prefetch window in front of the consumers, in chunks (0: off)
constexpr int kPdPfLen = 8;         // ... extended in pieces of this many chunks per row
constexpr int kPdStagePitch = 33;

inline bool phik_dmma_supported(int nx, int ny) { return nx % 4 == 0 && nx >= kPdChunk && ny >= 1; }

// C_x table re-laid for the tile kernel: rows padded to a multiple of the chunk,
// pitch 36, and within every group of 16 columns column (4q + s) is stored at
// row (4s + q) so that lanes (q, g) of a B-fragment load hit distinct banks.
// fold: only the left nx/2 columns, and the 32 orders re-ordered as
// [0, 2, .., 30 | 1, 3, .., 31] (even orders multiply the mirror sum, odd the difference).
__global__ void phik_permute_cx(const double* __restrict__ cx, int ncols, int rows_padded, int fold,
                                double* __restrict__ out)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_padded * kPdPitch) return;
  const int j = idx / kPdPitch, k = idx % kPdPitch;
  const int G = j >> 4, w = j & 15, q = w >> 2, s = w & 3;
  const int row = 16 * G + 4 * s + q;
  const int order = fold ? (k < 16 ? 2 * k : 2 * (k - 16) + 1) : k;
  out[(size_t)row * kPdPitch + k] = (j < ncols && k < 32) ? cx[(size_t)j * 32 + order] : 0.0;
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
  const double* cxp;   // permuted C_x, [nchunks*CH][36]
  const double* cy;    // C_y, [ny][32]
  double* parts;       // [gridDim.x][1024]
  int nx, ny;          // grid
  int ncols;           // logical columns walked by the chunks: nx, or nx / 2 with the mirror fold
  int nchunks;         // chunk iterations per row block
  long long total;     // row_blocks * nchunks chunk iterations, row-block-major
  int ahead, pflen;    // L2 prefetch window / piece, in chunks (0: off)
};

// Work distribution.  The (row block, chunk) iterations are numbered row-block-major
// and CTA b takes the contiguous range [b * total / grid, (b + 1) * total / grid): every
// SM streams whole 64-row bands left to right (long sequential runs per DRAM row, one
// C_y fold per band) and the load is balanced to one iteration whatever the grid shape.
// A band shared by two CTAs is simply folded with C_y by both -- the fold is linear.
struct PdCursor
{
  int rb, k;
  __device__ __forceinline__ void init(const PhikDmmaParams& p, long long it)
  {
    rb = (int)(it / p.nchunks);
    k = (int)(it - (long long)rb * p.nchunks);
  }
  __device__ __forceinline__ void advance(const PhikDmmaParams& p)
  {
    if (++k == p.nchunks)
    {
      k = 0;
      rb++;
    }
  }
};

// Geometry of one chunk iteration.  Unfolded: 128 columns, a consumer warp owns 64
// of them (4 groups of 16).  Folded: 64 column PAIRS (64 left columns + their 64
// mirror columns), a consumer warp owns 32 pairs (2 groups) -- the same 16 doubles
// of Phi per lane and chunk, so the register double buffer is unchanged.
template <bool FOLD>
struct PdGeom
{
  static constexpr int kChunk = FOLD ? 64 : 128;
  static constexpr int kGroups = kChunk / 32;             // 16-column groups per warp and chunk
  static constexpr int kChunkBytes = kChunk * kPdPitch * 8;  // C_x stage
  static constexpr int kSmemBytes = kPdStages * kChunkBytes + kPdRows * kPdStagePitch * 8 + 128;
};

template <bool FOLD>
__global__ void __launch_bounds__(kPdThreads, 1) phik_dmma_kernel(const PhikDmmaParams p)
{
  using G = PdGeom<FOLD>;
  constexpr int CH = G::kChunk, GR = G::kGroups, NV = FOLD ? 2 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* const cxs0 = reinterpret_cast<double*>(smem_raw);
  double* const tstage = reinterpret_cast<double*>(smem_raw + kPdStages * G::kChunkBytes);  // [64][33]
  uint64_t* const full =
      reinterpret_cast<uint64_t*>(smem_raw + kPdStages * G::kChunkBytes + kPdRows * kPdStagePitch * 8);
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

  // this CTA's contiguous range of chunk iterations
  const long long it_lo = p.total * (long long)blockIdx.x / (long long)gridDim.x;
  const long long it_hi = p.total * ((long long)blockIdx.x + 1) / (long long)gridDim.x;
  const int total_it = (int)(it_hi - it_lo);

  if (warp == kPdWarps)
  {
    // ===== producer warp =====
    // Besides feeding the C_x ring it keeps a window of the band `ahead` chunks in
    // front of the consumers warm in L2 (cp.async.bulk.prefetch.L2), extended in
    // pieces of `pflen` chunks per row so one instruction covers several KB.
    PdCursor cx_cur;
    cx_cur.init(p, it_lo);
    int pf_rb = -1, pf_k = 0;
    for (int it = 0; it < total_it; it++)
    {
      const int s = it % kPdStages;
      if (it >= kPdStages) mbar_wait(&empty[s], ((it / kPdStages) - 1) & 1);
      if (lane == 0)
      {
        mbar_expect_tx(&full[s], G::kChunkBytes);
        tma_bulk_g2s(cxs0 + s * (G::kChunkBytes / 8), p.cxp + (size_t)cx_cur.k * CH * kPdPitch, G::kChunkBytes,
                     &full[s]);
      }
      if (p.ahead > 0)
      {
        if (cx_cur.rb != pf_rb)
        {
          pf_rb = cx_cur.rb;
          pf_k = cx_cur.k;
        }
        while (pf_k < min(p.nchunks, cx_cur.k + p.ahead))
        {
          const int col0 = pf_k * CH, col1 = min((pf_k + p.pflen) * CH, p.ncols);
          pf_k += p.pflen;
          if (col1 <= col0) continue;
#pragma unroll
          for (int r = 0; r < 2; r++)
          {
            const int row = pf_rb * kPdRows + lane + 32 * r;
            if (row < p.ny)
            {
              const double* base = p.phi + (size_t)row * p.nx;
              tma_prefetch_l2(base + col0, (uint32_t)(col1 - col0) * 8u);
              if (FOLD) tma_prefetch_l2(base + (p.nx - col1), (uint32_t)(col1 - col0) * 8u);
            }
          }
        }
      }
      cx_cur.advance(p);
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
  ld_cur.init(p, it_lo);
  int cur_rb = ld_cur.rb, cur_k = ld_cur.k;  // the chunk being computed

  // Phi fragments are double-buffered in registers: the 32-byte loads of chunk
  // it + 1 are all issued BEFORE the DMMAs of chunk it (ptxas otherwise sinks
  // them behind the last use of a shared window and exposes the full latency).
  // [.][0]: the lane's 4 columns of group grp; [.][1] (fold): their mirror columns.
  double va[GR][NV][4], vb[GR][NV][4];
  auto load_chunk = [&](const int row, const int col0, const bool row_ok, double (&v)[GR][NV][4]) {
    const double* src = p.phi + (size_t)row * p.nx;
#pragma unroll
    for (int grp = 0; grp < GR; grp++)
    {
      const int c = col0 + 16 * grp + 4 * q;  // first of this lane's 4 logical columns
      const bool ok = row_ok && c < p.ncols;
      ldg_stream4(src + c, ok, v[grp][0]);
      if (FOLD) ldg_stream4(src + (p.nx - 4 - c), ok, v[grp][NV - 1]);  // columns nx-1-c-3 .. nx-1-c
    }
  };
  if (total_it > 0)
  {
    const int row = ld_cur.rb * kPdRows + rg * 8 + g;
    load_chunk(row, ld_cur.k * CH + half * (CH / 2), row < p.ny, va);
    ld_cur.advance(p);
  }

  auto iteration = [&](const int it, double (&vc)[GR][NV][4], double (&vn)[GR][NV][4]) {
    const int s = it % kPdStages;
    const bool has_next = it + 1 < total_it;
    const int nrow = ld_cur.rb * kPdRows + rg * 8 + g, ncol0 = ld_cur.k * CH + half * (CH / 2);
    const int next_rb = ld_cur.rb;
    if (has_next) ld_cur.advance(p);
    load_chunk(nrow, ncol0, has_next && nrow < p.ny, vn);

    mbar_wait(&full[s], (it / kPdStages) & 1);
    const double* bbase = cxs0 + s * (G::kChunkBytes / 8) + (half * (CH / 2) + q) * kPdPitch + g;
#pragma unroll
    for (int grp = 0; grp < GR; grp++)
#pragma unroll
      for (int st = 0; st < 4; st++)
      {
        const double* brow = bbase + (16 * grp + 4 * st) * kPdPitch;
        if (FOLD)
        {
          // column c + st pairs with its mirror nx-1-c-st = (nx-4-c) + (3-st)
          const double lo = vc[grp][0][st], hi = vc[grp][NV - 1][3 - st];
          const double ev = lo + hi, od = lo - hi;
          dmma884(T[0][0], T[0][1], ev, brow[0]);   // orders 0, 2, .., 14
          dmma884(T[1][0], T[1][1], ev, brow[8]);   // orders 16, .., 30
          dmma884(T[2][0], T[2][1], od, brow[16]);  // orders 1, 3, .., 15
          dmma884(T[3][0], T[3][1], od, brow[24]);  // orders 17, .., 31
        }
        else
        {
          const double a = vc[grp][0][st];
#pragma unroll
          for (int t = 0; t < 4; t++) dmma884(T[t][0], T[t][1], a, brow[8 * t]);
        }
      }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the C_x stage

    if (++cur_k == p.nchunks || !has_next)
    {
      // End of the band (or of this CTA's range).  Sum the two column halves' tiles in
      // the stage (C layout: row 8*rg + g, cols 8t + 2q + e), then fold with C_y.
      cur_k = 0;
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
    // with the fold the partial's columns are in [even orders | odd orders] order; phik_finalize undoes it
    const int m = warp >> 2, t = warp & 3;
    double* out = p.parts + (size_t)blockIdx.x * 1024 + (8 * m + g) * 32 + 8 * t + 2 * q;
    out[0] = P[0];
    out[1] = P[1];
  }
}

// Launches one persistent CTA per SM (fewer when there is less work than that);
// returns the number of partial blocks written (= grid size) or -1 on a launch failure.
template <bool FOLD>
inline int phik_dmma_launch_t(const double* phi, int nx, int ny, const double* cxp, const double* cy, double* parts,
                              int max_parts, cudaStream_t stream)
{
  using G = PdGeom<FOLD>;
  PhikDmmaParams p{};
  p.phi = phi;
  p.cxp = cxp;
  p.cy = cy;
  p.parts = parts;
  p.nx = nx;
  p.ny = ny;
  p.ncols = FOLD ? nx / 2 : nx;
  p.nchunks = (p.ncols + G::kChunk - 1) / G::kChunk;
  {
    // L2 prefetch window / piece of the producer warp, in chunk iterations (tuned on B200, 8192^2, nb = 32)
    static const int ahead = getenv("EB_PHIK_AHEAD") ? atoi(getenv("EB_PHIK_AHEAD")) : -1;
    static const int pflen = getenv("EB_PHIK_PFLEN") ? atoi(getenv("EB_PHIK_PFLEN")) : -1;
    p.ahead = ahead >= 0 ? ahead : kPdAhead;
    p.pflen = pflen > 0 ? pflen : kPdPfLen;
  }
  const int row_blocks = (ny + kPdRows - 1) / kPdRows;
  p.total = (long long)row_blocks * p.nchunks;
  const int grid = (int)std::min<long long>(p.total, max_parts);
  static bool configured[64] = {};  // per instantiation and device (the attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63])
  {
    if (cudaFuncSetAttribute(phik_dmma_kernel<FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes) !=
        cudaSuccess)
      return -1;
    configured[dev & 63] = true;
  }
  phik_dmma_kernel<FOLD><<<grid, kPdThreads, G::kSmemBytes, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return grid;
}

inline int phik_dmma_launch(const double* phi, int nx, int ny, const double* cxp, const double* cy, double* parts,
                            int max_parts, bool fold, cudaStream_t stream)
{
  return fold ? phik_dmma_launch_t<true>(phi, nx, ny, cxp, cy, parts, max_parts, stream) :
                phik_dmma_launch_t<false>(phi, nx, ny, cxp, cy, parts, max_parts, stream);
}

// the fold needs the mirrored 32-byte loads aligned (nx % 8 == 0 keeps nx / 2 a multiple of 4)
inline bool phik_fold_shape_ok(int nx) { return nx % 8 == 0; }
}  // namespace eb
