// phik_dmma.cuh -- phi_raw = C_y^T Phi C_x with FP64 tensor-core tiles (sm_100a).
//
// Replaces the hot loop of Basis::spatialCoeff (basis.cpp:122-133) for large
// dense densities (config C3: 8192 x 8192 grid, 32 x 32 basis).  Algorithmic
// work: 8*nx*ny bytes (Phi read once) and 2*nx*ny*nb + 2*ny*nb^2 flops.
//
// Decomposition.  A work unit is 64 rows x (span * 128) columns of Phi.  A
// persistent CTA (8 warps, one per SM) walks its units; inside a unit each warp
// owns 8 rows and sweeps the columns:
//     T[8 rows][32 kx] += Phi[8 rows][4 cols] * C_x[4 cols][32 kx]      (DMMA m8n8k4 x 4)
//  * Phi is streamed straight from HBM into the A fragments: each lane keeps a
//    rotating window of eight 32-byte loads in flight (64 KB per SM), a quad
//    reads one full 128-byte line per row, every sector is used exactly once.
//  * the C_x chunk (128 columns x 32 bases, 36 KB with the conflict-free row
//    pitch of 36 doubles) is staged in shared memory by the TMA bulk-copy engine
//    (cp.async.bulk -> UBLKCP) into a two-stage ring guarded by mbarriers; the
//    chunk rows are stored pre-permuted so the B-fragment loads are
//    bank-conflict free.
//  * at the end of a unit the warp folds its 8 x 32 tile into the running
//    32 x 32 partial with C_y:  P[ky][kx] += C_y[8 rows][ky]^T * T  (32 DMMAs,
//    0.4 % extra work), T transposed through a small shared-memory stage.
// Every CTA writes one 32 x 32 partial; phik_finalize sums them in a fixed
// order (deterministic) and normalises by P[0][0] = sum(Phi).
#pragma once

#include "common.cuh"

namespace eb
{
constexpr int kPdRows = 64;         // rows per unit (8 warps x 8)
constexpr int kPdChunk = 128;       // columns per C_x chunk
constexpr int kPdPitch = 36;        // doubles per chunk row: 36*8 B = 288 = 32 (mod 128) -> conflict-free
constexpr int kPdWarps = 8;
constexpr int kPdChunkBytes = kPdChunk * kPdPitch * 8;  // 36864
constexpr int kPdStagePitch = 33;
constexpr int kPdSmemBytes = 2 * kPdChunkBytes + kPdWarps * 8 * kPdStagePitch * 8 + 64;

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

// 32 contiguous bytes of Phi, streamed (read once: no L1 allocation)
__device__ __forceinline__ void ldg_stream4(const double* p, bool pred, double (&v)[4])
{
  if (pred)
  {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "l"(p + 2));
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
  int nx, ny, nchunks, span, nspans, nunits;
};

__global__ void __launch_bounds__(kPdWarps * 32, 1) phik_dmma_kernel(const PhikDmmaParams p)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* cxs[2] = { reinterpret_cast<double*>(smem_raw), reinterpret_cast<double*>(smem_raw + kPdChunkBytes) };
  double* stage_all = reinterpret_cast<double*>(smem_raw + 2 * kPdChunkBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * kPdChunkBytes + kPdWarps * 8 * kPdStagePitch * 8);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  double* stage = stage_all + warp * 8 * kPdStagePitch;

  if (threadIdx.x == 0)
  {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // iterations of this CTA: its units (grid-strided) x span chunks each
  const int my_units = (p.nunits > (int)blockIdx.x) ? (p.nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_it = my_units * p.span;

  // decode iteration -> (row of this lane, first column of the chunk, chunk id)
  auto decode = [&](int it, int& row, int& col0, int& chunk) {
    const int u = (int)blockIdx.x + (it / p.span) * (int)gridDim.x;
    const int rb = u / p.nspans, cs = u % p.nspans;
    chunk = cs * p.span + it % p.span;
    row = rb * kPdRows + warp * 8 + g;
    col0 = chunk * kPdChunk;
  };

  double P[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) P[a][b][0] = P[a][b][1] = 0.0;

  double v[8][4];  // rotating window of Phi loads: group grp of the current / next chunk
  if (total_it > 0)
  {
    int row, col0, chunk;
    decode(0, row, col0, chunk);
    if (threadIdx.x == 0)
    {
      mbar_expect_tx(&full[0], kPdChunkBytes);
      tma_bulk_g2s(cxs[0], p.cxp + (size_t)chunk * kPdChunk * kPdPitch, kPdChunkBytes, &full[0]);
    }
    const double* src = p.phi + (size_t)row * p.nx + col0 + 4 * q;
#pragma unroll
    for (int grp = 0; grp < 8; grp++)
      ldg_stream4(src + 16 * grp, row < p.ny && col0 + 16 * grp + 4 * q < p.nx, v[grp]);
  }

  double T[4][2];
#pragma unroll
  for (int t = 0; t < 4; t++) T[t][0] = T[t][1] = 0.0;

  for (int it = 0; it < total_it; it++)
  {
    const int cur = it & 1;
    int row, col0, chunk;
    decode(it, row, col0, chunk);
    // next iteration's C_x chunk into the other stage (freed by the barrier at
    // the end of the previous iteration) and this lane's next Phi pointer
    const bool has_next = it + 1 < total_it;
    int nrow = 0, ncol0 = 0, nchunk = 0;
    if (has_next) decode(it + 1, nrow, ncol0, nchunk);
    if (has_next && threadIdx.x == 0)
    {
      mbar_expect_tx(&full[cur ^ 1], kPdChunkBytes);
      tma_bulk_g2s(cxs[cur ^ 1], p.cxp + (size_t)nchunk * kPdChunk * kPdPitch, kPdChunkBytes, &full[cur ^ 1]);
    }
    const double* nsrc = p.phi + (size_t)nrow * p.nx + ncol0 + 4 * q;
    const bool nrow_ok = has_next && nrow < p.ny;

    mbar_wait(&full[cur], (it >> 1) & 1);
    const double* cs = cxs[cur];
#pragma unroll
    for (int grp = 0; grp < 8; grp++)
    {
#pragma unroll
      for (int s = 0; s < 4; s++)
      {
        const double a = v[grp][s];
        const double* brow = cs + (16 * grp + 4 * s + q) * kPdPitch + g;
#pragma unroll
        for (int t = 0; t < 4; t++) dmma884(T[t][0], T[t][1], a, brow[8 * t]);
      }
      // refill this slot with the same group of the next chunk
      ldg_stream4(nsrc + 16 * grp, nrow_ok && ncol0 + 16 * grp + 4 * q < p.nx, v[grp]);
    }

    if ((it % p.span) == p.span - 1)
    {
      // unit finished: P += C_y[rows]^T * T.  T (C layout: row g, cols 8t+2q+e)
      // goes through the warp's stage to become B fragments (row 4h+q, col 8t+g).
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 4; t++)
      {
        stage[g * kPdStagePitch + 8 * t + 2 * q + 0] = T[t][0];
        stage[g * kPdStagePitch + 8 * t + 2 * q + 1] = T[t][1];
        T[t][0] = T[t][1] = 0.0;
      }
      __syncwarp();
      const int rbase = row - g;  // first row of this warp's 8
#pragma unroll
      for (int h = 0; h < 2; h++)
      {
        const int r = rbase + 4 * h + q;
        double a[4], b[4];
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] = r < p.ny ? __ldg(p.cy + (size_t)r * 32 + 8 * m + g) : 0.0;
#pragma unroll
        for (int t = 0; t < 4; t++) b[t] = stage[(4 * h + q) * kPdStagePitch + 8 * t + g];
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int t = 0; t < 4; t++) dmma884(P[m][t][0], P[m][t][1], a[m], b[t]);
      }
    }
    __syncthreads();  // everyone is done with cxs[cur]: it may be refilled next iteration
  }

  // reduce the 8 warps' partials in a fixed order (reusing the C_x ring)
  double* red = reinterpret_cast<double*>(smem_raw);  // [8][1024]
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int t = 0; t < 4; t++)
    {
      red[warp * 1024 + (8 * m + g) * 32 + 8 * t + 2 * q + 0] = P[m][t][0];
      red[warp * 1024 + (8 * m + g) * 32 + 8 * t + 2 * q + 1] = P[m][t][1];
    }
  __syncthreads();
  for (int e = threadIdx.x; e < 1024; e += blockDim.x)
  {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kPdWarps; w++) s += red[w * 1024 + e];
    p.parts[(size_t)blockIdx.x * 1024 + e] = s;
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
  phik_dmma_kernel<<<grid, kPdWarps * 32, kPdSmemBytes, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return grid;
}
}  // namespace eb
