// FP64 latency / issue-interval probes for B200 (sm_100a).  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu
// One warp on one SM: dependent-chain latency (cycles per op) of DFMA, DADD,
// DMMA (same accumulator), 64-bit SHFL, LDS.64; then the per-SMSP issue
// interval of DFMA / DMMA with ILP independent chains in 1..8 warps.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int kIters = 4096;

__global__ void lat_dfma(double* out, long long* cyc, double a, double b)
{
  double v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; i++) v = fma(v, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}
__global__ void lat_dadd(double* out, long long* cyc, double a)
{
  double v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; i++) v = v + a;
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}
__global__ void lat_dmma(double* out, long long* cyc, double a, double b)
{
  double d0 = threadIdx.x, d1 = 1.0;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; i++) dmma(d0, d1, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = d0 + d1;
}
__global__ void lat_shfl(double* out, long long* cyc)
{
  double v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; i++) v = __shfl_up_sync(0xffffffffu, v, 1);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}
__global__ void lat_shfl_add(double* out, long long* cyc)
{
  double v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; i++) v += __shfl_up_sync(0xffffffffu, v, 1);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}
__global__ void lat_lds(double* out, long long* cyc)
{
  __shared__ long long sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = (threadIdx.x + 1) & 63;
  __syncthreads();
  long long idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; i++) idx = sm[idx];
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = (double)idx;
}

// throughput on one SM: W warps per SMSP (blockDim = 128 * W), ILP chains per thread
template <int ILP>
__global__ void thr_dfma(double* out, long long* cyc, double a, double b)
{
  double v[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) v[j] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < kIters; i++)
#pragma unroll
    for (int j = 0; j < ILP; j++) v[j] = fma(v[j], a, b);
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += v[j];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = s;
}
template <int ILP>
__global__ void thr_dmma(double* out, long long* cyc, double a, double b)
{
  double v[ILP][2];
#pragma unroll
  for (int j = 0; j < ILP; j++) v[j][0] = v[j][1] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < kIters; i++)
#pragma unroll
    for (int j = 0; j < ILP; j++) dmma(v[j][0], v[j][1], a, b);
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += v[j][0] + v[j][1];
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = s;
}

template <class F>
static void run(const char* name, F launch, int ops_per_iter = 1)
{
  long long* cyc;
  cudaMallocManaged(&cyc, 8);
  launch(cyc);  // warm
  cudaDeviceSynchronize();
  launch(cyc);
  cudaDeviceSynchronize();
  printf("%-34s %8.2f cycles/op\n", name, (double)cyc[0] / (kIters * (double)ops_per_iter));
  cudaFree(cyc);
}

int main()
{
  double* out;
  cudaMalloc(&out, 8 * 4096);
  run("DFMA dependent", [&](long long* c) { lat_dfma<<<1, 32>>>(out, c, 1.0000001, 1e-9); });
  run("DADD dependent", [&](long long* c) { lat_dadd<<<1, 32>>>(out, c, 1e-9); });
  run("DMMA dependent (same acc)", [&](long long* c) { lat_dmma<<<1, 32>>>(out, c, 1e-9, 1e-9); });
  run("SHFL.64 dependent", [&](long long* c) { lat_shfl<<<1, 32>>>(out, c); });
  run("SHFL.64 + DADD dependent", [&](long long* c) { lat_shfl_add<<<1, 32>>>(out, c); });
  run("LDS.64 dependent", [&](long long* c) { lat_lds<<<1, 32>>>(out, c); });
  for (int w = 1; w <= 8; w *= 2)
  {
    char nm[64];
    snprintf(nm, 64, "DFMA %d warp/SMSP ILP1 (per op)", w);
    run(nm, [&](long long* c) { thr_dfma<1><<<1, 128 * w>>>(out, c, 1.0000001, 1e-9); }, w);
    snprintf(nm, 64, "DFMA %d warp/SMSP ILP2 (per op)", w);
    run(nm, [&](long long* c) { thr_dfma<2><<<1, 128 * w>>>(out, c, 1.0000001, 1e-9); }, 2 * w);
    snprintf(nm, 64, "DFMA %d warp/SMSP ILP4 (per op)", w);
    run(nm, [&](long long* c) { thr_dfma<4><<<1, 128 * w>>>(out, c, 1.0000001, 1e-9); }, 4 * w);
    snprintf(nm, 64, "DFMA %d warp/SMSP ILP8 (per op)", w);
    run(nm, [&](long long* c) { thr_dfma<8><<<1, 128 * w>>>(out, c, 1.0000001, 1e-9); }, 8 * w);
    snprintf(nm, 64, "DMMA %d warp/SMSP ILP1 (per op)", w);
    run(nm, [&](long long* c) { thr_dmma<1><<<1, 128 * w>>>(out, c, 1e-9, 1e-9); }, w);
    snprintf(nm, 64, "DMMA %d warp/SMSP ILP2 (per op)", w);
    run(nm, [&](long long* c) { thr_dmma<2><<<1, 128 * w>>>(out, c, 1e-9, 1e-9); }, 2 * w);
    snprintf(nm, 64, "DMMA %d warp/SMSP ILP4 (per op)", w);
    run(nm, [&](long long* c) { thr_dmma<4><<<1, 128 * w>>>(out, c, 1e-9, 1e-9); }, 4 * w);
  }
  return 0;
}
