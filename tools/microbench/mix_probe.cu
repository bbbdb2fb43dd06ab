// Do DMMA and DFMA share one FP64 pipe on B200?  And what does the shared-memory data pipe sustain?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mix_probe mix_probe.cu
// k_mix: in every CTA, warps with (warp % 2 == 0) run DMMA chains, the others DFMA chains; the kernel reports the
// cycles both roles needed.  If the two instruction classes had separate pipes, the mixed run would take as
// long as the slower role alone; on one shared pipe the times add up.
// k_lds: LDS.64 / LDS.128 / STS.64 throughput per SM (conflict-free), in bytes per cycle.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA (same per-warp work as in 0 / 1)
__global__ void __launch_bounds__(512) k_mix(double* out, int iters, int mode)
{
  const int warp = threadIdx.x >> 5;
  const bool do_dmma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  if (do_dmma)
  {
    double d[4][2];
    for (int i = 0; i < 4; i++) d[i][0] = d[i][1] = 1.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) dmma(d[i][0], d[i][1], a, b);  // 32 DMMA = 512 pipe cycles
    double s = 0;
    for (int i = 0; i < 4; i++) s += d[i][0] + d[i][1];
    if (s == 123.456) out[0] = s;
  }
  else
  {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = 1.0 + i + threadIdx.x;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 32; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);  // 256 DFMA = 512 pipe cycles
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 123.456) out[0] = s;
  }
}

// kind 0: LDS.64, 1: LDS.128, 2: STS.64, 3: SHFL.32 (xor 1)
__global__ void __launch_bounds__(1024) k_lds(double* out, int iters, int kind)
{
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  unsigned v = threadIdx.x;
  if (kind == 0)
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 16; r++) acc += *(volatile double*)&sm[((it + r) & 63) * 32 + lane];
  else if (kind == 1)
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 16; r++)
      {
        double2 t;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"((unsigned)__cvta_generic_to_shared(&sm[((it + r) & 31) * 64 + 2 * lane])));
        acc += t.x + t.y;
      }
  else if (kind == 2)
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 16; r++) *(volatile double*)&sm[((it + r) & 63) * 32 + lane] = acc + r;
  else
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 16; r++) v += __shfl_xor_sync(0xffffffffu, v, 1 + (r & 15));
  if (acc == 123.456 || v == 0x12345678u) out[0] = acc + v;
}

template <class F>
double time_it(F launch)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++)
  {
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (rep && ms < best) best = ms;
  }
  return best * 1e-3;
}

int main()
{
  int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* d; cudaMalloc(&d, 8);
  const int iters = 2048;
  for (int threads = 256; threads <= 512; threads *= 2)
  {
    double t[3];
    for (int mode = 0; mode < 3; mode++) t[mode] = time_it([&] { k_mix<<<sms, threads>>>(d, iters, mode); });
    printf("%2d warps/SM: all-DMMA %.3f ms, all-DFMA %.3f ms, half/half %.3f ms  (one shared pipe predicts %.3f, separate pipes %.3f)\n",
           threads / 32, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, 0.5 * (t[0] + t[1]) * 1e3, 0.5 * (t[0] > t[1] ? t[0] : t[1]) * 1e3);
  }
  const char* names[4] = { "LDS.64 ", "LDS.128", "STS.64 ", "SHFL.32" };
  const int bytes[4] = { 256, 512, 256, 128 };
  for (int kind = 0; kind < 4; kind++)
    for (int threads = 256; threads <= 1024; threads *= 2)
    {
      const double s = time_it([&] { k_lds<<<sms, threads>>>(d, iters, kind); });
      const double instr = (double)iters * 16 * (threads / 32);  // warp instructions per SM
      const double cyc = s * khz * 1e3;
      printf("%s %4d threads/SM: %.2f cycles per warp instruction per SM, %.1f B/cycle/SM (at %d MHz nominal)\n", names[kind], threads,
             cyc / instr, bytes[kind] * instr / cyc, khz / 1000);
    }
  return 0;
}
