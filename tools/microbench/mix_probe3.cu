// Does a busy FP64 pipe (DMMA / DFMA) take issue slots away from OTHER warps' integer / shared-memory instructions on B200?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mix_probe3 mix_probe3.cu
// Per CTA (one per SM): warps with (warp % 2 == 0) run role A, the others role B; both roles alone, then together.
// If the two instruction classes only shared the scheduler's one-instruction-per-cycle issue port, the mixed run would
// take max(tA, tB) (+ a little); if the FP64 pipe blocked the port while busy it would take tA + tB.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// role: 0 idle, 1 DMMA (32 per iteration = 512 pipe cycles), 2 DFMA (256 = 512+ pipe cycles), 3 IMAD (512, ILP 8),
//       4 LDS (128 x 64-bit, conflict-free), 5 FSEL/ISETP mix (512)
__device__ __forceinline__ double run_role(int role, int iters, double* sm, int lane)
{
  double out = 0.0;
  if (role == 1)
  {
    double d[4][2];
    for (int i = 0; i < 4; i++) d[i][0] = d[i][1] = 1.0;
    const double a = 1.0 + 1e-9 * lane, b = 1e-9;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) dmma(d[i][0], d[i][1], a, b);
    for (int i = 0; i < 4; i++) out += d[i][0] + d[i][1];
  }
  else if (role == 2)
  {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = 1.0 + i + lane;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 32; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], 1.0000001, 1e-9);
    for (int i = 0; i < 8; i++) out += a[i];
  }
  else if (role == 3)
  {
    unsigned v[8];
    for (int i = 0; i < 8; i++) v[i] = lane * 7 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 64; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = v[i] * 1664525u + (unsigned)r;
    for (int i = 0; i < 8; i++) out += v[i];
  }
  else if (role == 4)
  {
    double acc[4] = { 0, 0, 0, 0 };
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 128; r++) acc[r & 3] += *(volatile double*)&sm[((it + r) & 63) * 32 + lane];
    out = acc[0] + acc[1] + acc[2] + acc[3];
  }
  else if (role == 5)
  {
    float f[8];
    for (int i = 0; i < 8; i++) f[i] = 1.0f + i + lane;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int r = 0; r < 64; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = f[i] > (float)r ? f[i] - 0.5f : f[i] + 1.5f;
    for (int i = 0; i < 8; i++) out += f[i];
  }
  return out;
}

__global__ void __launch_bounds__(512) k_mix(double* out, int iters, int roleA, int roleB)
{
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1e-9 * i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double v = run_role((warp & 1) ? roleB : roleA, iters, sm, lane);
  if (v == 123.456) out[0] = v;
}

double time_it(double* d, int threads, int iters, int a, int b)
{
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++)
  {
    cudaEventRecord(e0); k_mix<<<sms, threads>>>(d, iters, a, b); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  return best;
}

int main()
{
  double* d; cudaMalloc(&d, 8);
  const char* names[6] = { "idle", "DMMA", "DFMA", "IMAD", "LDS.64", "FSEL/FSETP" };
  const int iters = 2000;
  for (int threads = 256; threads <= 512; threads *= 2)
  {
    printf("--- %d warps per SM (even warps role A, odd warps role B)\n", threads / 32);
    for (int a = 1; a <= 2; a++)
      for (int b = 3; b <= 5; b++)
      {
        const double ta = time_it(d, threads, iters, a, 0), tb = time_it(d, threads, iters, 0, b), tm = time_it(d, threads, iters, a, b);
        printf("%-5s alone %.3f ms | %-10s alone %.3f ms | together %.3f ms   (max %.3f, sum %.3f)\n", names[a], ta, names[b], tb, tm,
               ta > tb ? ta : tb, ta + tb);
      }
  }
  return 0;
}
