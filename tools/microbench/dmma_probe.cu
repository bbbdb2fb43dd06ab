// DMMA / DFMA throughput probes for B200 (sm_100a).  Build:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
// Prints TFLOP/s (2 flop per FMA) for several operand patterns.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// NACC independent accumulators, constant operands
template <int NACC>
__global__ void __launch_bounds__(256) k_const(double* out, int iters)
{
  double d[NACC][2];
  for (int i = 0; i < NACC; i++) d[i][0] = d[i][1] = 1.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int r = 0; r < 32 / NACC; r++)
#pragma unroll
      for (int i = 0; i < NACC; i++) dmma(d[i][0], d[i][1], a, b);
  double s = 0;
  for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}

// 4 accumulators, A reused over 4 tiles, B distinct per tile, operands rotate through 8 registers
__global__ void __launch_bounds__(256) k_regs(double* out, int iters)
{
  double d[4][2], a[8], b[8];
  for (int i = 0; i < 4; i++) d[i][0] = d[i][1] = 1.0;
  for (int i = 0; i < 8; i++) { a[i] = 1.0 + 1e-9 * (threadIdx.x + i); b[i] = 1e-9 * (i + 1); }
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int s = 0; s < 8; s++)
#pragma unroll
      for (int t = 0; t < 4; t++) dmma(d[t][0], d[t][1], a[s], b[(s + t) & 7]);
  double s = 0;
  for (int i = 0; i < 4; i++) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}

// as the phi_k kernel: A from registers, B from shared memory (conflict-free pattern)
__global__ void __launch_bounds__(512) k_lds(double* out, int iters)
{
  __shared__ double sm[128 * 36];
  for (int i = threadIdx.x; i < 128 * 36; i += blockDim.x) sm[i] = 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  double d[4][2], a[8];
  for (int i = 0; i < 4; i++) d[i][0] = d[i][1] = 1.0;
  for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double* base = sm + q * 36 + g;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int s = 0; s < 8; s++)
#pragma unroll
      for (int t = 0; t < 4; t++) dmma(d[t][0], d[t][1], a[s], base[(4 * s) * 36 + 8 * t]);
  double s = 0;
  for (int i = 0; i < 4; i++) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}

// DFMA with NCH independent chains, constant multiplier
template <int NCH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters)
{
  double a[NCH];
  for (int i = 0; i < NCH; i++) a[i] = 1.0 + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int r = 0; r < 64 / NCH; r++)
#pragma unroll
      for (int i = 0; i < NCH; i++) a[i] = fma(a[i], m, c);
  double s = 0;
  for (int i = 0; i < NCH; i++) s += a[i];
  if (s == 123.456) out[0] = s;
}

// DFMA with 3 distinct register operands per instruction
__global__ void __launch_bounds__(256) k_dfma3(double* out, int iters)
{
  double a[8], m[8], c[8];
  for (int i = 0; i < 8; i++) { a[i] = 1.0 + i + threadIdx.x; m[i] = 1.0 + 1e-9 * i; c[i] = 1e-9 * i; }
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fma(a[i], m[(i + r) & 7], c[(i + 2 * r) & 7]);
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 123.456) out[0] = s;
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
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d; cudaMalloc(&d, 8);
  const int iters = 4096;
  auto tf_dmma = [&](double s, int grid, int threads, int per_iter) {
    return 2.0 * 256.0 * per_iter * iters * (double)grid * (threads / 32) / s / 1e12; };
  for (int wps = 1; wps <= 4; wps *= 2)  // warps per SMSP: 4, 8, 16 warps per SM (x1 CTA of 128/256/512)
  {
    const int threads = 128 * wps, grid = sms;
    printf("--- %d warps/SM (1 CTA x %d threads per SM)\n", threads / 32, threads);
    printf("dmma const  1 acc : %6.2f TF\n", tf_dmma(time_it([&] { k_const<1><<<grid, threads>>>(d, iters); }), grid, threads, 32));
    printf("dmma const  2 acc : %6.2f TF\n", tf_dmma(time_it([&] { k_const<2><<<grid, threads>>>(d, iters); }), grid, threads, 32));
    printf("dmma const  4 acc : %6.2f TF\n", tf_dmma(time_it([&] { k_const<4><<<grid, threads>>>(d, iters); }), grid, threads, 32));
    printf("dmma const  8 acc : %6.2f TF\n", tf_dmma(time_it([&] { k_const<8><<<grid, threads>>>(d, iters); }), grid, threads, 32));
    printf("dmma regs   4 acc : %6.2f TF\n", tf_dmma(time_it([&] { k_regs<<<grid, threads>>>(d, iters); }), grid, threads, 32));
    printf("dmma lds    4 acc : %6.2f TF\n", tf_dmma(time_it([&] { k_lds<<<grid, threads>>>(d, iters); }), grid, threads, 32));
  }
  for (int mult = 1; mult <= 8; mult *= 2)
  {
    const int grid = sms * mult, threads = 256;
    auto tf = [&](double s) { return 2.0 * 64.0 * iters * (double)grid * threads / s / 1e12; };
    printf("--- %d CTAs x 256 threads per SM\n", mult);
    printf("dfma 1 chain      : %6.2f TF\n", tf(time_it([&] { k_dfma<1><<<grid, threads>>>(d, iters); })));
    printf("dfma 2 chains     : %6.2f TF\n", tf(time_it([&] { k_dfma<2><<<grid, threads>>>(d, iters); })));
    printf("dfma 4 chains     : %6.2f TF\n", tf(time_it([&] { k_dfma<4><<<grid, threads>>>(d, iters); })));
    printf("dfma 8 chains     : %6.2f TF\n", tf(time_it([&] { k_dfma<8><<<grid, threads>>>(d, iters); })));
    printf("dfma 8ch 3 regs   : %6.2f TF\n", tf(time_it([&] { k_dfma3<<<grid, threads>>>(d, iters); })));
  }
  return 0;
}
