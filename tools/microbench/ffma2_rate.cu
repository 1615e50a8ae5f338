// Microbenchmark (B200, sm_100a): issue/throughput of packed FFMA2 vs scalar FFMA, to know which roof the fused RX
// kernel can reach. Build+run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 tools/microbench/ffma2_rate.cu && /tmp/ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2 (u64 a, u64 b, u64 c) { u64 r; asm volatile ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2 (u64 a, u64 b) { u64 r; asm volatile ("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1 (float a, float b, float c) { float r; asm volatile ("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE> __global__ void k (float *out, int iters, float seed)
{
  float a[8]; u64 p[8];
  for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; p[i] = ((u64) __float_as_uint (a[i]) << 32) | __float_as_uint (a[i] * 0.5f); }
  const float m = 0.999f, c = 0.001f;
  const u64 m2 = ((u64) __float_as_uint (m) << 32) | __float_as_uint (m), c2 = ((u64) __float_as_uint (c) << 32) | __float_as_uint (c);
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
      if (MODE == 0) a[i] = fma1 (a[i], m, c);
      if (MODE == 1) p[i] = fma2 (p[i], m2, c2);
      if (MODE == 2) p[i] = add2 (p[i], c2);
      if (MODE == 3) { a[i] = fma1 (a[i], m, c); p[i] = fma2 (p[i], m2, c2); }
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float ((unsigned) p[i]) + __uint_as_float ((unsigned) (p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run (const char *name, int lanes_per_instr, int instr_per_iter)
{
  float *out; cudaMalloc (&out, 148 * 8 * 256 * sizeof (float));
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>> (out, 100, 1.0f);
  cudaEventRecord (e0); k<MODE><<<148 * 8, 256>>> (out, iters, 1.0f); cudaEventRecord (e1); cudaEventSynchronize (e1);
  float ms; cudaEventElapsedTime (&ms, e0, e1);
  double instr = (double) 148 * 8 * 256 / 32 * iters * instr_per_iter;      // warp instructions
  double clk = 1.965e9;
  printf ("%-28s %8.3f ms  %7.2f warp-instr/clk/SM (at 1965 MHz)  %7.1f G lane-FMA-ops/s/SM-equivalent: %6.1f lanes/clk/SM\n", name, ms,
          instr / (ms * 1e-3) / clk / 148, instr * 32 * lanes_per_instr / (ms * 1e-3) / 1e9, instr * 32 * lanes_per_instr / (ms * 1e-3) / clk / 148);
  cudaFree (out);
}
int main ()
{
  run<0> ("FFMA  (scalar)", 1, 8);
  run<1> ("FFMA2 (packed)", 2, 8);
  run<2> ("FADD2 (packed)", 2, 8);
  run<3> ("FFMA + FFMA2 interleaved", 1, 16);   // lanes: 8*1 + 8*2 per 16 instr -> reported per-instr average of 1.5 below is not exact
  return 0;
}
