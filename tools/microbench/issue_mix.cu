// Microbenchmark (B200, sm_100a): does a packed FP32x2 instruction cost one issue slot or two? Measures the warp-instruction
// rate per SM of (A) FFMA2 alone, (B) FFMA2 interleaved 1:1 with integer LOP3, (C) LOP3 alone, (D) FFMA2 : LOP3 = 1:2,
// (E) scalar FFMA interleaved 1:1 with LOP3, at 4 and 8 warps per scheduler. Used to read the RX kernel's issue_active figure.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/issue_mix tools/microbench/issue_mix.cu && /tmp/issue_mix
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2 (u64 a, u64 b, u64 c) { u64 r; asm volatile ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1 (float a, float b, float c) { float r; asm volatile ("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned lop (unsigned a, unsigned b, unsigned c) { unsigned r; asm volatile ("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <int MODE> __global__ void k (float *out, int iters, float seed)
{
  float a[8]; u64 p[8]; unsigned q[16];
  for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; p[i] = ((u64) __float_as_uint (a[i]) << 32) | __float_as_uint (a[i] * 0.5f); }
  for (int i = 0; i < 16; i++) q[i] = threadIdx.x * 7 + i;
  u64 m2[2], c2[2]; unsigned z = threadIdx.x, w = blockIdx.x;
  for (int i = 0; i < 2; i++) { float m = 0.999f + 1e-4f * i + seed * 1e-9f, c = 0.001f * (i + 1) + seed * 1e-9f; m2[i] = ((u64) __float_as_uint (m) << 32) | __float_as_uint (m); c2[i] = ((u64) __float_as_uint (c) << 32) | __float_as_uint (c); }
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
      if (MODE == 0 || MODE == 1 || MODE == 3) p[i] = fma2 (p[i], m2[i & 1], c2[i & 1]);
      if (MODE == 4) a[i] = fma1 (a[i], __uint_as_float ((unsigned) m2[i & 1]), __uint_as_float ((unsigned) c2[i & 1]));
      if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 4) q[i] = lop (q[i], z, w);
      if (MODE == 3 || MODE == 2) q[i + 8] = lop (q[i + 8], z, w);
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i] + __uint_as_float ((unsigned) p[i]) + __uint_as_float ((unsigned) (p[i] >> 32));
  unsigned t = 0; for (int i = 0; i < 16; i++) t ^= q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float) t;
}

template <int MODE> void run (const char *name, int instr_per_iter, int warps_per_sched)
{
  const int threads = 32 * 4 * warps_per_sched;
  float *out; cudaMalloc (&out, 148 * threads * sizeof (float));
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  const int iters = 40000;
  k<MODE><<<148, threads>>> (out, 100, 1.0f);
  cudaEventRecord (e0); k<MODE><<<148, threads>>> (out, iters, 1.0f); cudaEventRecord (e1); cudaEventSynchronize (e1);
  float ms; cudaEventElapsedTime (&ms, e0, e1);
  double instr = (double) 148 * threads / 32 * iters * instr_per_iter;
  printf ("%-34s %d warps/sched %8.3f ms  %6.2f warp-instr/clk/SM (of 4.00, at 1965 MHz)\n", name, warps_per_sched, ms, instr / (ms * 1e-3) / 1.965e9 / 148);
  cudaFree (out);
}
int main ()
{
  for (int w = 4; w <= 8; w += 4)
  {
    run<0> ("A FFMA2 alone", 8, w);
    run<1> ("B FFMA2 : LOP3 = 1 : 1", 16, w);
    run<2> ("C LOP3 alone", 16, w);
    run<3> ("D FFMA2 : LOP3 = 1 : 2", 24, w);
    run<4> ("E FFMA (scalar) : LOP3 = 1 : 1", 16, w);
  }
  return 0;
}
