// Microbenchmark (B200, sm_100a): throughput of the legacy warp-level integer tensor-core path
// mma.sync.m16n8k32.s32.s8.s8.s32 (SASS IMMA), which the exact q15 FIR of the RX-SSB-q15 chain runs on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/imma tools/microbench/imma_rate.cu && /tmp/imma
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma (int (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
  asm volatile ("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC> __global__ void k (int *out, int iters)
{
  int d[NACC][4] = {}; unsigned a[4], b[2];
  for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 0x01010101u + i;
  b[0] = threadIdx.x; b[1] = ~threadIdx.x;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int j = 0; j < NACC; j++) imma (d[j], a, b);
  int s = 0; for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += d[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> void run (int warps)
{
  int *out; cudaMalloc (&out, 148 * warps * 32 * 4);
  cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
  const int iters = 20000;
  k<NACC><<<148, warps * 32>>> (out, 10);
  cudaEventRecord (e0); k<NACC><<<148, warps * 32>>> (out, iters); cudaEventRecord (e1); cudaEventSynchronize (e1);
  float ms; cudaEventElapsedTime (&ms, e0, e1);
  double mmas = (double) 148 * warps * iters * NACC;
  printf ("%d independent accumulators, %2d warps/SM: %8.3f ms  %6.3f IMMA/clk/SM  %7.1f int8 TOPS (2*16*8*32 per IMMA)\n", NACC, warps, ms,
          mmas / (ms * 1e-3) / 1.965e9 / 148, mmas * 2 * 16 * 8 * 32 / (ms * 1e-3) / 1e12);
  cudaFree (out);
}
int main () { run<4> (4); run<8> (4); run<8> (8); run<8> (16); run<6> (16); return 0; }
