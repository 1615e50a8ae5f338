// Microbenchmark (B200, sm_100a): issue-to-completion rate of tcgen05.mma (cta_group::1, M = 128, SS mode) as a function of
// the shared-memory layout of the operands (no swizzle / 32 / 64 / 128-byte swizzle), N, and kind (i8 / f16).
// The RX-SSB-f32 tensor-core kernel needs a layout whose 8-row groups may alias (row-group stride < group size):
// that rules out the 128-byte swizzle, so the question is what the other layouts cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_rate tools/microbench/umma_rate.cu && /tmp/umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }
struct Cfg { uint32_t layout, a_lbo, a_sbo, b_lbo, b_sbo, n, f16, iters, a_step, b_step, ksteps, alt; };

__global__ void __launch_bounds__ (128, 1) rate (Cfg c, long long *out)
{
  extern __shared__ __align__ (1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4 *> (smem)[i] = make_uint4 (0, 0, 0, 0);
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0)
  {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&bar)) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32)
  {
    asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32 (&tmem_s)) : "memory");
    asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads ();
  asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (threadIdx.x < 32)
  {
    // the warp stays converged and ONE elected lane issues (descriptors in uniform registers): under `if (threadIdx.x == 0)`
    // the compiler wraps every tcgen05.mma in a vote / broadcast loop and the ISSUE, not the tensor pipe, sets the pace
    const uint32_t idesc = c.f16 ? ((1u << 4) | ((c.n >> 3) << 17) | (8u << 24))                    // f16 x f16 -> f32
                                 : ((2u << 4) | (1u << 7) | (1u << 10) | ((c.n >> 3) << 17) | (8u << 24));   // s8 x s8 -> s32
    const uint64_t hi = ((uint64_t) (c.layout & 7) << 61) | (1ull << 46);
    const uint32_t a0 = smem_u32 (smem), b0 = smem_u32 (smem) + 96 * 1024;
    const long long t0 = clock64 ();
    uint32_t elected;
    asm volatile ("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
    if (elected)
    {
      for (uint32_t it = 0; it < c.iters; it++)
        for (uint32_t ks = 0; ks < c.ksteps; ks++)
        {
          const uint64_t da = hi | (uint64_t) (((a0 + ks * c.a_step) & 0x3FFFF) >> 4) | ((uint64_t) (c.a_lbo >> 4) << 16) | ((uint64_t) (c.a_sbo >> 4) << 32);
          const uint64_t db = hi | (uint64_t) (((b0 + ks * c.b_step) & 0x3FFFF) >> 4) | ((uint64_t) (c.b_lbo >> 4) << 16) | ((uint64_t) (c.b_sbo >> 4) << 32);
          const uint32_t d = tmem + ((c.alt && (ks & 1)) ? 256u : 0u);              // alt: two independent accumulators, alternately
          if (c.f16)
            asm volatile ("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile ("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        }
      asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32 (&bar)) : "memory");
    }
    __syncwarp ();
    asm volatile ("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra D;\n bra W;\n D:\n}\n" ::"r"(smem_u32 (&bar)) : "memory");
    if (threadIdx.x == 0) out[blockIdx.x] = clock64 () - t0;
  }
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads ();
  if (threadIdx.x < 32) asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static void run (const char *name, Cfg c, int ctas)
{
  long long *d; cudaMalloc (&d, 148 * 8);
  c.iters = 40;
  cudaFuncSetAttribute (rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate<<<ctas, 128, 200 * 1024>>> (c, d);
  cudaError_t e = cudaDeviceSynchronize ();
  if (e != cudaSuccess) { printf ("%-64s CUDA error %s\n", name, cudaGetErrorString (e)); exit (1); }
  long long h[148]; cudaMemcpy (h, d, ctas * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < ctas; i++) avg += (double) h[i]; avg /= ctas;
  const double per = avg / (c.iters * c.ksteps);
  const double macs = 128.0 * c.n * (c.f16 ? 16 : 32);
  printf ("%-64s N=%3u %s  %7.1f clk/MMA  %7.0f MAC/clk/SM  (%d CTAs)\n", name, c.n, c.f16 ? "f16" : "i8 ", per, macs / per, ctas);
  cudaFree (d);
}

int main ()
{
  // K per instruction = 32 bytes for both kinds. Operand tiles: A 128 rows, B n rows.
  for (int ctas : { 1, 148 })
  {
    for (uint32_t n : { 48u, 96u, 144u, 192u, 256u })
    {
      run ("no swizzle, A aliased, two independent accumulators alternately", Cfg{ 0, 128, 768, 128, 256, n, 0, 0, 256, 4608, 11, 1 }, ctas);
      // no swizzle: core matrix 8 rows x 16 B; LBO = 128 between the two K chunks, SBO = 256 between row groups; K-step advances by 4608 (B) / 256 (A, as the kernel's aliased plane)
      run ("no swizzle, A aliased (SBO 768), step as in the kernel", Cfg{ 0, 128, 768, 128, 256, n, 0, 0, 256, 4608, 11 }, ctas);
      run ("no swizzle, A plain (SBO 256)", Cfg{ 0, 128, 256, 128, 256, n, 0, 0, 4096, 8192, 8 }, ctas);
      // 32-byte swizzle: 8 rows x 32 B = 256 B atoms; SBO = 256 (plain) or 768 (aliased: three 16-frame pieces per block)
      run ("swizzle 32B, A plain (SBO 256)", Cfg{ 6, 16, 256, 16, 256, n, 0, 0, 4096, 8192, 8 }, ctas);
      run ("swizzle 32B, A aliased (SBO 768)", Cfg{ 6, 16, 768, 16, 256, n, 0, 0, 256, 8192, 8 }, ctas);
      // 128-byte swizzle: 8 rows x 128 B atoms, SBO = 1024, K-step advances 32 B inside the row (4 steps per atom)
      run ("swizzle 128B, plain (SBO 1024), 4 K-steps inside one atom", Cfg{ 2, 16, 1024, 16, 1024, n, 0, 0, 32, 32, 4 }, ctas);
      run ("swizzle 128B, f16", Cfg{ 2, 16, 1024, 16, 1024, n, 1, 0, 32, 32, 4 }, ctas);
      run ("no swizzle, f16, plain", Cfg{ 0, 128, 256, 128, 256, n, 1, 0, 4096, 8192, 8 }, ctas);
    }
  }
  return 0;
}
