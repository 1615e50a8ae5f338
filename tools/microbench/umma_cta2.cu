// Probe + microbenchmark (B200, sm_100a): tcgen05.mma.cta_group::2 kind::i8 (M = 256 over a CTA pair) with the operand layout of
// the RX-SSB-f32 tensor-core kernel (no-swizzle K-major core matrices, A row groups aliased with SBO = 768).
//   1. correctness: which half of B does each CTA of the pair supply, where do the rows of D land, does the aliased A layout work;
//   2. rate: clocks per MMA as a function of N for cta_group::2 (A: 128 rows per CTA, B: N / 2 rows per CTA) against cta_group::1
//      at the same N — an SS-mode MMA is bound by the fetch of its operands from shared memory (tools/microbench/umma_rate.cu:
//      (4 KB of A + 32 N bytes of B) at ~74 B/clk), so halving the B bytes each SM reads should show up here.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_cta2 tools/microbench/umma_cta2.cu && /tmp/umma_cta2
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }
__device__ __forceinline__ uint64_t make_desc (uint32_t addr, uint32_t lbo, uint32_t sbo)
{
  return (uint64_t) ((addr & 0x3FFFFu) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc (int M, int N, int a_signed, int b_signed)
{
  return (2u << 4) | ((uint32_t) a_signed << 7) | ((uint32_t) b_signed << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait (uint64_t *bar, unsigned parity)
{
  asm volatile ("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n"
                ::"r"(smem_u32 (bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync ()
{
  asm volatile ("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank () { uint32_t r; asm volatile ("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tmem_ld16 (uint32_t addr, uint32_t *v)
{
  asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr));
  asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Cfg { uint32_t n, a_lbo, a_sbo, a_step, b_step, ksteps, iters, a_bytes, b_bytes, cta2; };

// ------------------------------------------------------------------------------------------------------------------------------
// one kernel for both purposes: each CTA copies its own A (a_bytes) and its own share of B (b_bytes) into shared memory, the
// leader issues iters x ksteps MMAs (the first of all fresh, the rest accumulating), both CTAs dump their 128 TMEM lanes
// ------------------------------------------------------------------------------------------------------------------------------
template <int kCta2>
__global__ void __launch_bounds__ (128, 1) mma_kernel (Cfg c, const uint8_t *A, const uint8_t *B, int *D, long long *clk)
{
  extern __shared__ __align__ (1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = kCta2 ? cluster_rank () : 0u;
  const uint32_t pair = kCta2 ? blockIdx.x / 2 : blockIdx.x;
  unsigned char *sA = smem, *sB = smem + 96 * 1024;
  const uint8_t *gA = A + (size_t) (kCta2 ? rank : 0) * c.a_bytes, *gB = B + (size_t) (kCta2 ? rank : 0) * c.b_bytes;
  for (uint32_t i = tid; i < c.a_bytes / 16; i += 128) reinterpret_cast<uint4 *> (sA)[i] = reinterpret_cast<const uint4 *> (gA)[i];
  for (uint32_t i = tid; i < c.b_bytes / 16; i += 128) reinterpret_cast<uint4 *> (sB)[i] = reinterpret_cast<const uint4 *> (gB)[i];
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0)
  {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&bar)) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0)
  {
    if (kCta2)
    {
      asm volatile ("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32 (&tmem_s)) : "memory");
      asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    else
    {
      asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32 (&tmem_s)) : "memory");
      asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads ();
  if (kCta2) cluster_sync ();
  asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  long long t0 = 0;
  if (warp == 0 && rank == 0)
  {
    const uint32_t idesc = make_idesc (kCta2 ? 256 : 128, (int) c.n, 1, 1);
    const uint32_t a0 = smem_u32 (sA), b0 = smem_u32 (sB);
    t0 = clock64 ();
    uint32_t elected;
    asm volatile ("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(elected));
    if (elected)
    {
      for (uint32_t it = 0; it < c.iters; it++)
        for (uint32_t ks = 0; ks < c.ksteps; ks++)
        {
          const uint64_t da = make_desc (a0 + ks * c.a_step, c.a_lbo, c.a_sbo), db = make_desc (b0 + ks * c.b_step, 128, 256);
          const uint32_t acc = (it | ks) ? 1u : 0u;
          if (kCta2)
            asm volatile ("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile ("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      if (kCta2)
        asm volatile ("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32 (&bar)), "h"((uint16_t) 3) : "memory");
      else
        asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32 (&bar)) : "memory");
    }
    __syncwarp ();
  }
  mbar_wait (&bar, 0);                                    // both CTAs: the commit is multicast to the pair
  if (tid == 0 && rank == 0) clk[pair] = clock64 () - t0;
  asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (D)
    for (uint32_t c0 = 0; c0 < c.n; c0 += 16)
    {
      uint32_t v[16];
      tmem_ld16 (tmem + ((uint32_t) (warp * 32) << 16) + c0, v);
      for (int j = 0; j < 16; j++) D[((size_t) (pair * 2 + rank) * 128 + warp * 32 + lane) * c.n + c0 + j] = (int) v[j];
    }
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads ();
  if (kCta2) cluster_sync ();
  if (warp == 0)
  {
    if (kCta2) asm volatile ("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static cudaError_t launch (bool cta2, int ctas, const Cfg &c, const uint8_t *dA, const uint8_t *dB, int *dD, long long *dClk)
{
  const size_t smem = 200 * 1024;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3 ((unsigned) ctas); lc.blockDim = dim3 (128); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cta2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  if (cta2)
  {
    cudaFuncSetAttribute (mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaLaunchKernelEx (&lc, mma_kernel<1>, c, dA, dB, dD, dClk);
  }
  else
  {
    cudaFuncSetAttribute (mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaLaunchKernelEx (&lc, mma_kernel<0>, c, dA, dB, dD, dClk);
  }
  return cudaDeviceSynchronize ();
}

// byte (row r, k) of a K-major no-swizzle operand
static inline size_t addr_of (uint32_t off, uint32_t lbo, uint32_t sbo, int r, int k) { return off + (size_t) (r % 8) * 16 + (size_t) (r / 8) * sbo + (size_t) (k / 16) * lbo + (k % 16); }

int main ()
{
  srand (4321);
  int fails = 0;
  // ------------------------------- 1. correctness, N = 160, two K-steps, plain and aliased A -------------------------------
  for (int aliased = 0; aliased < 2; aliased++)
  {
    const int N = 160, KS = 2;
    Cfg c{};
    c.n = N; c.a_lbo = 128; c.a_sbo = aliased ? 768 : 256; c.a_step = aliased ? 256 : 4096; c.b_step = (N / 2 / 8) * 256; c.ksteps = KS; c.iters = 1;
    c.a_bytes = 16 * 1024; c.b_bytes = KS * c.b_step; c.cta2 = 1;
    std::vector<uint8_t> A (2 * c.a_bytes), B (2 * c.b_bytes);
    for (auto &x : A) x = (uint8_t) rand (); for (auto &x : B) x = (uint8_t) rand ();
    uint8_t *dA, *dB; int *dD; long long *dClk;
    cudaMalloc (&dA, A.size ()); cudaMalloc (&dB, B.size ()); cudaMalloc (&dD, 2 * 128 * N * 4); cudaMalloc (&dClk, 8 * 148);
    cudaMemcpy (dA, A.data (), A.size (), cudaMemcpyHostToDevice); cudaMemcpy (dB, B.data (), B.size (), cudaMemcpyHostToDevice);
    cudaMemset (dD, 0xEE, 2 * 128 * N * 4);
    cudaError_t e = launch (true, 2, c, dA, dB, dD, dClk);
    if (e != cudaSuccess) { printf ("correctness launch (aliased=%d): CUDA error %s\n", aliased, cudaGetErrorString (e)); return 1; }
    std::vector<int> D (2 * 128 * N);
    cudaMemcpy (D.data (), dD, D.size () * 4, cudaMemcpyDeviceToHost);
    // hypotheses: H0: B rows [0, N/2) from CTA 0, [N/2, N) from CTA 1; H1: the other way round
    for (int hyp = 0; hyp < 2; hyp++)
    {
      long bad = 0;
      for (int rank = 0; rank < 2; rank++) for (int r = 0; r < 128; r++) for (int n = 0; n < N; n++)
      {
        const int owner = (n < N / 2) ? hyp : 1 - hyp, nl = n % (N / 2);
        long s = 0;
        for (int ks = 0; ks < KS; ks++) for (int k = 0; k < 32; k++)
          s += (long) (int8_t) A[(size_t) rank * c.a_bytes + addr_of (ks * c.a_step, c.a_lbo, c.a_sbo, r, k)] *
               (int8_t) B[(size_t) owner * c.b_bytes + addr_of (ks * c.b_step, 128, 256, nl, k)];
        if (s != D[((size_t) rank * 128 + r) * N + n]) { if (bad < 3) printf ("    hyp %d rank %d r %d n %d want %ld got %d\n", hyp, rank, r, n, s, D[((size_t) rank * 128 + r) * N + n]); bad++; }
      }
      printf ("cta_group::2 M=256 N=%d, A %s, hypothesis %d (B rows [0,N/2) from CTA %d): %ld mismatches of %d\n", N, aliased ? "aliased (SBO 768)" : "plain", hyp, hyp, bad, 2 * 128 * N);
      if (hyp == 0 && bad) fails++;
    }
    cudaFree (dA); cudaFree (dB); cudaFree (dD); cudaFree (dClk);
  }
  // ------------------------------- 2. rate -------------------------------
  {
    uint8_t *dA, *dB; long long *dClk;
    cudaMalloc (&dA, 64 * 1024); cudaMalloc (&dB, 256 * 1024); cudaMalloc (&dClk, 8 * 148);
    cudaMemset (dA, 1, 64 * 1024); cudaMemset (dB, 1, 256 * 1024);
    for (int ctas : { 2, 148 })
      for (uint32_t n : { 64u, 96u, 128u, 160u, 192u, 256u })
        for (int cta2 = 0; cta2 < 2; cta2++)
        {
          Cfg c{};
          const uint32_t nb = cta2 ? n / 2 : n;                     // B rows held by one CTA
          c.n = n; c.a_lbo = 128; c.a_sbo = 768; c.a_step = 256; c.b_step = (nb / 8) * 256; c.ksteps = 11; c.iters = 40;
          c.a_bytes = 16 * 1024; c.b_bytes = c.ksteps * c.b_step; c.cta2 = cta2;
          cudaError_t e = launch (cta2 != 0, ctas, c, dA, dB, nullptr, dClk);
          if (e != cudaSuccess) { printf ("rate launch: CUDA error %s\n", cudaGetErrorString (e)); return 1; }
          const int np = cta2 ? ctas / 2 : ctas;
          std::vector<long long> h (np); cudaMemcpy (h.data (), dClk, np * 8, cudaMemcpyDeviceToHost);
          double avg = 0; for (auto v : h) avg += (double) v; avg /= np;
          const double per = avg / (c.iters * c.ksteps);
          printf ("%s  N=%3u  %7.1f clk/MMA  (%6.0f MAC/clk/SM; operand bytes per SM and MMA: %5u)  (%d CTAs)\n", cta2 ? "cta_group::2 M=256" : "cta_group::1 M=128", n, per,
                  128.0 * n * 32 / per, 4096u + 32u * nb, ctas);
        }
  }
  printf (fails ? "PROBE FAILED\n" : "PROBE OK\n");
  return fails ? 1 : 0;
}
