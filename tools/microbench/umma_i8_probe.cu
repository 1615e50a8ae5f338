// Probe (B200, sm_100a): does tcgen05.mma kind::i8 do what the RX-SSB-f32 tensor-core FIR needs?
//   * no-swizzle K-major shared-memory descriptors with a caller-chosen LBO / SBO, including the ALIASED layout where
//     the 8-row groups of A overlap in shared memory (SBO = 6 chunks: row group q is the same byte plane 48 frames on),
//   * s8 x s8 and u8 x s8 products into int32 TMEM accumulators, accumulate flag, N = 48 / 96 / 144 sub-ranges of B,
//   * tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.x16.
// Test 1: one MMA, plain layout, both readings of (LBO, SBO). Test 2: a whole supertile of the FIR (23 MMAs) against an
// integer FIR computed on the host from the raw samples and 24-bit taps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/microbench/umma_i8_probe.cu && /tmp/umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }
__device__ __forceinline__ uint64_t make_desc (uint32_t addr, uint32_t lbo, uint32_t sbo)
{
  // SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48), layout type 0 (no swizzle) [61,64)
  return (uint64_t) ((addr & 0x3FFFFu) >> 4) | ((uint64_t) (lbo >> 4) << 16) | ((uint64_t) (sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc (int M, int N, int a_signed, int b_signed)
{
  // InstrDescriptor: c_format S32 = 2 [4,6), a_format [7,10), b_format [10,13) (0 = u8, 1 = s8), K-major both, N >> 3 [17,23), M >> 4 [24,29)
  return (2u << 4) | ((uint32_t) a_signed << 7) | ((uint32_t) b_signed << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8 (uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
  asm volatile ("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t *bar, unsigned parity)
{
  asm volatile ("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n"
                ::"r"(smem_u32 (bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16 (uint32_t addr, uint32_t *v)
{
  asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(addr));
  asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Op { uint32_t a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, n, a_signed, b_signed, col, acc; };
constexpr int kMaxOps = 32;
struct Params { const uint8_t *a; const uint8_t *b; uint32_t a_bytes, b_bytes; int *d; int n_ops; int cols_out; Op ops[kMaxOps]; };

__global__ void __launch_bounds__ (128, 1) probe (const __grid_constant__ Params P)
{
  extern __shared__ __align__ (1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char *sA = smem, *sB = smem + ((P.a_bytes + 1023u) & ~1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < P.a_bytes / 16; i += 128) reinterpret_cast<uint4 *> (sA)[i] = reinterpret_cast<const uint4 *> (P.a)[i];
  for (uint32_t i = tid; i < P.b_bytes / 16; i += 128) reinterpret_cast<uint4 *> (sB)[i] = reinterpret_cast<const uint4 *> (P.b)[i];
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0)
  {
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32 (&bar)) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0)
  {
    asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32 (&tmem_base_s)) : "memory");
    asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads ();
  asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0)
  {
    for (int i = 0; i < P.n_ops; i++)
    {
      const Op &o = P.ops[i];
      umma_i8 (tmem + o.col, make_desc (smem_u32 (sA) + o.a_off, o.a_lbo, o.a_sbo), make_desc (smem_u32 (sB) + o.b_off, o.b_lbo, o.b_sbo),
               make_idesc (128, (int) o.n, (int) o.a_signed, (int) o.b_signed), o.acc);
    }
    asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32 (&bar)) : "memory");
  }
  mbar_wait (&bar, 0);
  asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < P.cols_out; c0 += 16)
  {
    uint32_t v[16];
    tmem_ld16 (tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) c0, v);
    for (int j = 0; j < 16; j++) P.d[(warp * 32 + lane) * P.cols_out + c0 + j] = (int) v[j];
  }
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads ();
  if (warp == 0) asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

static int run (const std::vector<uint8_t> &A, const std::vector<uint8_t> &B, Params P, std::vector<int> &D)
{
  uint8_t *dA, *dB; int *dD;
  cudaMalloc (&dA, A.size ()); cudaMalloc (&dB, B.size ()); cudaMalloc (&dD, 128 * P.cols_out * 4);
  cudaMemcpy (dA, A.data (), A.size (), cudaMemcpyHostToDevice); cudaMemcpy (dB, B.data (), B.size (), cudaMemcpyHostToDevice);
  cudaMemset (dD, 0xEE, 128 * P.cols_out * 4);
  P.a = dA; P.b = dB; P.d = dD; P.a_bytes = (uint32_t) A.size (); P.b_bytes = (uint32_t) B.size ();
  const size_t smem = ((A.size () + 1023) & ~(size_t) 1023) + B.size () + 1024;
  cudaFuncSetAttribute (probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  probe<<<1, 128, smem>>> (P);
  cudaError_t e = cudaDeviceSynchronize ();
  if (e != cudaSuccess) { printf ("  CUDA error: %s\n", cudaGetErrorString (e)); return 1; }
  D.resize (128 * P.cols_out);
  cudaMemcpy (D.data (), dD, D.size () * 4, cudaMemcpyDeviceToHost);
  cudaFree (dA); cudaFree (dB); cudaFree (dD);
  return 0;
}

// address model under test: byte (row r, k) of a K-major no-swizzle operand
static inline size_t addr_of (uint32_t off, uint32_t lbo, uint32_t sbo, int r, int k) { return off + (size_t) (r % 8) * 16 + (size_t) (r / 8) * sbo + (size_t) (k / 16) * lbo + (k % 16); }

int main ()
{
  srand (1234);
  int fails = 0;
  // ---------------- test 1: one MMA (M=128, N=48, K=32), plain layout, s8 x s8, then u8 x s8 accumulate ----------------
  for (int variant = 0; variant < 2; variant++)
  {
    // A: 16 row groups x 2 k-chunks x 128 B; model: LBO = distance between k-chunks, SBO = distance between row groups
    const uint32_t a_lbo = 128, a_sbo = 256, b_lbo = 128, b_sbo = 256;
    std::vector<uint8_t> A (16 * 256 * 2), B (6 * 256);
    for (auto &x : A) x = (uint8_t) rand (); for (auto &x : B) x = (uint8_t) rand ();
    Params P{}; P.cols_out = 48; P.n_ops = 2;
    // variant 1 swaps the two fields in the descriptor (if the hardware reads them the other way round, variant 1 matches)
    P.ops[0] = Op{ 0, 0, variant ? a_sbo : a_lbo, variant ? a_lbo : a_sbo, variant ? b_sbo : b_lbo, variant ? b_lbo : b_sbo, 48, 1, 1, 0, 0 };
    P.ops[1] = P.ops[0]; P.ops[1].a_off = 16 * 256; P.ops[1].a_signed = 0; P.ops[1].acc = 1;
    std::vector<int> D;
    if (run (A, B, P, D)) { fails++; continue; }
    long bad = 0;
    for (int r = 0; r < 128; r++) for (int n = 0; n < 48; n++)
    {
      long s = 0;
      for (int k = 0; k < 32; k++)
      {
        s += (long) (int8_t) A[addr_of (0, a_lbo, a_sbo, r, k)] * (int8_t) B[addr_of (0, b_lbo, b_sbo, n, k)];
        s += (long) (uint8_t) A[addr_of (16 * 256, a_lbo, a_sbo, r, k)] * (int8_t) B[addr_of (0, b_lbo, b_sbo, n, k)];
      }
      if (s != D[r * 48 + n]) { if (bad < 4) printf ("    r=%d n=%d want %ld got %d\n", r, n, s, D[r * 48 + n]); bad++; }
    }
    printf ("test 1 variant %d (%s): %ld mismatches of %d\n", variant, variant ? "descriptor fields swapped" : "LBO = k-chunk stride, SBO = row-group stride", bad, 128 * 48);
    if (variant == 0 && bad) fails++;
  }
  // ---------------- test 2: one supertile of the FIR ----------------
  {
    // raw: 8 channels x 896 frames int16 I/Q (128 history + 768 new); taps: 129 complex, 24-bit, balanced base-256 digits
    const int J = 8, F = 896, NT = 129;
    std::vector<int16_t> raw (J * F * 2);
    for (auto &x : raw) x = (int16_t) (rand () & 0xFFFF);
    raw[0] = 32767; raw[1] = -32768; raw[2] = -32768; raw[3] = 32767;
    std::vector<int32_t> hr (NT), hi (NT);
    for (int d = 0; d < NT; d++) { hr[d] = (rand () % 16000000) - 8000000; hi[d] = (rand () % 16000000) - 8000000; }
    hr[0] = 8323071; hi[0] = -8323072;     // the extremes the digit split must hold: 2^23 - 2^16 - 1
    auto digits = [] (int32_t h, int8_t *dg) { int32_t l = ((h + 128) & 255) - 128; int32_t r1 = (h - l) >> 8; int32_t m = ((r1 + 128) & 255) - 128; int32_t hh = (r1 - m) >> 8; dg[0] = (int8_t) hh; dg[1] = (int8_t) m; dg[2] = (int8_t) l; if (hh < -128 || hh > 127) printf ("digit overflow\n"); };
    // A planes: [plane][chunk 0..111][channel j][16 B]: chunk = 8 frames, bytes (I, Q) per frame
    std::vector<uint8_t> A (2 * 112 * 128), B (11 * 18 * 256);
    for (int j = 0; j < J; j++) for (int f = 0; f < F; f++) for (int rail = 0; rail < 2; rail++)
    {
      const uint16_t v = (uint16_t) raw[(j * F + f) * 2 + rail];
      const size_t o = (size_t) (f / 8) * 128 + j * 16 + (f % 8) * 2 + rail;
      A[o] = (uint8_t) (v >> 8); A[112 * 128 + o] = (uint8_t) (v & 255);
    }
    // B: [k-step 0..10][row group 0..17][k-chunk 0..1][row 0..7][16 B]; row n' = digit * 48 + n
    for (int ks = 0; ks < 11; ks++) for (int dg = 0; dg < 3; dg++) for (int n = 0; n < 48; n++) for (int kk = 0; kk < 32; kk++)
    {
      const int m = 32 * ks + kk, f = m / 2, rail = m & 1, d = 128 + n - f;
      int8_t v = 0;
      if (d >= 0 && d <= 128) { int8_t g[3]; digits (rail == 0 ? hr[d] : -hi[d], g); v = g[dg]; }
      const int row = dg * 48 + n;
      B[(size_t) ks * 18 * 256 + (row / 8) * 256 + (kk / 16) * 128 + (row % 8) * 16 + (kk % 16)] = (uint8_t) v;
    }
    Params P{}; P.cols_out = 192; P.n_ops = 0;
    for (int ks = 0; ks < 11; ks++)
    {
      const uint32_t a_hi = (uint32_t) ks * 2 * 128, a_lo = 112 * 128 + a_hi, b0 = (uint32_t) ks * 18 * 256;
      if (ks == 0)
      {
        P.ops[P.n_ops++] = Op{ a_hi, b0, 128, 768, 128, 256, 48, 1, 1, 0, 0 };                  // xh * hh -> cols 0..47, fresh
        P.ops[P.n_ops++] = Op{ a_lo, b0, 128, 768, 128, 256, 144, 0, 1, 48, 0 };                // xl * [hh|hm|hl] -> cols 48..191, fresh
        P.ops[P.n_ops++] = Op{ a_hi, b0 + 6 * 256, 128, 768, 128, 256, 96, 1, 1, 48, 1 };       // xh * [hm|hl] -> cols 48..143, accumulate
      }
      else
      {
        P.ops[P.n_ops++] = Op{ a_hi, b0, 128, 768, 128, 256, 144, 1, 1, 0, 1 };
        P.ops[P.n_ops++] = Op{ a_lo, b0, 128, 768, 128, 256, 144, 0, 1, 48, 1 };
      }
    }
    std::vector<int> D;
    if (run (A, B, P, D)) fails++;
    else
    {
      long bad = 0;
      for (int q = 0; q < 16; q++) for (int j = 0; j < J; j++) for (int n = 0; n < 48; n++)
      {
        // y[t] = sum_d hr[d] I[t-d] - hi[d] Q[t-d], t = 128 + 48 q + n in the 896-frame window
        const int t = 128 + 48 * q + n;
        long long want = 0;
        for (int d = 0; d < NT; d++) want += (long long) hr[d] * raw[(j * F + t - d) * 2] - (long long) hi[d] * raw[(j * F + t - d) * 2 + 1];
        const int r = q * 8 + j;
        const int *dv = &D[r * 192];
        const long long got = ((long long) dv[n] << 24) + ((long long) dv[48 + n] << 16) + ((long long) dv[96 + n] << 8) + dv[144 + n];
        if (got != want) { if (bad < 6) printf ("    q=%d j=%d n=%d want %lld got %lld (%d %d %d %d)\n", q, j, n, want, got, dv[n], dv[48 + n], dv[96 + n], dv[144 + n]); bad++; }
      }
      printf ("test 2 (FIR supertile, aliased row groups SBO = 768, 23 MMAs): %ld mismatches of %d\n", bad, 16 * 8 * 48);
      if (bad) fails++;
    }
  }
  printf (fails ? "PROBE FAILED\n" : "PROBE OK\n");
  return fails ? 1 : 0;
}
