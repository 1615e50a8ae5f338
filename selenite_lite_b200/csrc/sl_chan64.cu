// CHAN-64-f32: wideband polyphase FFT channelizer feeding per-channel demodulator + AGC, one fused kernel
// (BASELINE config 4: 192 kHz wideband streams -> 64 narrowband channels of 3 kHz each -> demod + AGC).
//
//   int16 I/Q wideband --unpack--> 64-branch polyphase filter (8 taps/branch) --> 64-point FFT per hop --> Re or | |
//        --> per-1-ms AGC (3 narrowband samples) --> int16 L=R, written channel-major [stream][bin][hop]
//   oracle stage per box (reference = /root/reference/Drivers/CMSIS/DSP/Source/...):
//     unpack    SupportFunctions/arm_q15_to_float.c:65     branch FIR  FilteringFunctions/arm_fir_f32.c:553 (numTaps 8, on
//     FFT       TransformFunctions/arm_cfft_f32.c:562 (64)             the stride-64 commutated input, I and Q separately)
//     envelope  ComplexMathFunctions/arm_cmplx_mag_f32.c:72 (AM)  AGC  arm_abs_f32.c:63, arm_max_f32.c:58, arm_scale_f32.c:77
//     pack      SupportFunctions/arm_float_to_q15.c:64
//   The COMPOSITION is ours (SURVEY.md §0 / Appendix B "CHAN-64"); the arithmetic of each box is the reference's.
//
// Algorithmic bytes: 4 B in per wideband frame + 64 bins x 4 B out per 64 frames = 8 B per wideband complex sample.
// Work per wideband sample: 16 FMA (polyphase) + ~20 (FFT-64 as 8 x 8) + ~5 (AGC, pack): ~43 FP32 lane-ops, i.e. the
// FP32 pipe (35 T lane-ops/s) and the HBM roof (6.5 TB/s / 8 B) are about equal for this chain.
//
// Structure. Work item = (stream, tile of 48 hops = 3072 wideband frames). Items are dealt tile-major to a persistent,
// co-resident grid of warp-specialised CTAs (3 producer warps + 2 AGC warps, 3 CTAs per SM); nothing in the steady state
// is a CTA-wide barrier:
//   * one thread stages each tile's raw int16 frames (plus the 7 hops of FIR history that precede it) into a
//     double-buffered shared-memory tile with ONE bulk asynchronous copy (cp.async.bulk + mbarrier, SASS UBLKCP), two
//     items ahead of the compute.
//   * producer warps, 16 hops each, in two rounds of 8 hops. Polyphase: lane r owns branches r and r+32 with their 8+8
//     coefficients in registers and a sliding window of 8 complex inputs per branch, so each input sample is read from
//     shared memory once per round and each tap is one FMA in the oracle's accumulation order. The 8 x 64 branch
//     outputs go through a warp-private scratch; the 64-point FFTs run as 8 x 8 on groups of 8 lanes, two hops per lane
//     packed in FP32x2 registers (FADD2/FMUL2/FFMA2), with a skewed 8 x 8 exchange that is bank-conflict-free both ways.
//     The detector output goes to a double-buffered audio tile [bin][hop]; named barriers (full / empty per buffer)
//     connect the two sides as in the RX kernel.
//   * 64 AGC threads (one per bin) work one tile behind. The envelope recurrence env = max(peak, env*decay) is a
//     max-plus scan whose carry-in comes from the previous tile of the same stream, which another CTA processes
//     concurrently: tiles are chained with a DECOUPLED LOOK-BACK (publish the zero-start envelope at once, then the
//     inclusive one) so that no tile waits for a predecessor's full AGC pass, and because fl(max(a,b)*d) =
//     max(fl(a*d), fl(b*d)) the folded result is bit-identical to the oracle's sequential walk.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "sl_internal.h"

namespace sl {

namespace {
#include "sl_stress.cuh"

constexpr int kBins = 64;                        // branches = FFT length = decimation
constexpr int kTaps = 8;                         // taps per branch (prototype = 512 taps)
constexpr int kHistHops = kTaps - 1;             // hops of input history a tile needs
constexpr int kTileHops = 48;                    // hops per tile = 16 AGC blocks of 3
constexpr int kBlk = 3;                          // AGC block: 1 ms = 192 wideband frames = 3 narrowband samples
constexpr int kTileBlocks = kTileHops / kBlk;
#ifndef SL_CHAN_FFTWARPS
#define SL_CHAN_FFTWARPS 3                          /* producer warps per CTA: 3 (16 hops each, 3 CTAs per SM) or 6 (8 hops each, 2 CTAs per SM) */
#endif
constexpr int kFftWarps = SL_CHAN_FFTWARPS, kAgcWarps = 2;
static_assert (kFftWarps == 3 || kFftWarps == 6, "48 hops split into rounds of 8");
constexpr int kFftThreads = 32 * kFftWarps, kAgcThreads = 32 * kAgcWarps, kThreads = kFftThreads + kAgcThreads;
constexpr int kHopsPerWarp = kTileHops / kFftWarps;  // 16 = two rounds of 8
constexpr int kRowUnits = 72;                    // 64-bit units per hop-pair row of the FFT scratch (64 + 8: rows of the two
                                                 // hop pairs a half-warp touches land in different bank halves)
constexpr int kAudioStride = 52;                 // floats per bin row of the audio tile: 16-byte aligned rows, and 52 c mod 32
                                                 // runs through all multiples of 4 for c = 0..7 (conflict-free both ways)
constexpr int kLookBackMax = 8;                  // deepest fold before a tile insists on an inclusive predecessor

constexpr size_t kRawWords = (size_t) (kTileHops + kHistHops) * kBins;
constexpr size_t kRawBytes = 2 * kRawWords * 4;                                  // double-buffered
constexpr size_t kScratchBytes = (size_t) kFftWarps * 2 * 4 * kRowUnits * 8;
constexpr size_t kAudioWords = (size_t) kBins * kAudioStride;
constexpr size_t kAudioBytes = 2 * kAudioWords * 4;                              // double-buffered
#ifndef SL_CHAN_FIR_X2
#define SL_CHAN_FIR_X2 0                          /* 1: the polyphase FIR on FP32x2, one fma.rn.f32x2 per tap for (re, im): bit-identical, measured equal (288.6 vs 290.5 Gsamples/s) */
#endif
#ifndef SL_CHAN_EARLY_LOOK
#define SL_CHAN_EARLY_LOOK 0                      /* 1: look-back words fetched before the wait for the tile's audio: measured slower (283 vs 290 Gsamples/s) */
#endif
#ifndef SL_CHAN_LAST_REFILLS
#define SL_CHAN_LAST_REFILLS 0                    /* 1: no barrier among the producer warps, the last one done with a raw buffer refills it: measured slower (284 vs 290) */
#endif
#ifndef SL_CHAN_RD_UNROLL
#define SL_CHAN_RD_UNROLL 1
#endif
constexpr int kRdUnroll = SL_CHAN_RD_UNROLL;
#ifndef SL_CHAN_PACK
#define SL_CHAN_PACK 1                            /* who scales, packs and stores a tile. 0: the AGC thread of each channel row (every store instruction touches 32 lines);
                                                     1: the three producer warps after the FFTs of the next tile, gains handed over through shared memory, consecutive lanes
                                                     on consecutive 16-byte pieces of a row; 2: the AGC warps, same lane mapping as 1.
                                                     Measured (Gsamples/s): 0 -> 283, 1 -> 290 (and then the PRODUCER side is the limit: 305 with the AGC ablated, the AGC
                                                     side alone 566), 2 -> 237 (one more barrier and a shared-memory round trip on the AGC warps' critical path) */
#endif
#define SL_CHAN_PACK_FFT (SL_CHAN_PACK == 1)
constexpr int kGainStride = 20;                  // floats per row of the gain tile: 16 block gains + pad (80-byte rows: conflict-free 16-byte stores of 8 lanes)
constexpr size_t kGainBytes = SL_CHAN_PACK != 0 ? (size_t) kBins * kGainStride * 4 : 0;
constexpr size_t kSmemBytes = kRawBytes + kScratchBytes + kAudioBytes + 64 * 8 /* twiddles */ + 32 /* mbarriers */ + kGainBytes;

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk (float lo, float hi) { u64 r; asm ("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo_of (u64 a) { float x; [[maybe_unused]] float y; asm ("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x; }
__device__ __forceinline__ float hi_of (u64 a) { [[maybe_unused]] float x; float y; asm ("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return y; }
__device__ __forceinline__ u64 add2 (u64 a, u64 b) { u64 r; asm ("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2 (u64 a, u64 b) { u64 r; asm ("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2 (u64 a, u64 b) { u64 r; asm ("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2 (u64 a, u64 b, u64 c) { u64 r; asm ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// 8-point forward DFT of two columns at once (lo / hi halves), natural order in and out
__device__ __forceinline__ void dft8x2 (u64 *re, u64 *im)
{
  const u64 H = pk (0.70710678118654752f, 0.70710678118654752f), NH = pk (-0.70710678118654752f, -0.70710678118654752f);
  const u64 a0r = add2 (re[0], re[4]), a0i = add2 (im[0], im[4]), t0r = sub2 (re[0], re[4]), t0i = sub2 (im[0], im[4]);
  const u64 a1r = add2 (re[1], re[5]), a1i = add2 (im[1], im[5]), t1r = sub2 (re[1], re[5]), t1i = sub2 (im[1], im[5]);
  const u64 a2r = add2 (re[2], re[6]), a2i = add2 (im[2], im[6]), t2r = sub2 (re[2], re[6]), t2i = sub2 (im[2], im[6]);
  const u64 a3r = add2 (re[3], re[7]), a3i = add2 (im[3], im[7]), t3r = sub2 (re[3], re[7]), t3i = sub2 (im[3], im[7]);
  {
    const u64 s0r = add2 (a0r, a2r), s0i = add2 (a0i, a2i), s1r = sub2 (a0r, a2r), s1i = sub2 (a0i, a2i);
    const u64 s2r = add2 (a1r, a3r), s2i = add2 (a1i, a3i), dr = sub2 (a1r, a3r), di = sub2 (a1i, a3i);
    re[0] = add2 (s0r, s2r); im[0] = add2 (s0i, s2i);
    re[4] = sub2 (s0r, s2r); im[4] = sub2 (s0i, s2i);
    re[2] = add2 (s1r, di); im[2] = sub2 (s1i, dr);
    re[6] = sub2 (s1r, di); im[6] = add2 (s1i, dr);
  }
  {
    const u64 s0r = add2 (t0r, t2i), s0i = sub2 (t0i, t2r), s1r = sub2 (t0r, t2i), s1i = add2 (t0i, t2r);
    const u64 p = add2 (t1r, t1i), q = sub2 (t1i, t1r), u = sub2 (t3i, t3r), v = add2 (t3r, t3i);
    const u64 pu = add2 (p, u), qv = sub2 (q, v), pmu = sub2 (p, u), qpv = add2 (q, v);
    re[1] = fma2 (H, pu, s0r); im[1] = fma2 (H, qv, s0i);
    re[5] = fma2 (NH, pu, s0r); im[5] = fma2 (NH, qv, s0i);
    re[3] = fma2 (H, qpv, s1r); im[3] = fma2 (NH, pmu, s1i);
    re[7] = fma2 (NH, qpv, s1r); im[7] = fma2 (H, pmu, s1i);
  }
}

__device__ __forceinline__ void unpack_iq (uint32_t iq, float &i, float &q)
{
  const uint32_t u = iq ^ 0x80008000u;                      // exact int16 -> float through the mantissa of 2^23
  i = __uint_as_float (__byte_perm (u, 0x4B000000u, 0x7610)) - 8421376.0f;
  q = __uint_as_float (__byte_perm (u, 0x4B000000u, 0x7632)) - 8421376.0f;
}
// output stores that do not allocate in L1: the lanes of a warp store to 32 different lines, which would evict what the
// shared-memory / L1 data pipe is needed for (measured on the tensor-core RX kernel: 366 -> 408 Gsamples/s)
__device__ __forceinline__ void st_na (uint4 *p, uint4 v)
{
  asm volatile ("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_lr (float x_times_32768)
{
  short v;                                                  // arm_float_to_q15.c:147: truncate toward zero, saturate
  asm ("cvt.rzi.sat.s16.f32 %0, %1;" : "=h"(v) : "f"(x_times_32768));
  return __byte_perm ((uint32_t) (uint16_t) v, 0u, 0x1010);
}

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }
__device__ __forceinline__ void mbar_init (uint64_t *bar, unsigned count)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32 (bar)), "r"(count) : "memory");
  asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx (uint64_t *bar, unsigned bytes)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t *bar, unsigned parity)
{
  sl_jitter ();
  asm volatile (
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
      ::"r"(smem_u32 (bar)), "r"(parity) : "memory");
}
// one bulk asynchronous copy global -> shared, completion counted in bytes on the mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s (void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32 (dst)), "l"(src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ unsigned long long ld_look (const unsigned long long *p)
{
  unsigned long long v; asm volatile ("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_look (unsigned long long *p, unsigned status, float v)
{
  sl_jitter ();
  asm volatile ("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long) status << 32) | (unsigned long long) __float_as_uint (v)) : "memory");
}
// named barriers (like the RX kernel): 1,2 = audio buffer 0/1 full; 3,4 = audio buffer 0/1 empty (all 160 threads take
// part in each, one side syncs, the other arrives); 5 = the FFT warps among themselves; 6 = the AGC warps
__device__ __forceinline__ void bar_sync (int id, int n) { sl_jitter (); asm volatile ("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive (int id, int n) { sl_jitter (); asm volatile ("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct KParams
{
  const uint32_t *in; uint32_t *out;            // in [S][frames] u32 I/Q; out [S][64][hops] u32 L=R
  const uint32_t *hist_in; uint32_t *hist_out;  // [S][7*64] raw frames carried between calls (ping-pong)
  const float *env_in; float *env_out;          // [S][64] carried envelope (ping-pong: a late tile may finish before an
                                                // early one has read its carry-in)
  unsigned long long *look;                     // look-back words [S][tiles][64]: status (0 = none, 1 = zero-start aggregate, 2 = inclusive) << 32 | float bits.
                                                // One 8-byte word carries flag AND value, so a reader needs no fence and one round trip.
  const float *coef;                            // [64][8]: e_r[p] = h[64 p + 63 - r] / 32768
  const float2 *tw;                             // [8][8]: W_64^(k1*b)
  float *audio_dbg, *gain_dbg;                  // optional: [S][64][hops], [S][64][hops/3]
  uint32_t streams, hops, tiles, envelope;
  float target, decay, floor, gmax;
};

// n steps of the oracle's release walk from a zero peak history: x <- fl(x * decay), exactly as chains.inc.c does per block
__device__ __forceinline__ float decay_n (float x, float decay, int n)
{
  for (int i = 0; i < n; i++) x = x * decay;
  return x;
}

__global__ void __launch_bounds__ (kThreads, kFftWarps == 3 ? 3 : 2) chan64_f32_kernel (const __grid_constant__ KParams P)
{
  extern __shared__ __align__ (128) unsigned char smem[];
  uint32_t *sRaw = reinterpret_cast<uint32_t *> (smem);
  u64 *sScr = reinterpret_cast<u64 *> (smem + kRawBytes);
  float *sAudio = reinterpret_cast<float *> (smem + kRawBytes + kScratchBytes);
  float2 *sTw = reinterpret_cast<float2 *> (smem + kRawBytes + kScratchBytes + kAudioBytes);
  uint64_t *sBar = reinterpret_cast<uint64_t *> (smem + kRawBytes + kScratchBytes + kAudioBytes + 64 * 8);   // [2]: raw buffer full
  unsigned *sCnt = reinterpret_cast<unsigned *> (smem + kRawBytes + kScratchBytes + kAudioBytes + 64 * 8 + 16);   // [2]: producer warps done with a raw buffer
#if SL_CHAN_PACK != 0
  float *sGain = reinterpret_cast<float *> (smem + kRawBytes + kScratchBytes + kAudioBytes + 64 * 8 + 32);     // [64][kGainStride]: gains x 32768 of the tile being packed
#endif

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned total = P.streams * P.tiles;

  if (tid == 0) { mbar_init (sBar, 1); mbar_init (sBar + 1, 1); sCnt[0] = sCnt[1] = 0u; }
  if (tid < 64) sTw[tid] = P.tw[tid];
  __syncthreads ();

  if (warp < kFftWarps)
  {
    // =============================== producer side: staging, polyphase, FFT, detector ===============================
    // bulk copy of item `it` into raw buffer `buf` (thread 0 only): history rows + the tile's hops, contiguous in the
    // stream except for tile 0, whose history comes from the carried state
    auto issue_load = [&] (unsigned it, int buf) {
      const uint32_t tile = it / P.streams, s = it % P.streams;
      const uint32_t hops_here = min ((uint32_t) kTileHops, P.hops - tile * kTileHops);
      const uint32_t *src = P.in + (size_t) s * P.hops * kBins + (size_t) tile * kTileHops * kBins;
      uint32_t *dst = sRaw + (size_t) buf * kRawWords;
      mbar_expect_tx (sBar + buf, (hops_here + kHistHops) * kBins * 4);
      if (tile == 0)
      {
        bulk_g2s (dst, P.hist_in + (size_t) s * kHistHops * kBins, kHistHops * kBins * 4, sBar + buf);
        bulk_g2s (dst + kHistHops * kBins, src, hops_here * kBins * 4, sBar + buf);
      }
      else
        bulk_g2s (dst, src - kHistHops * kBins, (hops_here + kHistHops) * kBins * 4, sBar + buf);
    };
    if (tid == 0)
    {
      if (blockIdx.x < total) issue_load (blockIdx.x, 0);
      if (blockIdx.x + gridDim.x < total) issue_load (blockIdx.x + gridDim.x, 1);
    }
    // branch coefficients of this lane: branches r = lane and lane + 32
    float c0[kTaps], c1[kTaps];
#pragma unroll
    for (int p = 0; p < kTaps; p++) { c0[p] = P.coef[lane * kTaps + p]; c1[p] = P.coef[(lane + 32) * kTaps + p]; }
    u64 *scr_re = sScr + (size_t) warp * 2 * 4 * kRowUnits, *scr_im = scr_re + 4 * kRowUnits;
    const int g = lane >> 3, b = lane & 7;
    unsigned seq = 0;
#if SL_CHAN_PACK_FFT
    // scale (arm_scale_f32), pack (arm_float_to_q15), store of a finished tile by the 96 producer threads: the tile is 64 rows x 12
    // float4, thread t takes the float4 t, t + 96, ... — consecutive lanes store consecutive 16-byte pieces of a channel row (192
    // contiguous bytes per row and tile), where one thread per row made every store instruction touch 32 different lines. A
    // float4 q4 of a row holds samples 4 q4 .. 4 q4 + 3 of blocks (4 q4) / 3 and the next one: (g0 g0 g0 g1), (g0 g0 g1 g1) or
    // (g0 g1 g1 g1) by q4 mod 3.
    auto pack_item = [&] (unsigned pitem, int pb) {
      const uint32_t ptile = pitem / P.streams, ps = pitem % P.streams;
      const int nq4 = (int) (min ((uint32_t) kTileHops, P.hops - ptile * kTileHops) / 4);
      const float *aud = sAudio + (size_t) pb * kAudioWords;
      uint32_t *orow0 = P.out + (size_t) ps * kBins * P.hops + (size_t) ptile * kTileHops;
      bar_sync (3 + pb, kThreads);                                   // the AGC warps have written this tile's gains
#pragma unroll
      for (int i = 0; i < kBins * (kTileHops / 4) / kFftThreads; i++)
      {
        const int idx = tid + kFftThreads * i, row = idx / (kTileHops / 4), q4 = idx - row * (kTileHops / 4);
        if (q4 < nq4)
        {
          const float4 v = *reinterpret_cast<const float4 *> (aud + row * kAudioStride + 4 * q4);
          const int ga = (4 * q4) / 3, m = q4 - 3 * (q4 / 3);
          const float g0 = sGain[row * kGainStride + ga], g1 = sGain[row * kGainStride + ga + 1];
          st_na (reinterpret_cast<uint4 *> (orow0 + (size_t) row * P.hops) + q4,
                 make_uint4 (pack_lr (v.x * g0), pack_lr (v.y * (m == 2 ? g1 : g0)), pack_lr (v.z * (m == 0 ? g0 : g1)), pack_lr (v.w * g1)));
        }
      }
      bar_arrive (6 + pb, kThreads);                                 // the gain tile may be overwritten
    };
    static_assert (kBins * (kTileHops / 4) % kFftThreads == 0, "the pack loop covers the tile evenly");
#endif

    for (unsigned item = blockIdx.x; item < total; item += gridDim.x, seq++)
    {
      const int buf = seq & 1;
      const uint32_t tile = item / P.streams, s = item % P.streams;
      const uint32_t hops_here = min ((uint32_t) kTileHops, P.hops - tile * kTileHops);
      const uint32_t *raw = sRaw + (size_t) buf * kRawWords;
      float *audio = sAudio + (size_t) buf * kAudioWords;
      mbar_wait (sBar + buf, (seq >> 1) & 1);                        // raw tile has landed
#if !SL_CHAN_PACK_FFT
      if (seq >= 2) bar_sync (3 + buf, kThreads);                    // the AGC warps have drained this audio buffer
#endif                                                               // (else: these warps packed tile seq - 2 out of it themselves, after the AGC warps were done with it)

      const int h_base = warp * kHopsPerWarp;                        // first hop of this warp inside the tile
#ifdef SL_CHAN_ABLATE_FFT                                            // (profiling aid: the kernel without polyphase filter and FFTs)
      if (false)
#else
      if ((uint32_t) h_base < hops_here)
#endif
      {
#if SL_CHAN_FIR_X2
        // sliding windows: at hop m the oldest sample x_r[m-7] sits in slot (m+1)&7, the newest in slot m&7; a slot holds (re, im) as
        // one FP32x2 register pair, so a tap is ONE fma.rn.f32x2 for both rails (each lane IEEE-identical to the scalar fmaf)
        u64 w0[kTaps], w1[kTaps];
        const uint32_t *rawp = raw + (size_t) h_base * kBins + lane;  // row (hop_local + 7) holds hop hop_local
#pragma unroll
        for (int j = 0; j < kHistHops; j++)
        {
          float re, im;
          unpack_iq (rawp[j * kBins], re, im); w0[j + 1] = pk (re, im);
          unpack_iq (rawp[j * kBins + 32], re, im); w1[j + 1] = pk (re, im);
        }
#pragma unroll kRdUnroll
        for (int rd = 0; rd < kHopsPerWarp / 8; rd++)
        {
          const int h0 = h_base + 8 * rd;
          // ---- polyphase branch FIRs (arm_fir_f32.c: acc = sum_k state[n+k] * pCoeffs[k], oldest sample first,
          //      pCoeffs[k] = e_r[7-k]) for 8 hops; the two hops of a pair are stored together
#pragma unroll
          for (int m = 0; m < 8; m += 2)
          {
            u64 a0[2], a1[2];
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
              const int mm = m + e;
              float re, im;
              unpack_iq (rawp[(8 * rd + mm + kHistHops) * kBins], re, im); w0[mm & 7] = pk (re, im);
              unpack_iq (rawp[(8 * rd + mm + kHistHops) * kBins + 32], re, im); w1[mm & 7] = pk (re, im);
              u64 p0 = pk (0.f, 0.f), p1 = pk (0.f, 0.f);
#pragma unroll
              for (int k = 0; k < kTaps; k++)
              {
                const int slot = (mm + 1 + k) & 7;                   // x_r[m - 7 + k]
                p0 = fma2 (w0[slot], pk (c0[kTaps - 1 - k], c0[kTaps - 1 - k]), p0);
                p1 = fma2 (w1[slot], pk (c1[kTaps - 1 - k], c1[kTaps - 1 - k]), p1);
              }
              a0[e] = p0; a1[e] = p1;
            }
            // scratch row = hop pair, unit = branch; lo/hi = even/odd hop of the pair
            scr_re[(m >> 1) * kRowUnits + lane] = pk (lo_of (a0[0]), lo_of (a0[1])); scr_im[(m >> 1) * kRowUnits + lane] = pk (hi_of (a0[0]), hi_of (a0[1]));
            scr_re[(m >> 1) * kRowUnits + lane + 32] = pk (lo_of (a1[0]), lo_of (a1[1])); scr_im[(m >> 1) * kRowUnits + lane + 32] = pk (hi_of (a1[0]), hi_of (a1[1]));
          }
#else
        // sliding windows: at hop m the oldest sample x_r[m-7] sits in slot (m+1)&7, the newest in slot m&7
        float w0r[kTaps], w0i[kTaps], w1r[kTaps], w1i[kTaps];
        const uint32_t *rawp = raw + (size_t) h_base * kBins + lane;  // row (hop_local + 7) holds hop hop_local
#pragma unroll
        for (int j = 0; j < kHistHops; j++)
        {
          unpack_iq (rawp[j * kBins], w0r[j + 1], w0i[j + 1]);
          unpack_iq (rawp[j * kBins + 32], w1r[j + 1], w1i[j + 1]);
        }
#pragma unroll kRdUnroll
        for (int rd = 0; rd < kHopsPerWarp / 8; rd++)
        {
          const int h0 = h_base + 8 * rd;
          // ---- polyphase branch FIRs (arm_fir_f32.c: acc = sum_k state[n+k] * pCoeffs[k], oldest sample first,
          //      pCoeffs[k] = e_r[7-k]) for 8 hops; the two hops of a pair are stored together
#pragma unroll
          for (int m = 0; m < 8; m += 2)
          {
            float a0r[2], a0i[2], a1r[2], a1i[2];
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
              const int mm = m + e;
              unpack_iq (rawp[(8 * rd + mm + kHistHops) * kBins], w0r[mm & 7], w0i[mm & 7]);
              unpack_iq (rawp[(8 * rd + mm + kHistHops) * kBins + 32], w1r[mm & 7], w1i[mm & 7]);
              float p0r = 0.f, p0i = 0.f, p1r = 0.f, p1i = 0.f;
#pragma unroll
              for (int k = 0; k < kTaps; k++)
              {
                const int slot = (mm + 1 + k) & 7;                   // x_r[m - 7 + k]
                p0r = fmaf (w0r[slot], c0[kTaps - 1 - k], p0r); p0i = fmaf (w0i[slot], c0[kTaps - 1 - k], p0i);
                p1r = fmaf (w1r[slot], c1[kTaps - 1 - k], p1r); p1i = fmaf (w1i[slot], c1[kTaps - 1 - k], p1i);
              }
              a0r[e] = p0r; a0i[e] = p0i; a1r[e] = p1r; a1i[e] = p1i;
            }
            // scratch row = hop pair, unit = branch; lo/hi = even/odd hop of the pair
            scr_re[(m >> 1) * kRowUnits + lane] = pk (a0r[0], a0r[1]); scr_im[(m >> 1) * kRowUnits + lane] = pk (a0i[0], a0i[1]);
            scr_re[(m >> 1) * kRowUnits + lane + 32] = pk (a1r[0], a1r[1]); scr_im[(m >> 1) * kRowUnits + lane + 32] = pk (a1i[0], a1i[1]);
          }
#endif
          __syncwarp ();
          // ---- 64-point FFT (arm_cfft_f32 len 64, forward) of 8 hops: lane (g, b) = hop pair g, column b
          u64 xr[8], xi[8];
#pragma unroll
          for (int a = 0; a < 8; a++) { xr[a] = scr_re[g * kRowUnits + 8 * a + b]; xi[a] = scr_im[g * kRowUnits + 8 * a + b]; }
          dft8x2 (xr, xi);                                           // over a: index k1
#pragma unroll
          for (int k1 = 1; k1 < 8; k1++)
          {
            const float2 w = sTw[k1 * 8 + b];                        // W_64^(k1 b)
            const u64 wr = pk (w.x, w.x), wi = pk (w.y, w.y);
            const u64 tr = sub2 (mul2 (xr[k1], wr), mul2 (xi[k1], wi));
            xi[k1] = fma2 (xr[k1], wi, mul2 (xi[k1], wr)); xr[k1] = tr;
          }
          __syncwarp ();
          // skewed 8 x 8 exchange inside the 8-lane group: (k1, b) is stored at column (b + k1) & 7 of row k1
#pragma unroll
          for (int k1 = 0; k1 < 8; k1++)
          {
            const int u = g * kRowUnits + 8 * k1 + ((b + k1) & 7);
            scr_re[u] = xr[k1]; scr_im[u] = xi[k1];
          }
          __syncwarp ();
#pragma unroll
          for (int bb = 0; bb < 8; bb++)
          {
            const int u = g * kRowUnits + 8 * b + ((bb + b) & 7);    // this lane is k1 = b now
            xr[bb] = scr_re[u]; xi[bb] = scr_im[u];
          }
          dft8x2 (xr, xi);                                           // over b: index k2, bin = k1 + 8 k2
          __syncwarp ();
          // ---- detector: product (real part) or envelope (arm_cmplx_mag_f32); audio tile [bin][hop]
#pragma unroll
          for (int k2 = 0; k2 < 8; k2++)
          {
            float lo, hi;
            if (P.envelope)
            {
              const float rl = lo_of (xr[k2]), il = lo_of (xi[k2]), rh = hi_of (xr[k2]), ih = hi_of (xi[k2]);
              lo = __fsqrt_rn (__fadd_rn (__fmul_rn (rl, rl), __fmul_rn (il, il)));
              hi = __fsqrt_rn (__fadd_rn (__fmul_rn (rh, rh), __fmul_rn (ih, ih)));
            }
            else { lo = lo_of (xr[k2]); hi = hi_of (xr[k2]); }
            *reinterpret_cast<float2 *> (audio + (b + 8 * k2) * kAudioStride + h0 + 2 * g) = make_float2 (lo, hi);
          }
        }
      }
      // carried FIR history for the next call: the last 7 hops of the stream (rows hops_here .. hops_here + 6)
      if (tile == P.tiles - 1)
        for (int i = tid; i < kHistHops * kBins; i += kFftThreads) P.hist_out[(size_t) s * kHistHops * kBins + i] = raw[hops_here * kBins + i];
      const unsigned refill = item + 2 * gridDim.x;
#if SL_CHAN_LAST_REFILLS
      // the warp that is LAST done with raw[buf] (a shared-memory counter) issues the refill: nobody waits for the slowest warp
      __syncwarp ();
      if (lane == 0)
      {
        __threadfence_block ();
        if (atomicAdd (sCnt + buf, 1u) == (unsigned) kFftWarps - 1u)
        {
          sCnt[buf] = 0u;
          __threadfence_block ();
          if (refill < total) { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); issue_load (refill, buf); }
        }
      }
#else
      bar_sync (5, kFftThreads);                                     // every FFT warp is done with raw[buf]; audio[buf] complete
      if (tid == 0 && refill < total) { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); issue_load (refill, buf); }
#endif
      bar_arrive (1 + buf, kThreads);
#if SL_CHAN_PACK_FFT
      if (seq >= 1) pack_item (item - gridDim.x, buf ^ 1);           // the previous tile, while the AGC warps work on this one
#endif
    }
#if SL_CHAN_PACK_FFT
    if (seq >= 1) pack_item (blockIdx.x + (seq - 1) * gridDim.x, (int) ((seq - 1) & 1));
#endif
  }
  else
  {
    // ================================ consumer side: AGC, one thread per bin ================================
    const int k = tid - kFftThreads;
    const float decay = P.decay;
    unsigned seq = 0;
    for (unsigned item = blockIdx.x; item < total; item += gridDim.x, seq++)
    {
      const int buf = seq & 1;
      const uint32_t tile = item / P.streams, s = item % P.streams;
      const uint32_t hops_here = min ((uint32_t) kTileHops, P.hops - tile * kTileHops);
      const int nblk = hops_here / kBlk;
      const float4 *row = reinterpret_cast<const float4 *> (sAudio + (size_t) buf * kAudioWords + k * kAudioStride);
#if SL_CHAN_EARLY_LOOK
      // the look-back words of the predecessors do not depend on this tile: fetched before the wait for its audio, so the round trip
      // to L2 runs behind the barrier, the peak detection and the zero-start walk (re-fetched below if they do not suffice yet)
      const unsigned long long *lk = P.look + (size_t) s * P.tiles * kBins + k;
      unsigned long long wd[kLookBackMax];
#pragma unroll
      for (int d = 0; d < kLookBackMax; d++) { const int j = (int) tile - 1 - d; wd[d] = (j >= 0) ? ld_look (lk + (size_t) j * kBins) : (2ull << 32); }
#endif
      bar_sync (1 + buf, kThreads);                                  // audio tile complete
      // block peaks (arm_abs_f32 + arm_max_f32 over 3 samples); 4 blocks = 12 samples = 3 float4
      float pkv[kTileBlocks];
#ifdef SL_CHAN_ABLATE_AGC                                             // (profiling aid: no peak detection — WRONG results, timing only)
#pragma unroll
      for (int q = 0; q < kTileBlocks; q++) pkv[q] = 1.0f;
      if (false)
#endif
#pragma unroll
      for (int q = 0; q < kTileBlocks / 4; q++)
      {
        const float4 v0 = row[3 * q], v1 = row[3 * q + 1], v2 = row[3 * q + 2];
        pkv[4 * q + 0] = fmaxf (fmaxf (fabsf (v0.x), fabsf (v0.y)), fabsf (v0.z));
        pkv[4 * q + 1] = fmaxf (fmaxf (fabsf (v0.w), fabsf (v1.x)), fabsf (v1.y));
        pkv[4 * q + 2] = fmaxf (fmaxf (fabsf (v1.z), fabsf (v1.w)), fabsf (v2.x));
        pkv[4 * q + 3] = fmaxf (fmaxf (fabsf (v2.y), fabsf (v2.z)), fabsf (v2.w));
      }
      // zero-start envelope at the end of the tile: published at once so that successors never wait for this tile's
      // own carry-in
      float e0 = 0.f;
#pragma unroll
      for (int q = 0; q < kTileBlocks; q++) if (q < nblk) e0 = fmaxf (pkv[q], e0 * decay);
      const size_t slot = ((size_t) s * P.tiles + tile) * kBins + k;
      st_look (P.look + slot, 1u, e0);

      // carry-in by look-back over the predecessors of this stream: the words of the last kLookBackMax tiles of this bin are
      // fetched TOGETHER (one round trip whatever the depth) and re-fetched until the nearest inclusive one and every
      // aggregate after it are there. Virtual tile -1 is the call's carried state (always inclusive).
      float env;
#ifdef SL_CHAN_ABLATE_LOOKBACK                                       // (profiling aid: no carry-in from other CTAs — WRONG results, timing only)
      env = 0.f;
      if (false)
#endif
      {
#if !SL_CHAN_EARLY_LOOK
        const unsigned long long *lk = P.look + (size_t) s * P.tiles * kBins + k;
        unsigned long long wd[kLookBackMax];
#endif
        int dstar;
        unsigned spins = 0;
        for (bool first = true;; first = false)
        {
          if (!SL_CHAN_EARLY_LOOK || !first)
          {
#pragma unroll
            for (int d = 0; d < kLookBackMax; d++) { const int j = (int) tile - 1 - d; wd[d] = (j >= 0) ? ld_look (lk + (size_t) j * kBins) : (2ull << 32); }
          }
          dstar = -1;
          bool ok = true;
#pragma unroll
          for (int d = kLookBackMax - 1; d >= 0; d--) { if ((unsigned) (wd[d] >> 32) >= 2u) dstar = d; }      // nearest inclusive predecessor
          if (dstar < 0) ok = false;
#pragma unroll
          for (int d = 0; d < kLookBackMax; d++) if (d < dstar && (unsigned) (wd[d] >> 32) < 1u) ok = false;   // everything nearer has an aggregate
          if (ok) break;
          if (++spins > (1u << 23)) __trap ();                                // (seconds: a predecessor tile never published — an error, not a hung GPU)
          __nanosleep (40);
        }
        // fold forward: env_end(i) = max(E0(i), decay^16(env_end(i-1))), every predecessor tile is a full one
        env = 0.f;
#pragma unroll
        for (int d = kLookBackMax - 1; d >= 0; d--)
        {
          const int j = (int) tile - 1 - d;
          const float v = (j < 0) ? __ldcg (P.env_in + (size_t) s * kBins + k) : __uint_as_float ((unsigned) wd[d]);
          if (d == dstar) env = v;
          else if (d < dstar) env = fmaxf (v, decay_n (env, decay, kTileBlocks));
        }
      }
      // the real walk (oracle order), gains per block
      float gain[kTileBlocks];
#pragma unroll
      for (int q = 0; q < kTileBlocks; q++)
        if (q < nblk)
        {
          env = fmaxf (pkv[q], env * decay);
          // g = target / max(env, floor): reciprocal + one Newton step (<= 1 ulp from the oracle's IEEE division, far inside
          // the 1e-5 float tolerance; the envelope itself stays bit-exact)
          const float den = fmaxf (env, P.floor);
          float rc; asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));
          rc = fmaf (rc, fmaf (-den, rc, 1.0f), rc);
          const float gq = P.target * rc;
          gain[q] = fminf (fmaf (rc, fmaf (-den, gq, P.target), gq), P.gmax);
        }
      st_look (P.look + slot, 2u, env);
      if (tile == P.tiles - 1) __stcg (P.env_out + (size_t) s * kBins + k, env);

      const size_t orow = ((size_t) s * kBins + k) * P.hops + (size_t) tile * kTileHops;
#if SL_CHAN_PACK == 2
      // scale (arm_scale_f32), pack (arm_float_to_q15), store, by the 64 AGC threads together: the gains (x 32768: exact) go through
      // shared memory, then thread t takes the float4 t, t + 64, ... of the 64 x 12 float4 tile — consecutive lanes store consecutive
      // 16-byte pieces of a channel row (192 contiguous bytes per row and tile) instead of 32 different lines per instruction. A
      // float4 q4 of a row holds samples 4 q4 .. 4 q4 + 3 of blocks (4 q4) / 3 and the next one: (g0 g0 g0 g1), (g0 g0 g1 g1) or
      // (g0 g1 g1 g1) by q4 mod 3. (One gain tile: every AGC thread has passed the next tile's audio barrier before any writes it again.)
      {
        float4 *gr = reinterpret_cast<float4 *> (sGain + k * kGainStride);
#pragma unroll
        for (int q = 0; q < kTileBlocks / 4; q++)
          gr[q] = make_float4 (4 * q < nblk ? gain[4 * q] * 32768.0f : 0.f, 4 * q + 1 < nblk ? gain[4 * q + 1] * 32768.0f : 0.f,
                               4 * q + 2 < nblk ? gain[4 * q + 2] * 32768.0f : 0.f, 4 * q + 3 < nblk ? gain[4 * q + 3] * 32768.0f : 0.f);
      }
      if (P.gain_dbg || P.audio_dbg)
      {
#pragma unroll
        for (int q = 0; q < kTileBlocks / 4; q++)
          if (4 * q < nblk)
          {
            if (P.gain_dbg)
            {
              float *gd = P.gain_dbg + ((size_t) s * kBins + k) * (P.hops / kBlk) + (size_t) tile * kTileBlocks + 4 * q;
              gd[0] = gain[4 * q]; gd[1] = gain[4 * q + 1]; gd[2] = gain[4 * q + 2]; gd[3] = gain[4 * q + 3];
            }
            if (P.audio_dbg)
            {
              float4 *ad = reinterpret_cast<float4 *> (P.audio_dbg + orow) + 3 * q;
              ad[0] = row[3 * q]; ad[1] = row[3 * q + 1]; ad[2] = row[3 * q + 2];
            }
          }
      }
      bar_sync (6, kAgcThreads);                                      // every row's gains are in shared memory
      {
        const int nq4 = (int) hops_here / 4;
        const float *aud = sAudio + (size_t) buf * kAudioWords;
        uint32_t *orow0 = P.out + (size_t) s * kBins * P.hops + (size_t) tile * kTileHops;
#pragma unroll
        for (int i = 0; i < kBins * (kTileHops / 4) / kAgcThreads; i++)
        {
          const int idx = k + kAgcThreads * i, r = idx / (kTileHops / 4), q4 = idx - r * (kTileHops / 4);
          if (q4 < nq4)
          {
            const float4 v = *reinterpret_cast<const float4 *> (aud + r * kAudioStride + 4 * q4);
            const int ga = (4 * q4) / 3, m = q4 - 3 * (q4 / 3);
            const float g0 = sGain[r * kGainStride + ga], g1 = sGain[r * kGainStride + ga + 1];
            st_na (reinterpret_cast<uint4 *> (orow0 + (size_t) r * P.hops) + q4,
                   make_uint4 (pack_lr (v.x * g0), pack_lr (v.y * (m == 2 ? g1 : g0)), pack_lr (v.z * (m == 0 ? g0 : g1)), pack_lr (v.w * g1)));
          }
        }
      }
      bar_arrive (3 + buf, kThreads);                                // audio buffer drained
#elif SL_CHAN_PACK_FFT
      // hand the gains (x 32768: exact) to the producer warps, which scale, pack and store the tile; the gain tile has one buffer,
      // free once the previous tile is packed — long ago: that started when this tile's audio was complete
      if (seq >= 1) bar_sync (6 + (buf ^ 1), kThreads);
      {
        float4 *gr = reinterpret_cast<float4 *> (sGain + k * kGainStride);
#pragma unroll
        for (int q = 0; q < kTileBlocks / 4; q++)
          gr[q] = make_float4 (4 * q < nblk ? gain[4 * q] * 32768.0f : 0.f, 4 * q + 1 < nblk ? gain[4 * q + 1] * 32768.0f : 0.f,
                               4 * q + 2 < nblk ? gain[4 * q + 2] * 32768.0f : 0.f, 4 * q + 3 < nblk ? gain[4 * q + 3] * 32768.0f : 0.f);
      }
      if (P.gain_dbg || P.audio_dbg)
      {
#pragma unroll
        for (int q = 0; q < kTileBlocks / 4; q++)
          if (4 * q < nblk)
          {
            if (P.gain_dbg)
            {
              float *gd = P.gain_dbg + ((size_t) s * kBins + k) * (P.hops / kBlk) + (size_t) tile * kTileBlocks + 4 * q;
              gd[0] = gain[4 * q]; gd[1] = gain[4 * q + 1]; gd[2] = gain[4 * q + 2]; gd[3] = gain[4 * q + 3];
            }
            if (P.audio_dbg)
            {
              float4 *ad = reinterpret_cast<float4 *> (P.audio_dbg + orow) + 3 * q;
              ad[0] = row[3 * q]; ad[1] = row[3 * q + 1]; ad[2] = row[3 * q + 2];
            }
          }
      }
      bar_arrive (3 + buf, kThreads);                                // (after the debug taps: they read the audio row once more)
#else
      // scale (arm_scale_f32), pack (arm_float_to_q15), store channel-major: 12 samples = 3 float4 in, 3 uint4 out
      uint4 *dst = reinterpret_cast<uint4 *> (P.out + orow);
#pragma unroll
      for (int q = 0; q < kTileBlocks / 4; q++)
        if (4 * q < nblk)
        {
          const float4 v0 = row[3 * q], v1 = row[3 * q + 1], v2 = row[3 * q + 2];
          const float g0 = gain[4 * q] * 32768.0f, g1 = gain[4 * q + 1] * 32768.0f, g2 = gain[4 * q + 2] * 32768.0f, g3 = gain[4 * q + 3] * 32768.0f;
          st_na (dst + 3 * q, make_uint4 (pack_lr (v0.x * g0), pack_lr (v0.y * g0), pack_lr (v0.z * g0), pack_lr (v0.w * g1)));
          st_na (dst + 3 * q + 1, make_uint4 (pack_lr (v1.x * g1), pack_lr (v1.y * g1), pack_lr (v1.z * g2), pack_lr (v1.w * g2)));
          st_na (dst + 3 * q + 2, make_uint4 (pack_lr (v2.x * g2), pack_lr (v2.y * g3), pack_lr (v2.z * g3), pack_lr (v2.w * g3)));
          if (P.gain_dbg)
          {
            float *gd = P.gain_dbg + ((size_t) s * kBins + k) * (P.hops / kBlk) + (size_t) tile * kTileBlocks + 4 * q;
            gd[0] = gain[4 * q]; gd[1] = gain[4 * q + 1]; gd[2] = gain[4 * q + 2]; gd[3] = gain[4 * q + 3];
          }
          if (P.audio_dbg)
          {
            float4 *ad = reinterpret_cast<float4 *> (P.audio_dbg + orow) + 3 * q;
            ad[0] = v0; ad[1] = v1; ad[2] = v2;
          }
        }
      bar_arrive (3 + buf, kThreads);                                // audio buffer drained
#endif
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct Chan64State
{
  uint32_t streams = 0;
  slb_chan_params prm{};
  float *d_coef = nullptr; float2 *d_tw = nullptr;
  uint32_t *d_hist[2] = { nullptr, nullptr }; int parity = 0;
  float *d_env[2] = { nullptr, nullptr };
  float *dbg_audio = nullptr, *dbg_gain = nullptr;
  bool attr_set = false;
};

static const double kPi = 3.14159265358979323846;

static double bessel_i0 (double x)
{
  double s = 1.0, t = 1.0;
  for (int k = 1; k < 64; k++) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; if (t < 1e-18 * s) break; }
  return s;
}

// Frozen default (ours; nothing in the reference): 512-tap Kaiser(beta = 8) windowed-sinc prototype with its -6 dB point
// at half the bin spacing (fs/128), unit DC gain; AGC as RX-SSB-f32 at the same 1 ms block cadence.
int design_default_chan (uint32_t fs, slb_chan_params *p)
{
  if (!p || (fs != 48000u && fs != 96000u && fs != 192000u)) return SLB_ERR_ARG;
  std::memset (p, 0, sizeof *p);
  p->bins = kBins; p->taps_per_branch = kTaps; p->agc_block = kBlk; p->envelope = 0;
  p->agc_target = 0.25f; p->agc_decay = (float) std::exp (-1.0 / 300.0); p->agc_floor = 1.0e-4f; p->agc_gmax = 100.0f;
  const int n = kBins * kTaps;
  const double fc = 0.5 / kBins, mid = (n - 1) / 2.0, beta = 8.0;
  std::vector<double> h (n);
  double dc = 0.0;
  for (int i = 0; i < n; i++)
  {
    const double m = i - mid, x = 2.0 * fc * m;
    const double sinc = (std::fabs (x) < 1e-12) ? 1.0 : std::sin (kPi * x) / (kPi * x);
    const double r = m / mid;
    h[i] = 2.0 * fc * sinc * bessel_i0 (beta * std::sqrt (std::fmax (0.0, 1.0 - r * r))) / bessel_i0 (beta);
    dc += h[i];
  }
  for (int i = 0; i < n; i++) p->proto[i] = (float) (h[i] / dc);
  return SLB_OK;
}

static int chan_upload (slb_ctx *ctx, Chan64State *st)
{
  // e_r[p] = h[64 p + 63 - r]; 1/32768 of arm_q15_to_float is a power of two and is folded in (changes no bit)
  std::vector<float> coef ((size_t) kBins * kTaps);
  for (int r = 0; r < kBins; r++) for (int p = 0; p < kTaps; p++) coef[(size_t) r * kTaps + p] = st->prm.proto[kBins * p + (kBins - 1 - r)] * (1.0f / 32768.0f);
  std::vector<float2> tw (64);
  for (int k1 = 0; k1 < 8; k1++) for (int b = 0; b < 8; b++)
  {
    const double a = -2.0 * kPi * (double) (k1 * b) / 64.0;
    tw[k1 * 8 + b] = make_float2 ((float) std::cos (a), (float) std::sin (a));
  }
  if (cudaMemcpy (st->d_coef, coef.data (), coef.size () * sizeof (float), cudaMemcpyHostToDevice) != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, "chan64: coefficient upload failed");
  if (cudaMemcpy (st->d_tw, tw.data (), tw.size () * sizeof (float2), cudaMemcpyHostToDevice) != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, "chan64: twiddle upload failed");
  return SLB_OK;
}

void chan64_destroy (Chan64State *st)
{
  if (!st) return;
  cudaFree (st->d_coef); cudaFree (st->d_tw); cudaFree (st->d_hist[0]); cudaFree (st->d_hist[1]); cudaFree (st->d_env[0]); cudaFree (st->d_env[1]);
  delete st;
}

int chan64_reset (slb_ctx *ctx, Chan64State *st)
{
  const size_t hb = (size_t) st->streams * kHistHops * kBins * 4;
  if (cudaMemset (st->d_hist[0], 0, hb) != cudaSuccess || cudaMemset (st->d_hist[1], 0, hb) != cudaSuccess ||
      cudaMemset (st->d_env[0], 0, (size_t) st->streams * kBins * 4) != cudaSuccess || cudaMemset (st->d_env[1], 0, (size_t) st->streams * kBins * 4) != cudaSuccess)
    return ctx_fail (ctx, SLB_ERR_CUDA, "chan64: state reset failed");
  st->parity = 0;
  return SLB_OK;
}

int chan64_create (slb_ctx *ctx, uint32_t streams, uint32_t fs, Chan64State **out)
{
  Chan64State *st = new Chan64State ();
  st->streams = streams;
  design_default_chan (fs, &st->prm);
  const size_t hb = (size_t) streams * kHistHops * kBins * 4;
  if (cudaMalloc (&st->d_coef, (size_t) kBins * kTaps * 4) != cudaSuccess || cudaMalloc (&st->d_tw, 64 * sizeof (float2)) != cudaSuccess ||
      cudaMalloc (&st->d_hist[0], hb) != cudaSuccess || cudaMalloc (&st->d_hist[1], hb) != cudaSuccess ||
      cudaMalloc (&st->d_env[0], (size_t) streams * kBins * 4) != cudaSuccess || cudaMalloc (&st->d_env[1], (size_t) streams * kBins * 4) != cudaSuccess)
  { chan64_destroy (st); return ctx_fail (ctx, SLB_ERR_CUDA, "chan64: allocation failed"); }
  int rc = chan_upload (ctx, st);
  if (rc == SLB_OK) rc = chan64_reset (ctx, st);
  if (rc != SLB_OK) { chan64_destroy (st); return rc; }
  *out = st;
  return SLB_OK;
}

int chan64_set_params (slb_ctx *ctx, Chan64State *st, const slb_chan_params *p)
{
  if (p->bins != (uint32_t) kBins || p->taps_per_branch != (uint32_t) kTaps || p->agc_block != (uint32_t) kBlk)
    return ctx_fail (ctx, SLB_ERR_UNSUPPORTED, "this build has a kernel for bins=64, taps_per_branch=8, agc_block=3 only");
  if (!(p->agc_decay > 0.f && p->agc_decay <= 1.f) || !(p->agc_floor > 0.f) || !(p->agc_gmax > 0.f) || !(p->agc_target > 0.f))
    return ctx_fail (ctx, SLB_ERR_ARG, "AGC constants out of range");
  st->prm = *p;
  return chan_upload (ctx, st);
}
const slb_chan_params *chan64_params (const Chan64State *st) { return &st->prm; }
void chan64_set_debug (Chan64State *st, float *audio, float *gain) { st->dbg_audio = audio; st->dbg_gain = gain; }
size_t chan64_state_bytes (const Chan64State *st) { return (size_t) st->streams * (kHistHops * kBins * 4 + kBins * 4); }
int chan64_state_save (Chan64State *st, char *dst)
{
  const size_t hb = (size_t) st->streams * kHistHops * kBins * 4;
  if (cudaMemcpy (dst, st->d_hist[st->parity], hb, cudaMemcpyDeviceToHost) != cudaSuccess) return SLB_ERR_CUDA;
  if (cudaMemcpy (dst + hb, st->d_env[st->parity], (size_t) st->streams * kBins * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return SLB_ERR_CUDA;
  return SLB_OK;
}
int chan64_state_load (Chan64State *st, const char *src)
{
  const size_t hb = (size_t) st->streams * kHistHops * kBins * 4;
  st->parity = 0;
  if (cudaMemcpy (st->d_hist[0], src, hb, cudaMemcpyHostToDevice) != cudaSuccess) return SLB_ERR_CUDA;
  if (cudaMemcpy (st->d_env[0], src + hb, (size_t) st->streams * kBins * 4, cudaMemcpyHostToDevice) != cudaSuccess) return SLB_ERR_CUDA;
  return SLB_OK;
}

// one launch over streams [s0, s0 + ns) of the context (the host bulk path cuts the batch into stream groups)
int chan64_launch (slb_ctx *ctx, Chan64State *st, const int16_t *d_in, int16_t *d_out, uint32_t s0, uint32_t ns, uint32_t frames,
                   int sm_count, void *stream_, bool with_debug)
{
  cudaStream_t stream = (cudaStream_t) stream_;
  if (frames == 0 || frames % (kBins * 12) != 0) return ctx_fail (ctx, SLB_ERR_ARG, "frames must be a multiple of 768 (4 firmware blocks at 192 kHz)");
  KParams P{};
  P.streams = ns; P.hops = frames / kBins; P.tiles = (P.hops + kTileHops - 1) / kTileHops;
  // look-back scratch of this launch (per stream group: concurrent launches of one context use disjoint regions)
  const size_t per_stream = (size_t) P.tiles * kBins * 8;
  char *scr = static_cast<char *> (ctx_scratch (ctx, (size_t) st->streams * per_stream + 256));
  if (!scr) return SLB_ERR_CUDA;
  scr += (size_t) s0 * per_stream;
  P.look = reinterpret_cast<unsigned long long *> (scr);
  if (cudaMemsetAsync (P.look, 0, (size_t) ns * P.tiles * kBins * 8, stream) != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, "chan64: look-back reset failed");
  P.in = reinterpret_cast<const uint32_t *> (d_in); P.out = reinterpret_cast<uint32_t *> (d_out);
  P.hist_in = st->d_hist[st->parity] + (size_t) s0 * kHistHops * kBins; P.hist_out = st->d_hist[st->parity ^ 1] + (size_t) s0 * kHistHops * kBins;
  P.env_in = st->d_env[st->parity] + (size_t) s0 * kBins; P.env_out = st->d_env[st->parity ^ 1] + (size_t) s0 * kBins;
  P.coef = st->d_coef; P.tw = st->d_tw;
  P.audio_dbg = with_debug ? st->dbg_audio : nullptr; P.gain_dbg = with_debug ? st->dbg_gain : nullptr;
  P.envelope = st->prm.envelope;
  P.target = st->prm.agc_target; P.decay = st->prm.agc_decay; P.floor = st->prm.agc_floor; P.gmax = st->prm.agc_gmax;

  if (!st->attr_set)
  {
    cudaError_t e = cudaFuncSetAttribute (chan64_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBytes);
    if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
    st->attr_set = true;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, chan64_f32_kernel, kThreads, kSmemBytes);
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  if (per_sm < 1) per_sm = 1;
  const uint64_t items = (uint64_t) ns * P.tiles;
  uint64_t grid = (uint64_t) sm_count * per_sm;          // co-resident: the look-back may spin on a predecessor
  if (grid > items) grid = items;
  // cooperative launch (gang-scheduled grid): the look-back polls predecessors, see launch_rx_ssb_f32
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3 ((unsigned) grid); lc.blockDim = dim3 (kThreads); lc.dynamicSmemBytes = kSmemBytes; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  lc.attrs = at; lc.numAttrs = 1;
  e = cudaLaunchKernelEx (&lc, chan64_f32_kernel, P);
  if (e == cudaSuccess) e = cudaGetLastError ();
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  ctx_count_launch (ctx, 1);
  return SLB_OK;
}
void chan64_advance (Chan64State *st) { st->parity ^= 1; }

}  // namespace sl
