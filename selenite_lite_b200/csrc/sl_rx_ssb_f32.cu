// RX-SSB-f32: the fused receive chain, one kernel, every sample crosses HBM once (4 B in, 4 B out).
//
//   int16 I/Q  --unpack-->  overlap-save FFT filter (512-pt, hop 384)  --Re-->  2-stage df2T biquad  -->  per-48-frame AGC  -->  int16 L=R
//   oracle stage per box (reference = /root/reference/Drivers/CMSIS/DSP/Source/...):
//     unpack   SupportFunctions/arm_q15_to_float.c:65            FFT/IFFT  TransformFunctions/arm_cfft_f32.c:562 (len 512)
//     mask     ComplexMathFunctions/arm_cmplx_mult_cmplx_f32.c:72  biquad   FilteringFunctions/arm_biquad_cascade_df2T_f32.c:142
//     AGC      BasicMathFunctions/arm_abs_f32.c:63, StatisticsFunctions/arm_max_f32.c:58, arm_scale_f32.c:77 (+ our gain law)
//     pack     SupportFunctions/arm_float_to_q15.c:64 (truncating)
//   and it sits where the firmware would call it: between pbuf and the ring store in DSP_In_Buff_Write (Core/Src/dsp_if.c:286-289).
//
// Work decomposition: a work item is (channel, tile of 4 hops = 1536 frames). Persistent CTAs pull
// items from an atomic queue ordered tile-major, so the only cross-tile dependency — 5 floats of biquad/AGC state per
// channel — is almost always already published when a CTA reaches the recurrence phase; it is handed over through
// global memory with a release/acquire flag per channel. The FFT phases never wait.
#include <cuda_runtime.h>
#include <cstdint>
#include "sl_internal.h"

namespace sl {

namespace {

constexpr int kN = 512;               // FFT length
constexpr int kHop = 384;             // new frames per FFT frame
constexpr int kOvl = kN - kHop;       // 128 carried frames
constexpr int kHopsPerTile = 4;
constexpr int kTile = kHop * kHopsPerTile;      // 1536 frames
constexpr int kThreads = 256;                   // 64 threads per FFT frame, radix-8
constexpr int kRunsPerTile = kTile / kRun;      // 32 lanes x 48 samples in the recurrence phase
constexpr int kRunPad = kRun + 1;               // 49: lane stride in shared memory, conflict-free
constexpr int kFramePad = kN + kN / 8;          // 576: phys(i) = i + (i >> 3)

struct KParams
{
  const uint32_t *in; uint32_t *out;            // one u32 = one I/Q (or L/R) frame
  float *audio_dbg; float *gain_dbg;
  const uint32_t *ovl_in; uint32_t *ovl_out;
  float *state; unsigned *flag; unsigned *queue;
  const float2 *masks; const uint8_t *mask_slot; const float2 *twiddle;
  unsigned flag_base;
  uint32_t channels, frames, tiles_per_channel;
  float agc_target, agc_decay, agc_floor, agc_gmax;
  BiquadScanTables tab;
};

__device__ __forceinline__ int phys (int i) { return i + (i >> 3); }

__device__ __forceinline__ float2 cmul (float2 a, float2 b) { return make_float2 (a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd (float2 a, float2 b) { return make_float2 (a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub (float2 a, float2 b) { return make_float2 (a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi (float2 a) { return make_float2 (a.y, -a.x); }      // a * (-i)

// 4-point forward DFT, natural order in and out
__device__ __forceinline__ void dft4 (float2 &u0, float2 &u1, float2 &u2, float2 &u3)
{
  float2 s0 = cadd (u0, u2), s1 = csub (u0, u2), s2 = cadd (u1, u3), s3 = mul_mi (csub (u1, u3));
  u0 = cadd (s0, s2); u2 = csub (s0, s2); u1 = cadd (s1, s3); u3 = csub (s1, s3);
}

// 8-point forward DFT, natural order in and out (decimation in frequency, outputs renamed at compile time)
__device__ __forceinline__ void dft8 (float2 *v)
{
  const float h = 0.70710678118654752f;
  float2 a0 = cadd (v[0], v[4]), a1 = cadd (v[1], v[5]), a2 = cadd (v[2], v[6]), a3 = cadd (v[3], v[7]);
  float2 b0 = csub (v[0], v[4]), t1 = csub (v[1], v[5]), t2 = csub (v[2], v[6]), t3 = csub (v[3], v[7]);
  float2 b1 = make_float2 ((t1.x + t1.y) * h, (t1.y - t1.x) * h);       // * W8^1 = (1 - i)/sqrt2
  float2 b2 = mul_mi (t2);                                              // * W8^2 = -i
  float2 b3 = make_float2 ((t3.y - t3.x) * h, -(t3.x + t3.y) * h);      // * W8^3 = (-1 - i)/sqrt2
  dft4 (a0, a1, a2, a3);
  dft4 (b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// One Stockham radix-8 pass for the 64 threads of a frame: v[] already holds x[j + 64 r] (twiddled by the caller's
// choice); scatter to idxD + r*Ns with idxD = (j / Ns) * Ns * 8 + (j % Ns).
template <int Ns>
__device__ __forceinline__ void twiddle8 (float2 *v, int j, const float2 *tw)
{
  if (Ns == 1) return;
  const int kk = j & (Ns - 1);
  const int step = kk * (kN / (Ns * 8));          // W_{8 Ns}^{kk} = W_512^{step}
#pragma unroll
  for (int r = 1; r < 8; r++) v[r] = cmul (v[r], tw[(r * step) & (kN - 1)]);
}

template <int Ns>
__device__ __forceinline__ int scatter_base (int j) { return (j / Ns) * Ns * 8 + (j & (Ns - 1)); }

__device__ __forceinline__ unsigned ld_acquire (const unsigned *p)
{
  unsigned v;
  asm volatile ("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release (unsigned *p, unsigned v)
{
  asm volatile ("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t pack_lr (float x)
{
  // arm_float_to_q15.c:147 : (q15_t) __SSAT((q31_t)(x * 32768.0f), 16) — cast truncates toward zero
  int v = __float2int_rz (x * 32768.0f);
  v = max (-32768, min (32767, v));
  const uint32_t u = (uint32_t) v & 0xFFFFu;
  return u | (u << 16);                            // stereo endpoint, L = R
}

__global__ void __launch_bounds__ (kThreads) rx_ssb_f32_kernel (const __grid_constant__ KParams P)
{
  __shared__ float2 sX[kHopsPerTile][kFramePad];
  __shared__ float sAudio[kRunsPerTile * kRunPad];
  __shared__ float sGain[kRunsPerTile];
  __shared__ float2 sTw[kN];
  __shared__ unsigned sItem;

  const int tid = threadIdx.x;
  const int fid = tid >> 6;                        // which of the 4 FFT frames of the tile this thread works on
  const int j = tid & 63;                          // its index inside the 64-thread FFT group

  for (int i = tid; i < kN; i += kThreads) sTw[i] = P.twiddle[i];

  const unsigned total_items = P.channels * P.tiles_per_channel;

  while (true)
  {
    __syncthreads ();                              // also covers sTw on the first trip and smem reuse afterwards
    if (tid == 0) sItem = atomicAdd (P.queue, 1u);
    __syncthreads ();
    const unsigned item = sItem;
    if (item >= total_items) break;
    const uint32_t tile = item / P.channels, c = item % P.channels;      // tile-major order
    const uint32_t t0 = tile * kTile;
    const int hops = min ((uint32_t) kHopsPerTile, (P.frames - t0) / kHop);
    const int nsamp = hops * kHop;
    const uint32_t *in_c = P.in + (size_t) c * P.frames;
    const float2 *mask = P.masks + (size_t) P.mask_slot[c] * kN;

    // ---- P1: load + unpack (arm_q15_to_float: x / 32768). Frame f covers stream samples [t0 + 384 f - 128, +512).
    for (int f = 0; f < hops; f++)
    {
#pragma unroll
      for (int i = tid; i < kN; i += kThreads)
      {
        const int64_t t = (int64_t) t0 + (int64_t) f * kHop - kOvl + i;
        const uint32_t iq = (t >= 0) ? __ldg (in_c + t) : __ldg (P.ovl_in + (size_t) c * kOvl + (t + kOvl));
        const float re = (float) (int16_t) (iq & 0xFFFFu) * (1.0f / 32768.0f);
        const float im = (float) (int16_t) (iq >> 16) * (1.0f / 32768.0f);
        sX[f][phys (i)] = make_float2 (re, im);
      }
    }
    // carry the raw tail for the next call (ping-pong buffer: this launch reads ovl_in, writes ovl_out)
    if (t0 + nsamp == P.frames)
      for (int i = tid; i < kOvl; i += kThreads) P.ovl_out[(size_t) c * kOvl + i] = __ldg (in_c + P.frames - kOvl + i);
    __syncthreads ();

    float2 v[8];
    const bool active = fid < hops;

    // ---- P2: forward FFT, three Stockham radix-8 passes (Ns = 1, 8, 64), in place with a barrier between gather and scatter
#define SL_PASS(NS, LOAD_EXPR, STORE_STMT)                                                  \
    if (active) {                                                                            \
      _Pragma ("unroll") for (int r = 0; r < 8; r++) { const int idx = j + 64 * r; v[r] = LOAD_EXPR; } \
      twiddle8<NS> (v, j, sTw);                                                              \
      dft8 (v);                                                                              \
    }                                                                                        \
    __syncthreads ();                                                                        \
    if (active) {                                                                            \
      const int base = scatter_base<NS> (j);                                                 \
      _Pragma ("unroll") for (int r = 0; r < 8; r++) { const int idx = base + r * NS; STORE_STMT; } \
    }                                                                                        \
    __syncthreads ();

    SL_PASS (1, sX[fid][phys (idx)], sX[fid][phys (idx)] = v[r])
    SL_PASS (8, sX[fid][phys (idx)], sX[fid][phys (idx)] = v[r])
    SL_PASS (64, sX[fid][phys (idx)], sX[fid][phys (idx)] = v[r])

    // ---- P3+P4: spectral mask (arm_cmplx_mult_cmplx_f32) and inverse FFT. arm_cfft_f32 inverse = conj in, forward
    // transform, conj and 1/N out (arm_cfft_f32.c:571-580, :604-614); the 1/N is folded into the mask (power of two,
    // exact) and the final conj disappears because only the real part is kept.
    {
      auto load_masked = [&] (int idx) {
        const float2 y = cmul (sX[fid][phys (idx)], __ldg (mask + idx));
        return make_float2 (y.x, -y.y);
      };
      SL_PASS (1, load_masked (idx), sX[fid][phys (idx)] = v[r])
    }
    SL_PASS (8, sX[fid][phys (idx)], sX[fid][phys (idx)] = v[r])
    // last pass: keep the last 384 outputs of the frame, real part only -> audio in lane-run layout
    SL_PASS (64, sX[fid][phys (idx)], if (idx >= kOvl) { const int n = fid * kHop + idx - kOvl; sAudio[(n / kRun) * kRunPad + (n % kRun)] = v[r].x; })
#undef SL_PASS

    // ---- P5: recurrences, warp 0. Lane k owns run k (48 samples = one AGC block).
    if (tid < 32)
    {
      const int lane = tid, nruns = nsamp / kRun;
      float *run = sAudio + lane * kRunPad;
      const float *cf = P.tab.coef;
      float d1a = 0.f, d2a = 0.f, d1b = 0.f, d2b = 0.f;
      if (lane < nruns)
      {
        // zero-state response of the cascade; per-sample recurrences as arm_biquad_cascade_df2T_f32.c:551-562
#pragma unroll 4
        for (int n = 0; n < kRun; n++)
        {
          const float x = run[n];
          const float y0 = cf[0] * x + d1a;
          d1a = (cf[1] * x + cf[3] * y0) + d2a;
          d2a = cf[2] * x + cf[4] * y0;
          const float y1 = cf[5] * y0 + d1b;
          d1b = (cf[6] * y0 + cf[8] * y1) + d2b;
          d2b = cf[7] * y0 + cf[9] * y1;
          run[n] = y1;
        }
      }
      // wait for the previous tile of this channel (tile-major queue order makes this a formality)
      if (lane == 0) { const unsigned want = P.flag_base + tile; while (ld_acquire (P.flag + c) != want) __nanosleep (64); }
      __syncwarp ();
      const float *stc = P.state + (size_t) c * 8;
      const float s0 = __ldcg (stc + 0), s1 = __ldcg (stc + 1), s2 = __ldcg (stc + 2), s3 = __ldcg (stc + 3);
      const float env0 = __ldcg (stc + 4);

      // end state of run k given all earlier runs: z_k = zs_k + M z_{k-1}; lane 0 folds the carried state in
      float z0 = d1a, z1 = d2a, z2 = d1b, z3 = d2b;
      if (lane == 0)
      {
        const float *M = P.tab.Mpow[0];
        z0 += M[0] * s0 + M[1] * s1 + M[2] * s2 + M[3] * s3;
        z1 += M[4] * s0 + M[5] * s1 + M[6] * s2 + M[7] * s3;
        z2 += M[8] * s0 + M[9] * s1 + M[10] * s2 + M[11] * s3;
        z3 += M[12] * s0 + M[13] * s1 + M[14] * s2 + M[15] * s3;
      }
#pragma unroll
      for (int k = 0; k < 5; k++)
      {
        const int d = 1 << k;
        const float *M = P.tab.Mpow[k];
        const float p0 = __shfl_up_sync (0xffffffffu, z0, d), p1 = __shfl_up_sync (0xffffffffu, z1, d);
        const float p2 = __shfl_up_sync (0xffffffffu, z2, d), p3 = __shfl_up_sync (0xffffffffu, z3, d);
        if (lane >= d)
        {
          z0 += M[0] * p0 + M[1] * p1 + M[2] * p2 + M[3] * p3;
          z1 += M[4] * p0 + M[5] * p1 + M[6] * p2 + M[7] * p3;
          z2 += M[8] * p0 + M[9] * p1 + M[10] * p2 + M[11] * p3;
          z3 += M[12] * p0 + M[13] * p1 + M[14] * p2 + M[15] * p3;
        }
      }
      // start state of this lane's run = end state of the previous run
      float b0 = __shfl_up_sync (0xffffffffu, z0, 1), b1 = __shfl_up_sync (0xffffffffu, z1, 1);
      float b2 = __shfl_up_sync (0xffffffffu, z2, 1), b3 = __shfl_up_sync (0xffffffffu, z3, 1);
      if (lane == 0) { b0 = s0; b1 = s1; b2 = s2; b3 = s3; }

      float peak = 0.f;
      if (lane < nruns)
      {
        float *adbg = P.audio_dbg ? P.audio_dbg + (size_t) c * P.frames + t0 + lane * kRun : nullptr;
#pragma unroll 4
        for (int n = 0; n < kRun; n++)
        {
          const float *C = P.tab.Cresp[n];
          const float y = run[n] + (C[0] * b0 + C[1] * b1 + C[2] * b2 + C[3] * b3);
          run[n] = y;
          peak = fmaxf (peak, fabsf (y));                                   // arm_abs_f32 + arm_max_f32
          if (adbg) adbg[n] = y;
        }
      }
      // AGC envelope: sequential over blocks, exactly the oracle's order
      float env = env0, my_env = 0.f;
      for (int i = 0; i < nruns; i++)
      {
        const float pi = __shfl_sync (0xffffffffu, peak, i);
        env = fmaxf (pi, env * P.agc_decay);
        if (lane == i) my_env = env;
      }
      if (lane < nruns)
      {
        const float g = fminf (__fdiv_rn (P.agc_target, fmaxf (my_env, P.agc_floor)), P.agc_gmax);
        sGain[lane] = g;
        if (P.gain_dbg) P.gain_dbg[(size_t) c * (P.frames / kRun) + t0 / kRun + lane] = g;
      }
      if (lane == nruns - 1)
      {
        float *stw = P.state + (size_t) c * 8;
        __stcg (stw + 0, z0); __stcg (stw + 1, z1); __stcg (stw + 2, z2); __stcg (stw + 3, z3); __stcg (stw + 4, env);
        __threadfence ();
        st_release (P.flag + c, P.flag_base + tile + 1u);
      }
    }
    __syncthreads ();

    // ---- P6: gain (arm_scale_f32), pack (arm_float_to_q15), coalesced store L = R
    uint32_t *out_c = P.out + (size_t) c * P.frames + t0;
    for (int n = tid; n < nsamp; n += kThreads)
    {
      const int r = n / kRun;
      out_c[n] = pack_lr (sAudio[r * kRunPad + (n - r * kRun)] * sGain[r]);
    }
  }
}

}  // namespace

uint32_t rx_ssb_f32_launches_per_call () { return 1u; }

int launch_rx_ssb_f32 (const RxF32Launch &L, int sm_count, void *stream_)
{
  cudaStream_t stream = (cudaStream_t) stream_;
  if (L.frames % kHop != 0 || L.frames == 0 || L.channels == 0) return (int) cudaErrorInvalidValue;
  KParams P;
  P.in = reinterpret_cast<const uint32_t *> (L.in); P.out = reinterpret_cast<uint32_t *> (L.out);
  P.audio_dbg = L.audio_dbg; P.gain_dbg = L.gain_dbg;
  P.ovl_in = reinterpret_cast<const uint32_t *> (L.ovl_in); P.ovl_out = reinterpret_cast<uint32_t *> (L.ovl_out);
  P.state = L.state; P.flag = L.flag; P.queue = L.queue;
  P.masks = reinterpret_cast<const float2 *> (L.masks); P.mask_slot = L.mask_slot;
  P.twiddle = reinterpret_cast<const float2 *> (L.twiddle);
  P.flag_base = L.flag_base; P.channels = L.channels; P.frames = L.frames;
  P.tiles_per_channel = (L.frames + kTile - 1) / kTile;
  P.agc_target = L.agc_target; P.agc_decay = L.agc_decay; P.agc_floor = L.agc_floor; P.agc_gmax = L.agc_gmax;
  P.tab = *L.tables;

  cudaError_t e = cudaMemsetAsync (L.queue, 0, sizeof (unsigned), stream);
  if (e != cudaSuccess) return (int) e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, rx_ssb_f32_kernel, kThreads, 0);
  if (e != cudaSuccess) return (int) e;
  if (per_sm < 1) per_sm = 1;
  const uint64_t items = (uint64_t) L.channels * P.tiles_per_channel;
  uint64_t grid = (uint64_t) sm_count * per_sm;
  if (grid > items) grid = items;
  rx_ssb_f32_kernel<<<(unsigned) grid, kThreads, 0, stream>>> (P);
  return (int) cudaGetLastError ();
}

uint32_t rx_ssb_f32_tiles (uint32_t frames) { return (frames + kTile - 1) / kTile; }

}  // namespace sl
