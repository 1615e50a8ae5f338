// Batched stage library: every CMSIS-DSP routine SURVEY.md §8(a) lists as a stage of the path, as a stand-alone
// kernel over [channels][n] arrays resident in HBM, with the per-channel state the CMSIS instance would hold carried
// in a caller-owned device buffer. Reference = /root/reference/Drivers/CMSIS/DSP/Source/<group>/<name>.c (cited per
// kernel). These are the unfused building blocks (HBM-bound element-wise / FIR kernels, channel-parallel
// recurrences); the fused chains (sl_rx_ssb_f32.cu, sl_chains.cu) are what the hot path runs.
//
// Arithmetic contract: integer routines are bit-exact by construction. Float routines use explicit __fmul_rn /
// __fadd_rn in the reference's association order, so nvcc cannot contract them into FMAs and the sequential ones
// (FIR, biquads, statistics, complex products) are bit-exact against the host build (gcc -ffp-contract=off) too.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include <cstdio>
#include "sl_internal.h"

namespace sl {

// ---- saturation helpers: cmsis_gcc.h:1299 __SSAT takes an int32_t (64-bit accumulators truncate first) ----
__device__ __forceinline__ int ssat16 (int v) { return max (-32768, min (32767, v)); }

// =====================================================================================================================
// element-wise kernels: one grid-stride template, the op is a functor
// =====================================================================================================================
template <typename Op> __global__ void ew_kernel (Op op, size_t n)
{
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) op (i);
}
template <typename Op> static int launch_ew (Op op, size_t n, cudaStream_t st)
{
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256; if (blocks > 148u * 16u) blocks = 148u * 16u;
  ew_kernel<<<(unsigned) blocks, 256, 0, st>>> (op, n);
  return (int) cudaGetLastError ();
}

struct OpQ15ToFloat { const int16_t *s; float *d; __device__ void operator() (size_t i) const { d[i] = (float) s[i] / 32768.0f; } };           // arm_q15_to_float.c:87
struct OpFloatToQ15 { const float *s; int16_t *d; __device__ void operator() (size_t i) const { d[i] = (int16_t) ssat16 (__float2int_rz (__fmul_rn (s[i], 32768.0f))); } };   // arm_float_to_q15.c:147
struct OpScaleF32 { const float *s; float k; float *d; __device__ void operator() (size_t i) const { d[i] = __fmul_rn (s[i], k); } };           // arm_scale_f32.c:77
struct OpMultF32 { const float *a, *b; float *d; __device__ void operator() (size_t i) const { d[i] = __fmul_rn (a[i], b[i]); } };
struct OpAddF32 { const float *a, *b; float *d; __device__ void operator() (size_t i) const { d[i] = __fadd_rn (a[i], b[i]); } };
struct OpSubF32 { const float *a, *b; float *d; __device__ void operator() (size_t i) const { d[i] = __fsub_rn (a[i], b[i]); } };
struct OpAbsF32 { const float *a; float *d; __device__ void operator() (size_t i) const { d[i] = fabsf (a[i]); } };                              // arm_abs_f32.c:63
struct OpScaleQ15 { const int16_t *s; int k, ksh; int16_t *d; __device__ void operator() (size_t i) const { d[i] = (int16_t) ssat16 (((int) s[i] * k) >> ksh); } };  // arm_scale_q15.c:138
struct OpAddQ15 { const int16_t *a, *b; int16_t *d; __device__ void operator() (size_t i) const { d[i] = (int16_t) ssat16 ((int) a[i] + b[i]); } };  // arm_add_q15.c:115
struct OpSubQ15 { const int16_t *a, *b; int16_t *d; __device__ void operator() (size_t i) const { d[i] = (int16_t) ssat16 ((int) a[i] - b[i]); } };  // arm_sub_q15.c:115
struct OpAbsQ15 { const int16_t *a; int16_t *d; __device__ void operator() (size_t i) const { int v = a[i]; d[i] = (int16_t) (v > 0 ? v : (v == -32768 ? 32767 : -v)); } };
struct OpShiftQ15 { const int16_t *a; int sh; int16_t *d; __device__ void operator() (size_t i) const { int v = a[i]; d[i] = (int16_t) (sh >= 0 ? ssat16 (v << sh) : (v >> (-sh))); } };
struct OpCmulF32 { const float2 *a, *b; float2 *d; __device__ void operator() (size_t i) const {                                                   // arm_cmplx_mult_cmplx_f32.c:72
  float2 x = a[i], y = b[i]; d[i] = make_float2 (__fsub_rn (__fmul_rn (x.x, y.x), __fmul_rn (x.y, y.y)), __fadd_rn (__fmul_rn (x.x, y.y), __fmul_rn (x.y, y.x))); } };
struct OpCmulRealF32 { const float2 *a; const float *r; float2 *d; __device__ void operator() (size_t i) const { float2 x = a[i]; float k = r[i]; d[i] = make_float2 (__fmul_rn (x.x, k), __fmul_rn (x.y, k)); } };
struct OpConjF32 { const float2 *a; float2 *d; __device__ void operator() (size_t i) const { float2 x = a[i]; d[i] = make_float2 (x.x, -x.y); } };
struct OpMagSqF32 { const float2 *a; float *d; __device__ void operator() (size_t i) const { float2 x = a[i]; d[i] = __fadd_rn (__fmul_rn (x.x, x.x), __fmul_rn (x.y, x.y)); } };
struct OpMagF32 { const float2 *a; float *d; __device__ void operator() (size_t i) const {                                                          // arm_cmplx_mag_f32.c:72 via arm_sqrt_f32 (arm_math.h:5726)
  float2 x = a[i]; float v = __fadd_rn (__fmul_rn (x.x, x.x), __fmul_rn (x.y, x.y)); d[i] = v >= 0.0f ? __fsqrt_rn (v) : 0.0f; } };

// q15 square root with the exact integer results of arm_sqrt_q15.c:50-140: normalise by an even shift so the mantissa m lies in
// [0.25, 1), seed r ~ 1/sqrt(m) from the float bit pattern, refine r three times in q15 (every intermediate truncated to 16 bits
// where the reference truncates), sqrt = m * r, then undo half the normalisation shift.
__device__ __forceinline__ int16_t trunc16 (int v) { return (int16_t) v; }
__device__ __forceinline__ int16_t sqrt_q15 (int16_t x)
{
  if (x <= 0) return 0;
  const int lead = __clz ((int) x) - 17;                 // redundant sign bits of the 16-bit value
  const int even_shift = lead & ~1;                      // the reference shifts by lead or lead - 1, whichever is even
  const int16_t m = trunc16 ((int) x << even_shift);
  const int16_t m_half = trunc16 (m >> 1);
  const float seed = __int_as_float (0x5f3759df - (__float_as_int (__fmul_rn ((float) m, 1.0f / 32768.0f)) >> 1));
  int16_t r = trunc16 (__float2int_rz (__fmul_rn (seed, 16384.0f)));
#pragma unroll
  for (int step = 0; step < 3; step++)
  {
    const int16_t r2 = trunc16 (((int) r * r) >> 15);
    const int16_t t = trunc16 (((int) r2 * m_half) >> 15);
    r = trunc16 ((int) trunc16 (((int) r * (0x3000 - t)) >> 15) << 2);
  }
  const int16_t root = trunc16 ((int) trunc16 (((int) m * r) >> 15) << 1);
  return trunc16 (root >> (even_shift >> 1));
}
struct OpMagQ15 { const short2 *a; int16_t *d; __device__ void operator() (size_t i) const {                                                         // arm_cmplx_mag_q15.c:120-130, result in 2.14
  short2 x = a[i]; long long acc = (long long) ((int) x.x * x.x) + (long long) ((int) x.y * x.y); d[i] = sqrt_q15 ((int16_t) (acc >> 17)); } };

// arm_sin_f32.c:72-119 / arm_cos_f32.c : 512-entry table (sin(2 pi n / 512) written with 8 decimals) + linear interpolation
struct OpSinCosF32 { const float *x; float *d; const float *tab; int is_cos;
  __device__ void operator() (size_t i) const
  {
    const float v = x[i]; float in; int n;
    if (!is_cos)
    {
      if (v < 0.0f && v >= -1.9e-7f) { d[i] = v; return; }
      in = __fmul_rn (v, 0.159154943092f); n = (int) in; if (v < 0.0f) n--;
    }
    else { in = __fadd_rn (__fmul_rn (v, 0.159154943092f), 0.25f); n = (int) in; if (in < 0.0f) n--; }
    in = __fsub_rn (in, (float) n);
    const float findex = __fmul_rn (512.0f, in);
    const unsigned idx = ((unsigned) (uint16_t) (int) findex) & 0x1ffu;
    const float fract = __fsub_rn (findex, (float) idx);
    d[i] = __fadd_rn (__fmul_rn (__fsub_rn (1.0f, fract), tab[idx]), __fmul_rn (fract, tab[idx + 1]));
  } };

// =====================================================================================================================
// FIR family — time-parallel: one thread per output, taps accumulated oldest-first exactly like the reference loop.
// hist = the first ntaps-1 entries of the CMSIS state buffer (previous samples, oldest first), [channels][hist_len].
// =====================================================================================================================
template <typename T> __device__ __forceinline__ T fir_sample_at (const T *hist, const T *src, int hist_len, long long pos)
{ return pos < hist_len ? hist[pos] : src[pos - hist_len]; }     // pos indexes the virtual sequence [hist | src]

// mode 0: arm_fir_f32.c:553  1: arm_fir_q15.c:591-642 (q63 acc, >>15, sat16)  2: arm_fir_fast_q15.c:60 (wrapping 32-bit acc)
// 3: arm_fir_q31.c:60 (q63 acc >> 31). Decimation M (arm_fir_decimate_*.c) and interpolation L (arm_fir_interpolate_*.c:508-526).
template <typename T, int MODE>
__global__ void fir_kernel (const T *__restrict__ coeffs, int ntaps, int M, int L, const T *__restrict__ hist_in, T *__restrict__ hist_out,
                            const T *__restrict__ src, T *__restrict__ dst, uint32_t n)
{
  const uint32_t c = blockIdx.y;
  const int P = ntaps / L, hist_len = P - 1;                         // L = 1: P = ntaps
  const T *h = hist_in + (size_t) c * hist_len, *s = src + (size_t) c * n;
  const uint32_t n_out = (n / M) * L;
  T *d = dst + (size_t) c * n_out;
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += gridDim.x * blockDim.x)
  {
    const uint32_t i = (o / L) * M, j = o % L;                       // input position of the window start, polyphase branch
    const T *cf = coeffs + (L - 1 - j);
    if (MODE == 0)
    {
      float acc = 0.0f;
      for (int k = 0; k < P; k++) acc = __fadd_rn (acc, __fmul_rn ((float) fir_sample_at (h, s, hist_len, (long long) i + k), (float) cf[k * L]));
      d[o] = (T) acc;
    }
    else if (MODE == 1)
    {
      long long acc = 0;
      for (int k = 0; k < P; k++) acc += (int) fir_sample_at (h, s, hist_len, (long long) i + k) * (int) cf[k * L];
      d[o] = (T) ssat16 ((int) (acc >> 15));
    }
    else if (MODE == 2)
    {
      unsigned acc = 0;
      for (int k = 0; k < P; k++) acc += (unsigned) ((int) fir_sample_at (h, s, hist_len, (long long) i + k) * (int) cf[k * L]);
      d[o] = (T) ssat16 ((int) acc >> 15);
    }
    else
    {
      long long acc = 0;
      for (int k = 0; k < P; k++) acc += (long long) fir_sample_at (h, s, hist_len, (long long) i + k) * (long long) cf[k * L];
      d[o] = (T) (int) (acc >> 31);
    }
  }
  // new history = last hist_len samples of [hist | src]; written to a second buffer (other CTAs still read hist_in)
  if (blockIdx.x == 0)
    for (int k = threadIdx.x; k < hist_len; k += blockDim.x)
      hist_out[(size_t) c * hist_len + k] = fir_sample_at (h, s, hist_len, (long long) n + k);
}

// =====================================================================================================================
// normalised LMS (arm_lms_norm_f32.c:161) — the adaptive stage behind noise reduction (take `out`: the predictable part of the
// audio) and the auto-notch (take `err`: the audio minus its predictable, tonal part) when src is a delayed copy of ref.
// The coefficients change with every sample, so a channel is strictly sequential: one thread per channel, a warp = 32 channels.
// Window and coefficients live in shared memory ([tap][lane]: conflict-free); the window is kept twice, ntaps apart, so that
// the oldest-first walk of the reference (:359-366) is a contiguous run without index wrapping. Samples travel through
// 32 x 32 tiles (coalesced global accesses, padded rows). Every operation is rounded separately in the reference's order
// (the oracle build has no FMA contraction), so output, error, coefficients and state agree bit for bit.
// =====================================================================================================================
constexpr int kLmsTile = 32;
__global__ void __launch_bounds__ (32) lms_norm_f32_kernel (float *__restrict__ coeffs, int ntaps, float mu, float *__restrict__ state,
                                                            const float *__restrict__ src, const float *__restrict__ ref,
                                                            float *__restrict__ out, float *__restrict__ err, uint32_t channels, uint32_t n)
{
  extern __shared__ float lms_smem[];
  float *win = lms_smem;                                  // [2 ntaps][32]
  float *cf = win + 2 * ntaps * 32;                       // [ntaps][32]
  float *tx = cf + ntaps * 32;                            // four tiles [32][33]: src, ref, out, err
  float *tr = tx + 32 * 33, *to = tr + 32 * 33, *te = to + 32 * 33;
  const int lane = threadIdx.x;
  const uint32_t c0 = blockIdx.x * 32u, c = c0 + lane;
  const bool valid = c < channels;
  const uint32_t nch = min (32u, channels - c0);
  float energy = 0.f, x0 = 0.f;
  if (valid)
  {
    const float *st = state + (size_t) c * (ntaps + 1);
    for (int k = 0; k + 1 < ntaps; k++) { const float v = st[k]; win[(1 + k) * 32 + lane] = v; win[(1 + k + ntaps) * 32 + lane] = v; }
    energy = st[ntaps - 1]; x0 = st[ntaps];
    for (int k = 0; k < ntaps; k++) cf[k * 32 + lane] = coeffs[(size_t) c * ntaps + k];
  }
  int p = 0;
  for (uint32_t t0 = 0; t0 < n; t0 += kLmsTile)
  {
    const uint32_t tn = min ((uint32_t) kLmsTile, n - t0);
    __syncwarp ();
    for (uint32_t r = 0; r < nch; r++)
      if ((uint32_t) lane < tn)
      {
        tx[r * 33 + lane] = src[(size_t) (c0 + r) * n + t0 + lane];
        tr[r * 33 + lane] = ref[(size_t) (c0 + r) * n + t0 + lane];
      }
    __syncwarp ();
    if (valid)
      for (uint32_t i = 0; i < tn; i++)
      {
        const float in = tx[lane * 33 + i];
        win[p * 32 + lane] = in; win[(p + ntaps) * 32 + lane] = in;
        energy = __fsub_rn (energy, __fmul_rn (x0, x0));
        energy = __fadd_rn (energy, __fmul_rn (in, in));
        const float *w0 = win + (p + 1) * 32 + lane;     // the window, oldest first
        float sum = 0.0f;
        for (int k = 0; k < ntaps; k++) sum = __fadd_rn (sum, __fmul_rn (w0[k * 32], cf[k * 32 + lane]));
        const float e = __fsub_rn (tr[lane * 33 + i], sum);
        to[lane * 33 + i] = sum; te[lane * 33 + i] = e;
        const float w = __fdiv_rn (__fmul_rn (e, mu), __fadd_rn (energy, 0.000000119209289f));
        for (int k = 0; k < ntaps; k++) cf[k * 32 + lane] = __fadd_rn (cf[k * 32 + lane], __fmul_rn (w, w0[k * 32]));
        x0 = w0[0];
        p = (p + 1 == ntaps) ? 0 : p + 1;
      }
    __syncwarp ();
    for (uint32_t r = 0; r < nch; r++)
      if ((uint32_t) lane < tn)
      {
        out[(size_t) (c0 + r) * n + t0 + lane] = to[r * 33 + lane];
        err[(size_t) (c0 + r) * n + t0 + lane] = te[r * 33 + lane];
      }
  }
  if (valid)
  {
    float *st = state + (size_t) c * (ntaps + 1);
    for (int k = 0; k + 1 < ntaps; k++) st[k] = win[(p + 1 + k) * 32 + lane];
    st[ntaps - 1] = energy; st[ntaps] = x0;
    for (int k = 0; k < ntaps; k++) coeffs[(size_t) c * ntaps + k] = cf[k * 32 + lane];
  }
}

// =====================================================================================================================
// biquad cascades — channel-parallel: one thread per (channel [, stereo rail]), sample-sequential, stage-major order
// is irrelevant for a causal cascade so the cascade runs sample-major with the state in registers.
// =====================================================================================================================
constexpr int kMaxSt = SLB_MAX_STAGES;
struct BqCoefF32 { float c[5 * kMaxSt]; int ns; };
struct BqCoefQ15 { int16_t c[6 * kMaxSt]; int ns, shift; };
struct BqCoefQ31 { int c[5 * kMaxSt]; int ns; unsigned shift; };

// arm_biquad_cascade_df2T_f32.c:551-562 (mono), arm_biquad_cascade_stereo_df2T_f32.c:428-457 (rails = 2, interleaved)
__global__ void biquad_df2T_kernel (BqCoefF32 K, float *__restrict__ state, const float *__restrict__ src, float *__restrict__ dst, uint32_t channels, uint32_t n, int rails)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= channels * rails) return;
  const uint32_t c = t / rails, rail = t % rails;
  float d1[kMaxSt], d2[kMaxSt];
  float *st = state + (size_t) c * 2 * rails * K.ns;
  for (int s = 0; s < K.ns; s++) { d1[s] = st[2 * rails * s + 2 * rail]; d2[s] = st[2 * rails * s + 2 * rail + 1]; }
  const float *x = src + (size_t) c * n * rails + rail; float *y = dst + (size_t) c * n * rails + rail;
  for (uint32_t i = 0; i < n; i++)
  {
    float v = x[(size_t) i * rails];
#pragma unroll
    for (int s = 0; s < kMaxSt; s++)
    {
      if (s < K.ns)
      {
        const float *k = K.c + 5 * s;
        const float acc = __fadd_rn (__fmul_rn (k[0], v), d1[s]);
        d1[s] = __fadd_rn (__fadd_rn (__fmul_rn (k[1], v), __fmul_rn (k[3], acc)), d2[s]);
        d2[s] = __fadd_rn (__fmul_rn (k[2], v), __fmul_rn (k[4], acc));
        v = acc;
      }
    }
    y[(size_t) i * rails] = v;
  }
  for (int s = 0; s < K.ns; s++) { st[2 * rails * s + 2 * rail] = d1[s]; st[2 * rails * s + 2 * rail + 1] = d2[s]; }
}
// arm_biquad_cascade_df1_f32.c:367 : acc = ((((b0 x)+(b1 x1))+(b2 x2))+(a1 y1))+(a2 y2), state {x1,x2,y1,y2}
__global__ void biquad_df1_f32_kernel (BqCoefF32 K, float *__restrict__ state, const float *__restrict__ src, float *__restrict__ dst, uint32_t channels, uint32_t n)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  float x1[kMaxSt], x2[kMaxSt], y1[kMaxSt], y2[kMaxSt];
  float *st = state + (size_t) c * 4 * K.ns;
  for (int s = 0; s < K.ns; s++) { x1[s] = st[4 * s]; x2[s] = st[4 * s + 1]; y1[s] = st[4 * s + 2]; y2[s] = st[4 * s + 3]; }
  for (uint32_t i = 0; i < n; i++)
  {
    float v = src[(size_t) c * n + i];
#pragma unroll
    for (int s = 0; s < kMaxSt; s++)
      if (s < K.ns)
      {
        const float *k = K.c + 5 * s;
        const float acc = __fadd_rn (__fadd_rn (__fadd_rn (__fadd_rn (__fmul_rn (k[0], v), __fmul_rn (k[1], x1[s])), __fmul_rn (k[2], x2[s])), __fmul_rn (k[3], y1[s])), __fmul_rn (k[4], y2[s]));
        x2[s] = x1[s]; x1[s] = v; y2[s] = y1[s]; y1[s] = acc; v = acc;
      }
    dst[(size_t) c * n + i] = v;
  }
  for (int s = 0; s < K.ns; s++) { st[4 * s] = x1[s]; st[4 * s + 1] = x2[s]; st[4 * s + 2] = y1[s]; st[4 * s + 3] = y2[s]; }
}
// arm_biquad_cascade_df1_q15.c:300-400 : q63 acc, >> (15 - postShift) truncated to q31, sat16
__global__ void biquad_df1_q15_kernel (BqCoefQ15 K, int16_t *__restrict__ state, const int16_t *__restrict__ src, int16_t *__restrict__ dst, uint32_t channels, uint32_t n)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  int x1[kMaxSt], x2[kMaxSt], y1[kMaxSt], y2[kMaxSt];
  int16_t *st = state + (size_t) c * 4 * K.ns;
  for (int s = 0; s < K.ns; s++) { x1[s] = st[4 * s]; x2[s] = st[4 * s + 1]; y1[s] = st[4 * s + 2]; y2[s] = st[4 * s + 3]; }
  for (uint32_t i = 0; i < n; i++)
  {
    int v = src[(size_t) c * n + i];
#pragma unroll
    for (int s = 0; s < kMaxSt; s++)
      if (s < K.ns)
      {
        const int16_t *k = K.c + 6 * s;
        long long acc = (long long) ((int) k[0] * v);
        acc += (int) k[2] * x1[s]; acc += (int) k[3] * x2[s]; acc += (int) k[4] * y1[s]; acc += (int) k[5] * y2[s];
        const int y = ssat16 ((int) (acc >> K.shift));
        x2[s] = x1[s]; x1[s] = v; y2[s] = y1[s]; y1[s] = y; v = y;
      }
    dst[(size_t) c * n + i] = (int16_t) v;
  }
  for (int s = 0; s < K.ns; s++) { st[4 * s] = (int16_t) x1[s]; st[4 * s + 1] = (int16_t) x2[s]; st[4 * s + 2] = (int16_t) y1[s]; st[4 * s + 3] = (int16_t) y2[s]; }
}
// arm_biquad_cascade_df1_q31.c:100-200 : wrapping q63 accumulator, low 32 bits of acc >> (31 - postShift)
__global__ void biquad_df1_q31_kernel (BqCoefQ31 K, int *__restrict__ state, const int *__restrict__ src, int *__restrict__ dst, uint32_t channels, uint32_t n)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  int x1[kMaxSt], x2[kMaxSt], y1[kMaxSt], y2[kMaxSt];
  int *st = state + (size_t) c * 4 * K.ns;
  for (int s = 0; s < K.ns; s++) { x1[s] = st[4 * s]; x2[s] = st[4 * s + 1]; y1[s] = st[4 * s + 2]; y2[s] = st[4 * s + 3]; }
  for (uint32_t i = 0; i < n; i++)
  {
    int v = src[(size_t) c * n + i];
#pragma unroll
    for (int s = 0; s < kMaxSt; s++)
      if (s < K.ns)
      {
        const int *k = K.c + 5 * s;
        unsigned long long acc = (unsigned long long) ((long long) k[0] * v);
        acc += (unsigned long long) ((long long) k[1] * x1[s]); acc += (unsigned long long) ((long long) k[2] * x2[s]);
        acc += (unsigned long long) ((long long) k[3] * y1[s]); acc += (unsigned long long) ((long long) k[4] * y2[s]);
        const int y = (int) (unsigned) (acc >> K.shift);
        x2[s] = x1[s]; x1[s] = v; y2[s] = y1[s]; y1[s] = y; v = y;
      }
    dst[(size_t) c * n + i] = v;
  }
  for (int s = 0; s < K.ns; s++) { st[4 * s] = x1[s]; st[4 * s + 1] = x2[s]; st[4 * s + 2] = y1[s]; st[4 * s + 3] = y2[s]; }
}

// =====================================================================================================================
// block statistics — one thread per (channel, block), sequential in the reference's order (blocks are 48 frames, so
// there are plenty of them); kind 0 max (first maximum wins, arm_max_f32.c:58) 1 rms (arm_rms_f32.c:122) 2 power 3 mean
// =====================================================================================================================
__global__ void stats_f32_kernel (const float *__restrict__ src, float *__restrict__ out, uint32_t *__restrict__ idx, size_t nblocks, uint32_t block, int kind)
{
  const size_t b = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const float *s = src + b * block;
  if (kind == 0)
  {
    float m = s[0]; uint32_t mi = 0;
    for (uint32_t i = 1; i < block; i++) if (m < s[i]) { m = s[i]; mi = i; }
    out[b] = m; if (idx) idx[b] = mi;
  }
  else
  {
    float sum = 0.0f;
    for (uint32_t i = 0; i < block; i++) sum = __fadd_rn (sum, kind == 3 ? s[i] : __fmul_rn (s[i], s[i]));
    if (kind == 1) { const float v = __fdiv_rn (sum, (float) block); out[b] = v >= 0.0f ? __fsqrt_rn (v) : 0.0f; }
    else if (kind == 2) out[b] = sum;
    else out[b] = __fdiv_rn (sum, (float) block);
  }
}
// arm_max_q15.c, arm_rms_q15.c:107-131 (q63 sum, sat16((sum / N) >> 15), arm_sqrt_q15)
__global__ void stats_q15_kernel (const int16_t *__restrict__ src, int16_t *__restrict__ out, uint32_t *__restrict__ idx, size_t nblocks, uint32_t block, int kind)
{
  const size_t b = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const int16_t *s = src + b * block;
  if (kind == 0)
  {
    int16_t m = s[0]; uint32_t mi = 0;
    for (uint32_t i = 1; i < block; i++) if (m < s[i]) { m = s[i]; mi = i; }
    out[b] = m; if (idx) idx[b] = mi;
  }
  else
  {
    long long sum = 0;
    for (uint32_t i = 0; i < block; i++) sum += (int) s[i] * (int) s[i];
    out[b] = sqrt_q15 ((int16_t) ssat16 ((int) ((sum / (long long) block) >> 15)));
  }
}

// =====================================================================================================================
// batched complex FFT, N = 16..4096 (arm_cfft_f32.c:562-615 contract: in place, interleaved, forward unscaled, inverse
// = conj in / forward / conj and 1/N out, natural order). One CTA per transform, radix-2 Stockham in shared memory with
// twiddles from sincospif (float32 result within 2 ulp). Tolerance class, like every float FFT against the reference.
// =====================================================================================================================
__global__ void cfft_f32_kernel (float2 *__restrict__ data, int N, int log2n, int inverse)
{
  extern __shared__ float2 sm[];
  float2 *a = sm, *b = sm + N;
  float2 *x = data + (size_t) blockIdx.x * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { float2 v = x[i]; if (inverse) v.y = -v.y; a[i] = v; }
  __syncthreads ();
  for (int s = 0, ns = 1; s < log2n; s++, ns <<= 1)
  {
    for (int j = threadIdx.x; j < N / 2; j += blockDim.x)
    {
      const int k = j & (ns - 1);
      float sn, cs; sincospif (-(float) k / (float) ns, &sn, &cs);
      const float2 u = a[j], t = a[j + N / 2];
      const float2 w = make_float2 (t.x * cs - t.y * sn, t.x * sn + t.y * cs);
      const int o = ((j - k) << 1) + k;
      b[o] = make_float2 (u.x + w.x, u.y + w.y); b[o + ns] = make_float2 (u.x - w.x, u.y - w.y);
    }
    __syncthreads ();
    float2 *t = a; a = b; b = t;
  }
  const float sc = inverse ? 1.0f / (float) N : 1.0f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { float2 v = a[i]; x[i] = make_float2 (v.x * sc, inverse ? -v.y * sc : v.y); }
}

// =====================================================================================================================
// batched real FFT, N = 32..4096 (arm_rfft_fast_f32.c:288-315 contract). Forward: the N reals are transformed as N/2
// complex points (arm_cfft_f32 of half length) and split (stage_rfft_f32, :38-141): out[0] = X[0], out[1] = X[N/2] (both
// real), then (Re, Im) of X[1..N/2-1]. Inverse: merge (merge_rfft_f32, :143-227) then the half-length inverse transform,
// so that rfft followed by rifft returns the input. One CTA per transform, everything in shared memory.
// =====================================================================================================================
__global__ void rfft_fast_f32_kernel (const float *__restrict__ in, float *__restrict__ out, int N, int log2m, int inverse)
{
  extern __shared__ float2 sm[];
  const int M = N / 2;
  float2 *a = sm, *b = sm + M;
  const float *x = in + (size_t) blockIdx.x * N;
  float *y = out + (size_t) blockIdx.x * N;
  if (!inverse)
    for (int i = threadIdx.x; i < M; i += blockDim.x) a[i] = make_float2 (x[2 * i], x[2 * i + 1]);
  else
  {
    // Zm[k] = 0.5 [(X[k] + conj X[M-k]) + j e^{+2 pi j k / N} (X[k] - conj X[M-k])], conjugated for the forward engine below
    for (int k = threadIdx.x; k < M; k += blockDim.x)
    {
      const float2 xk = (k == 0) ? make_float2 (x[0], 0.f) : make_float2 (x[2 * k], x[2 * k + 1]);
      const float2 xm = (k == 0) ? make_float2 (x[1], 0.f) : make_float2 (x[2 * (M - k)], x[2 * (M - k) + 1]);
      const float sr = xk.x + xm.x, si = xk.y - xm.y, dr = xk.x - xm.x, di = xk.y + xm.y;   // X[k] +/- conj X[M-k]
      float sn, cs; sincospif (2.0f * (float) k / (float) N, &sn, &cs);
      // j e^{j t} (dr + j di) = (-sn dr - cs di) + j (cs dr - sn di)
      const float zr = 0.5f * (sr + (-sn * dr - cs * di)), zi = 0.5f * (si + (cs * dr - sn * di));
      a[k] = make_float2 (zr, -zi);
    }
  }
  __syncthreads ();
  for (int s = 0, ns = 1; s < log2m; s++, ns <<= 1)
  {
    for (int j = threadIdx.x; j < M / 2; j += blockDim.x)
    {
      const int k = j & (ns - 1);
      float sn, cs; sincospif (-(float) k / (float) ns, &sn, &cs);
      const float2 u = a[j], t = a[j + M / 2];
      const float2 w = make_float2 (t.x * cs - t.y * sn, t.x * sn + t.y * cs);
      const int o = ((j - k) << 1) + k;
      b[o] = make_float2 (u.x + w.x, u.y + w.y); b[o + ns] = make_float2 (u.x - w.x, u.y - w.y);
    }
    __syncthreads ();
    float2 *t = a; a = b; b = t;
  }
  if (inverse)
  {
    const float sc = 1.0f / (float) M;                 // arm_cfft_f32.c:604-614
    for (int i = threadIdx.x; i < M; i += blockDim.x) { y[2 * i] = a[i].x * sc; y[2 * i + 1] = -a[i].y * sc; }
  }
  else
  {
    // X[k] = 0.5 (Z[k] + conj Z[M-k]) - 0.5 j e^{-2 pi j k / N} (Z[k] - conj Z[M-k])
    for (int k = threadIdx.x; k < M; k += blockDim.x)
    {
      if (k == 0) { y[0] = a[0].x + a[0].y; y[1] = a[0].x - a[0].y; continue; }
      const float2 zk = a[k], zm = a[M - k];
      const float sr = zk.x + zm.x, si = zk.y - zm.y, dr = zk.x - zm.x, di = zk.y + zm.y;
      float sn, cs; sincospif (-2.0f * (float) k / (float) N, &sn, &cs);
      // -j e^{j t} (dr + j di) = (sn dr + cs di) + j (sn di - cs dr)
      y[2 * k] = 0.5f * (sr + (sn * dr + cs * di)); y[2 * k + 1] = 0.5f * (si + (sn * di - cs * dr));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------------------------------
static float g_sin_table[513];
static bool g_sin_ready = false;
const float *host_sin_table ()
{
  if (!g_sin_ready)
  {
    char buf[32];
    for (int i = 0; i <= 512; i++)
    {   // the reference's table holds sin(2 pi i / 512) written with 8 decimals (arm_common_tables.c:21895)
      std::snprintf (buf, sizeof buf, "%.8f", std::sin (2.0 * 3.14159265358979323846 * (double) i / 512.0));
      g_sin_table[i] = std::strtof (buf, nullptr);
    }
    g_sin_ready = true;
  }
  return g_sin_table;
}

}  // namespace sl

// =====================================================================================================================
// C ABI (include/selenite_b200.h, "stage library"). All array pointers are DEVICE pointers laid out [channels][n];
// coefficient pointers are HOST pointers (they are tiny and copied with the launch).
// =====================================================================================================================
using namespace sl;

#define ST_BEGIN(ctx)                                                                    \
  if (!(ctx)) return SLB_ERR_ARG;                                                        \
  { cudaError_t e0_ = cudaSetDevice (ctx_device (ctx)); if (e0_ != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e0_)); } \
  const size_t C = ctx_channels (ctx); (void) C;                                         \
  cudaStream_t st = (cudaStream_t) stream
#define ST_END(ctx, rc) do { int rc_ = (rc); if (rc_) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString ((cudaError_t) rc_)); ctx_count_launch (ctx); return SLB_OK; } while (0)

extern "C" {

int slb_st_q15_to_float (slb_ctx *ctx, const int16_t *src, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpQ15ToFloat{ src, dst }, C * n, st)); }
int slb_st_float_to_q15 (slb_ctx *ctx, const float *src, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpFloatToQ15{ src, dst }, C * n, st)); }
int slb_st_scale_f32 (slb_ctx *ctx, const float *src, float scale, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpScaleF32{ src, scale, dst }, C * n, st)); }
int slb_st_mult_f32 (slb_ctx *ctx, const float *a, const float *b, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpMultF32{ a, b, dst }, C * n, st)); }
int slb_st_add_f32 (slb_ctx *ctx, const float *a, const float *b, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpAddF32{ a, b, dst }, C * n, st)); }
int slb_st_sub_f32 (slb_ctx *ctx, const float *a, const float *b, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpSubF32{ a, b, dst }, C * n, st)); }
int slb_st_abs_f32 (slb_ctx *ctx, const float *a, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpAbsF32{ a, dst }, C * n, st)); }
int slb_st_scale_q15 (slb_ctx *ctx, const int16_t *src, int16_t scale_fract, int32_t shift, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpScaleQ15{ src, scale_fract, 15 - shift, dst }, C * n, st)); }
int slb_st_add_q15 (slb_ctx *ctx, const int16_t *a, const int16_t *b, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpAddQ15{ a, b, dst }, C * n, st)); }
int slb_st_sub_q15 (slb_ctx *ctx, const int16_t *a, const int16_t *b, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpSubQ15{ a, b, dst }, C * n, st)); }
int slb_st_abs_q15 (slb_ctx *ctx, const int16_t *a, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpAbsQ15{ a, dst }, C * n, st)); }
int slb_st_shift_q15 (slb_ctx *ctx, const int16_t *a, int32_t shift, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpShiftQ15{ a, shift, dst }, C * n, st)); }
int slb_st_cmplx_mult_cmplx_f32 (slb_ctx *ctx, const float *a, const float *b, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpCmulF32{ (const float2 *) a, (const float2 *) b, (float2 *) dst }, C * n, st)); }
int slb_st_cmplx_mult_real_f32 (slb_ctx *ctx, const float *a, const float *r, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpCmulRealF32{ (const float2 *) a, r, (float2 *) dst }, C * n, st)); }
int slb_st_cmplx_conj_f32 (slb_ctx *ctx, const float *a, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpConjF32{ (const float2 *) a, (float2 *) dst }, C * n, st)); }
int slb_st_cmplx_mag_f32 (slb_ctx *ctx, const float *a, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpMagF32{ (const float2 *) a, dst }, C * n, st)); }
int slb_st_cmplx_mag_squared_f32 (slb_ctx *ctx, const float *a, float *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpMagSqF32{ (const float2 *) a, dst }, C * n, st)); }
int slb_st_cmplx_mag_q15 (slb_ctx *ctx, const int16_t *a, int16_t *dst, uint32_t n, void *stream) { ST_BEGIN (ctx); ST_END (ctx, launch_ew (OpMagQ15{ (const short2 *) a, dst }, C * n, st)); }

static int sincos_common (slb_ctx *ctx, const float *x, float *dst, uint32_t n, void *stream, int is_cos)
{
  ST_BEGIN (ctx);
  float *tab = (float *) ctx_scratch_on (ctx, 513 * sizeof (float), st);
  if (!tab) return SLB_ERR_CUDA;
  cudaError_t e = cudaMemcpyAsync (tab, host_sin_table (), 513 * sizeof (float), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  ST_END (ctx, launch_ew (OpSinCosF32{ x, dst, tab, is_cos }, C * n, st));
}
int slb_st_sin_f32 (slb_ctx *ctx, const float *x, float *dst, uint32_t n, void *stream) { return sincos_common (ctx, x, dst, n, stream, 0); }
int slb_st_cos_f32 (slb_ctx *ctx, const float *x, float *dst, uint32_t n, void *stream) { return sincos_common (ctx, x, dst, n, stream, 1); }

}  // extern "C"

// ---- FIR family. hist: device [channels][hist_len] with hist_len = ntaps-1 (interpolator: ntaps/L - 1); updated in place.
template <typename T, int MODE>
static int fir_common (slb_ctx *ctx, const T *h_coeffs, uint32_t ntaps, uint32_t M, uint32_t L, T *hist, const T *src, T *dst, uint32_t n, void *stream)
{
  ST_BEGIN (ctx);
  if (!h_coeffs || !hist || !src || !dst || ntaps == 0 || M == 0 || L == 0 || n % M != 0 || ntaps % L != 0) return ctx_fail (ctx, SLB_ERR_ARG, "bad FIR arguments");
  const size_t hist_len = ntaps / L - 1;
  char *scr = (char *) ctx_scratch_on (ctx, ntaps * sizeof (T) + C * hist_len * sizeof (T) + 64, st);
  if (!scr) return SLB_ERR_CUDA;
  T *d_coeffs = (T *) scr; T *hist_new = (T *) (scr + ((ntaps * sizeof (T) + 63) / 64) * 64);
  cudaError_t e = cudaMemcpyAsync (d_coeffs, h_coeffs, ntaps * sizeof (T), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  const uint32_t n_out = (n / M) * L;
  unsigned bx = (n_out + 255) / 256; if (bx > 1024) bx = 1024; if (bx < 1) bx = 1;
  fir_kernel<T, MODE><<<dim3 (bx, (unsigned) C), 256, 0, st>>> (d_coeffs, (int) ntaps, (int) M, (int) L, hist, hist_new, src, dst, n);
  e = cudaGetLastError ();
  if (e == cudaSuccess && hist_len) e = cudaMemcpyAsync (hist, hist_new, C * hist_len * sizeof (T), cudaMemcpyDeviceToDevice, st);
  ST_END (ctx, (int) e);
}
extern "C" {

int slb_st_fir_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ntaps, float *hist, const float *src, float *dst, uint32_t n, void *stream) { return fir_common<float, 0> (ctx, coeffs, ntaps, 1, 1, hist, src, dst, n, stream); }
int slb_st_fir_q15 (slb_ctx *ctx, const int16_t *coeffs, uint32_t ntaps, int16_t *hist, const int16_t *src, int16_t *dst, uint32_t n, void *stream) { return fir_common<int16_t, 1> (ctx, coeffs, ntaps, 1, 1, hist, src, dst, n, stream); }
int slb_st_fir_fast_q15 (slb_ctx *ctx, const int16_t *coeffs, uint32_t ntaps, int16_t *hist, const int16_t *src, int16_t *dst, uint32_t n, void *stream) { return fir_common<int16_t, 2> (ctx, coeffs, ntaps, 1, 1, hist, src, dst, n, stream); }
int slb_st_fir_q31 (slb_ctx *ctx, const int32_t *coeffs, uint32_t ntaps, int32_t *hist, const int32_t *src, int32_t *dst, uint32_t n, void *stream) { return fir_common<int32_t, 3> (ctx, coeffs, ntaps, 1, 1, hist, src, dst, n, stream); }
int slb_st_fir_decimate_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ntaps, uint32_t M, float *hist, const float *src, float *dst, uint32_t n, void *stream) { return fir_common<float, 0> (ctx, coeffs, ntaps, M, 1, hist, src, dst, n, stream); }
int slb_st_fir_decimate_q15 (slb_ctx *ctx, const int16_t *coeffs, uint32_t ntaps, uint32_t M, int16_t *hist, const int16_t *src, int16_t *dst, uint32_t n, void *stream) { return fir_common<int16_t, 1> (ctx, coeffs, ntaps, M, 1, hist, src, dst, n, stream); }
int slb_st_fir_interpolate_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ntaps, uint32_t L, float *hist, const float *src, float *dst, uint32_t n, void *stream) { return fir_common<float, 0> (ctx, coeffs, ntaps, 1, L, hist, src, dst, n, stream); }
int slb_st_fir_interpolate_q15 (slb_ctx *ctx, const int16_t *coeffs, uint32_t ntaps, uint32_t L, int16_t *hist, const int16_t *src, int16_t *dst, uint32_t n, void *stream) { return fir_common<int16_t, 1> (ctx, coeffs, ntaps, 1, L, hist, src, dst, n, stream); }
// arm_fir_decimate_q31.c:60 / arm_fir_interpolate_q31.c:62: q63 accumulator of q31 x q31 products, result (q31) (acc >> 31), no saturation
int slb_st_fir_decimate_q31 (slb_ctx *ctx, const int32_t *coeffs, uint32_t ntaps, uint32_t M, int32_t *hist, const int32_t *src, int32_t *dst, uint32_t n, void *stream) { return fir_common<int32_t, 3> (ctx, coeffs, ntaps, M, 1, hist, src, dst, n, stream); }
int slb_st_fir_interpolate_q31 (slb_ctx *ctx, const int32_t *coeffs, uint32_t ntaps, uint32_t L, int32_t *hist, const int32_t *src, int32_t *dst, uint32_t n, void *stream) { return fir_common<int32_t, 3> (ctx, coeffs, ntaps, 1, L, hist, src, dst, n, stream); }

// ---- fixed-point FFTs (sl_fft_fixed.cu): data [channels][count][2 N], in place, natural output order
int slb_st_cfft_q15 (slb_ctx *ctx, int16_t *data, uint32_t N, uint32_t count, int ifft, void *stream)
{
  ST_BEGIN (ctx);
  if (!data || count == 0 || N < 16 || N > 4096 || (N & (N - 1))) return ctx_fail (ctx, SLB_ERR_ARG, "cfft_q15: N = 16 .. 4096, a power of two");
  const size_t tw_bytes = (size_t) 3 * N / 4 * 2 * sizeof (int16_t);
  int16_t *d_tw = (int16_t *) ctx_scratch_on (ctx, tw_bytes, st);
  if (!d_tw) return SLB_ERR_CUDA;
  cudaError_t e = cudaMemcpyAsync (d_tw, fft_twiddle_q15 (N), tw_bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  ST_END (ctx, launch_cfft_q15 (data, N, C * count, ifft, d_tw, st));
}
int slb_st_cfft_q31 (slb_ctx *ctx, int32_t *data, uint32_t N, uint32_t count, int ifft, void *stream)
{
  ST_BEGIN (ctx);
  if (!data || count == 0 || N < 16 || N > 4096 || (N & (N - 1))) return ctx_fail (ctx, SLB_ERR_ARG, "cfft_q31: N = 16 .. 4096, a power of two");
  const size_t tw_bytes = (size_t) 3 * N / 4 * 2 * sizeof (int32_t);
  int32_t *d_tw = (int32_t *) ctx_scratch_on (ctx, tw_bytes, st);
  if (!d_tw) return SLB_ERR_CUDA;
  cudaError_t e = cudaMemcpyAsync (d_tw, fft_twiddle_q31 (N), tw_bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  ST_END (ctx, launch_cfft_q31 (data, N, C * count, ifft, d_tw, st));
}
// design check (tests/test_fft_fixed_tables.py): the regenerated twiddle tables, 3 N / 4 (cos, sin) pairs
int slb_design_twiddle_q15 (uint32_t N, int16_t *out) { if (!out || N < 16 || N > 4096 || (N & (N - 1))) return SLB_ERR_ARG; std::memcpy (out, fft_twiddle_q15 (N), (size_t) 3 * N / 4 * 2 * sizeof (int16_t)); return SLB_OK; }
int slb_design_twiddle_q31 (uint32_t N, int32_t *out) { if (!out || N < 16 || N > 4096 || (N & (N - 1))) return SLB_ERR_ARG; std::memcpy (out, fft_twiddle_q31 (N), (size_t) 3 * N / 4 * 2 * sizeof (int32_t)); return SLB_OK; }

}  // extern "C"
namespace sl {
// arm_biquad_cascade_df1_q15 on [channels][n] for the RX-SSB-q15 chain's optional audio filter (sl_rx_ssb_q15.cu)
int launch_biquad_df1_q15 (const int16_t *coeffs6, uint32_t ns, int32_t postshift, int16_t *d_state, const int16_t *d_src, int16_t *d_dst, uint32_t channels, uint32_t n, void *stream)
{
  if (ns == 0 || ns > (uint32_t) kMaxSt) return (int) cudaErrorInvalidValue;
  BqCoefQ15 K; K.ns = (int) ns; K.shift = 15 - postshift; std::memcpy (K.c, coeffs6, 6 * ns * sizeof (int16_t));
  biquad_df1_q15_kernel<<<(channels + 63) / 64, 64, 0, (cudaStream_t) stream>>> (K, d_state, d_src, d_dst, channels, n);
  return (int) cudaGetLastError ();
}
}  // namespace sl
extern "C" {
// ---- normalised LMS. coeffs: device [channels][ntaps] (every channel adapts its own), state: device [channels][ntaps + 1] =
// the ntaps - 1 previous samples oldest first (the head of the CMSIS state buffer), then energy and x0 of the instance; both in place
int slb_st_lms_norm_f32 (slb_ctx *ctx, float *coeffs, uint32_t ntaps, float mu, float *state, const float *src, const float *ref,
                         float *out, float *err, uint32_t n, void *stream)
{
  ST_BEGIN (ctx);
  if (!coeffs || !state || !src || !ref || !out || !err || ntaps < 2 || ntaps > 128 || n == 0) return ctx_fail (ctx, SLB_ERR_ARG, "lms_norm: 2..128 taps");
  const size_t smem = ((size_t) 3 * ntaps * 32 + 4 * 32 * 33) * sizeof (float);
  cudaError_t e = cudaFuncSetAttribute (lms_norm_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return ctx_fail (ctx, SLB_ERR_CUDA, cudaGetErrorString (e));
  lms_norm_f32_kernel<<<(unsigned) ((C + 31) / 32), 32, smem, st>>> (coeffs, (int) ntaps, mu, state, src, ref, out, err, (uint32_t) C, n);
  ST_END (ctx, (int) cudaGetLastError ());
}

// ---- biquads. state: device, CMSIS layout per channel (df2T 2/stage, stereo df2T 4/stage, df1 4/stage)
static int bq_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ns, float *state, const float *src, float *dst, uint32_t n, void *stream, int kind)
{
  ST_BEGIN (ctx);
  if (!coeffs || !state || !src || !dst || ns == 0 || ns > (uint32_t) kMaxSt) return ctx_fail (ctx, SLB_ERR_ARG, "1..4 biquad stages");
  BqCoefF32 K; K.ns = (int) ns; std::memcpy (K.c, coeffs, 5 * ns * sizeof (float));
  const int rails = kind == 1 ? 2 : 1;
  const unsigned threads = (unsigned) C * rails, bx = (threads + 63) / 64;
  if (kind == 2) biquad_df1_f32_kernel<<<bx, 64, 0, st>>> (K, state, src, dst, (uint32_t) C, n);
  else biquad_df2T_kernel<<<bx, 64, 0, st>>> (K, state, src, dst, (uint32_t) C, n, rails);
  ST_END (ctx, (int) cudaGetLastError ());
}
int slb_st_biquad_df2T_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ns, float *state, const float *src, float *dst, uint32_t n, void *stream) { return bq_f32 (ctx, coeffs, ns, state, src, dst, n, stream, 0); }
int slb_st_biquad_stereo_df2T_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ns, float *state, const float *src, float *dst, uint32_t nframes, void *stream) { return bq_f32 (ctx, coeffs, ns, state, src, dst, nframes, stream, 1); }
int slb_st_biquad_df1_f32 (slb_ctx *ctx, const float *coeffs, uint32_t ns, float *state, const float *src, float *dst, uint32_t n, void *stream) { return bq_f32 (ctx, coeffs, ns, state, src, dst, n, stream, 2); }
int slb_st_biquad_df1_q15 (slb_ctx *ctx, const int16_t *coeffs, uint32_t ns, int32_t postshift, int16_t *state, const int16_t *src, int16_t *dst, uint32_t n, void *stream)
{
  ST_BEGIN (ctx);
  if (!coeffs || !state || !src || !dst || ns == 0 || ns > (uint32_t) kMaxSt) return ctx_fail (ctx, SLB_ERR_ARG, "1..4 biquad stages");
  BqCoefQ15 K; K.ns = (int) ns; K.shift = 15 - postshift; std::memcpy (K.c, coeffs, 6 * ns * sizeof (int16_t));
  biquad_df1_q15_kernel<<<(unsigned) (C + 63) / 64, 64, 0, st>>> (K, state, src, dst, (uint32_t) C, n);
  ST_END (ctx, (int) cudaGetLastError ());
}
int slb_st_biquad_df1_q31 (slb_ctx *ctx, const int32_t *coeffs, uint32_t ns, int32_t postshift, int32_t *state, const int32_t *src, int32_t *dst, uint32_t n, void *stream)
{
  ST_BEGIN (ctx);
  if (!coeffs || !state || !src || !dst || ns == 0 || ns > (uint32_t) kMaxSt) return ctx_fail (ctx, SLB_ERR_ARG, "1..4 biquad stages");
  BqCoefQ31 K; K.ns = (int) ns; K.shift = 31u - (unsigned) postshift; std::memcpy (K.c, coeffs, 5 * ns * sizeof (int32_t));
  biquad_df1_q31_kernel<<<(unsigned) (C + 63) / 64, 64, 0, st>>> (K, state, src, dst, (uint32_t) C, n);
  ST_END (ctx, (int) cudaGetLastError ());
}

// ---- block statistics: out [channels][n / block]
static int stats_f32 (slb_ctx *ctx, const float *src, uint32_t n, uint32_t block, float *out, uint32_t *idx, void *stream, int kind)
{
  ST_BEGIN (ctx);
  if (!src || !out || block == 0 || n % block) return ctx_fail (ctx, SLB_ERR_ARG, "n must be a multiple of block");
  const size_t nb = C * (n / block);
  stats_f32_kernel<<<(unsigned) ((nb + 127) / 128), 128, 0, st>>> (src, out, idx, nb, block, kind);
  ST_END (ctx, (int) cudaGetLastError ());
}
int slb_st_max_f32 (slb_ctx *ctx, const float *src, uint32_t n, uint32_t block, float *out, uint32_t *idx, void *stream) { return stats_f32 (ctx, src, n, block, out, idx, stream, 0); }
int slb_st_rms_f32 (slb_ctx *ctx, const float *src, uint32_t n, uint32_t block, float *out, void *stream) { return stats_f32 (ctx, src, n, block, out, nullptr, stream, 1); }
int slb_st_power_f32 (slb_ctx *ctx, const float *src, uint32_t n, uint32_t block, float *out, void *stream) { return stats_f32 (ctx, src, n, block, out, nullptr, stream, 2); }
int slb_st_mean_f32 (slb_ctx *ctx, const float *src, uint32_t n, uint32_t block, float *out, void *stream) { return stats_f32 (ctx, src, n, block, out, nullptr, stream, 3); }
static int stats_q15 (slb_ctx *ctx, const int16_t *src, uint32_t n, uint32_t block, int16_t *out, uint32_t *idx, void *stream, int kind)
{
  ST_BEGIN (ctx);
  if (!src || !out || block == 0 || n % block) return ctx_fail (ctx, SLB_ERR_ARG, "n must be a multiple of block");
  const size_t nb = C * (n / block);
  stats_q15_kernel<<<(unsigned) ((nb + 127) / 128), 128, 0, st>>> (src, out, idx, nb, block, kind);
  ST_END (ctx, (int) cudaGetLastError ());
}
int slb_st_max_q15 (slb_ctx *ctx, const int16_t *src, uint32_t n, uint32_t block, int16_t *out, uint32_t *idx, void *stream) { return stats_q15 (ctx, src, n, block, out, idx, stream, 0); }
int slb_st_rms_q15 (slb_ctx *ctx, const int16_t *src, uint32_t n, uint32_t block, int16_t *out, void *stream) { return stats_q15 (ctx, src, n, block, out, nullptr, stream, 1); }

// ---- batched complex FFT, in place: data [channels][count][2*N] floats
int slb_st_cfft_f32 (slb_ctx *ctx, float *data, uint32_t N, uint32_t count, int ifft, void *stream)
{
  ST_BEGIN (ctx);
  int lg = 0; while ((1u << lg) < N) lg++;
  if (!data || N < 16 || N > 4096 || (1u << lg) != N) return ctx_fail (ctx, SLB_ERR_ARG, "N must be a power of two in 16..4096 (arm_cfft_f32.c:582-599)");
  const size_t smem = (size_t) 2 * N * sizeof (float2);
  if (smem > 48 * 1024) cudaFuncSetAttribute (cfft_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const unsigned threads = N / 2 < 256 ? N / 2 : 256;
  cfft_f32_kernel<<<(unsigned) (C * count), threads, smem, st>>> ((float2 *) data, (int) N, lg, ifft ? 1 : 0);
  ST_END (ctx, (int) cudaGetLastError ());
}

// ---- batched real FFT: in [channels][count][N] floats -> out [channels][count][N] floats (packed as arm_rfft_fast_f32)
int slb_st_rfft_fast_f32 (slb_ctx *ctx, const float *in, float *out, uint32_t N, uint32_t count, int ifft, void *stream)
{
  ST_BEGIN (ctx);
  int lg = 0; while ((1u << lg) < N) lg++;
  if (!in || !out || in == out || N < 32 || N > 4096 || (1u << lg) != N) return ctx_fail (ctx, SLB_ERR_ARG, "N must be a power of two in 32..4096, out of place (arm_rfft_fast_f32.c:288)");
  const size_t smem = (size_t) N * sizeof (float2);
  if (smem > 48 * 1024) cudaFuncSetAttribute (rfft_fast_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const unsigned threads = N / 4 < 256 ? N / 4 : 256;
  rfft_fast_f32_kernel<<<(unsigned) (C * count), threads, smem, st>>> (in, out, (int) N, lg - 1, ifft ? 1 : 0);
  ST_END (ctx, (int) cudaGetLastError ());
}

}  // extern "C"
