// Firmware ring on the GPU: DSP_Buff_TypeDef (Core/Inc/dsp_if.h:87-94) keeps i[] and q[] de-interleaved; here they
// are [channels][ring_frames] int16 planes in HBM. The index arithmetic lives on the host (sl::RingPtrs); these
// kernels only move samples, one thread per (channel, frame).
#include <cuda_runtime.h>
#include "sl_internal.h"

namespace sl {

// DSP_In_Buff_Write / DSP_Out_Buff_Write sample stores (dsp_if.c:286-293 / :165-172): frames 0..n-1, then the
// last frame once more in the following slot.
__global__ void ring_write_kernel (const int16_t *__restrict__ blocks, uint32_t block_stride_frames, int16_t *__restrict__ ring_i,
                                   int16_t *__restrict__ ring_q, uint32_t ring_frames, uint32_t wr0, uint32_t frames)
{
  const uint32_t c = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > frames) return;
  const uint32_t src = (k < frames) ? k : frames - 1u;
  const uint32_t iq = reinterpret_cast<const uint32_t *> (blocks)[(size_t) c * block_stride_frames + src];
  const uint32_t slot = (wr0 + k) % ring_frames;
  ring_i[(size_t) c * ring_frames + slot] = (int16_t) (iq & 0xFFFFu);
  ring_q[(size_t) c * ring_frames + slot] = (int16_t) (iq >> 16);
}

// DSP_In_Buff_Read / DSP_Out_Buff_Read (dsp_if.c:328-339 / :206-217): re-interleave from the ring
__global__ void ring_read_kernel (int16_t *__restrict__ blocks, const int16_t *__restrict__ ring_i, const int16_t *__restrict__ ring_q,
                                  uint32_t ring_frames, uint32_t rd0, uint32_t frames)
{
  const uint32_t c = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= frames) return;
  const uint32_t slot = (rd0 + k) % ring_frames;
  const uint32_t i = (uint16_t) ring_i[(size_t) c * ring_frames + slot], q = (uint16_t) ring_q[(size_t) c * ring_frames + slot];
  reinterpret_cast<uint32_t *> (blocks)[(size_t) c * frames + k] = i | (q << 16);
}

// ---- per-channel cadence -------------------------------------------------------------------------
// Every channel carries its own {enable, rd, wr}; a call serves the channels whose producer / consumer fired this tick
// (active[c] != 0, or all). One thread per channel runs the firmware's pointer arithmetic — the same statements as
// sl::RingPtrs (Core/Src/dsp_if.c:116-180, :204-219, :250-301, :310-340) — and leaves the first slot in ptr[c][3]
// (kRingSkip for a skipped channel); the move kernels then read it per channel.
constexpr uint32_t kRingSkip = 0xFFFFFFFFu;

__global__ void ring_plan_kernel (uint32_t *__restrict__ ptr /* [C][4] */, const uint8_t *__restrict__ active, uint32_t channels, uint32_t N,
                                  int is_write, int is_out, uint32_t frames)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= channels) return;
  uint32_t *p = ptr + 4 * (size_t) c;
  if (active && !active[c]) { p[3] = kRingSkip; return; }
  uint32_t enable = p[0], rd = p[1], wr = p[2], first;
  if (is_write)
  {
    uint32_t gap = 0;
    if (is_out)
    {
      if (!enable) { wr = rd + N / 2; if (wr >= N) wr -= N; enable = 1; }       // dsp_if.c:124-134
      gap = wr; if (rd > wr) gap += N; gap -= rd;                                // dsp_if.c:136-143
    }
    else if (enable) { gap = wr; if (rd > wr) gap += N; gap -= rd; }             // dsp_if.c:254-264
    gap &= 0xFFFFu;
    if (gap > 3u * N / 4u) { if (wr < 1u) wr += N; wr--; }                       // dsp_if.c:145-153 / :266-274
    if (gap < N / 4u) { wr++; if (wr >= N) wr -= N; }                            // dsp_if.c:155-163 / :276-284
    first = wr;
    wr = (wr + frames + 1u) % N;                                                 // dsp_if.c:165-179 / :286-300
    if (wr < 1u) wr += N;
    wr--;
  }
  else
  {
    if (!is_out && !enable) { rd = wr + N / 2; if (rd >= N) rd = 0; enable = 1; }   // dsp_if.c:316-326
    first = rd;
    rd = (rd + frames) % N;                                                      // dsp_if.c:206-217 / :328-339
  }
  p[0] = enable; p[1] = rd; p[2] = wr; p[3] = first;
}

__global__ void ring_write_pc_kernel (const int16_t *__restrict__ blocks, uint32_t block_stride_frames, int16_t *__restrict__ ring_i,
                                      int16_t *__restrict__ ring_q, uint32_t ring_frames, const uint32_t *__restrict__ ptr, uint32_t frames,
                                      const uint32_t *__restrict__ src_off)
{
  const uint32_t c = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t wr0 = ptr[4 * (size_t) c + 3];
  if (k > frames || wr0 == kRingSkip) return;
  const uint32_t src = ((k < frames) ? k : frames - 1u) + (src_off ? src_off[c] : 0u);     // src_off: where this channel's block starts in its source row
  const uint32_t iq = reinterpret_cast<const uint32_t *> (blocks)[(size_t) c * block_stride_frames + src];
  const uint32_t slot = (wr0 + k) % ring_frames;
  ring_i[(size_t) c * ring_frames + slot] = (int16_t) (iq & 0xFFFFu);
  ring_q[(size_t) c * ring_frames + slot] = (int16_t) (iq >> 16);
}

__global__ void ring_read_pc_kernel (int16_t *__restrict__ blocks, const int16_t *__restrict__ ring_i, const int16_t *__restrict__ ring_q,
                                     uint32_t ring_frames, const uint32_t *__restrict__ ptr, uint32_t frames)
{
  const uint32_t c = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t rd0 = ptr[4 * (size_t) c + 3];
  if (k >= frames) return;
  uint32_t v = 0u;                                                               // a skipped channel's block reads as zeros
  if (rd0 != kRingSkip)
  {
    const uint32_t slot = (rd0 + k) % ring_frames;
    v = (uint32_t) (uint16_t) ring_i[(size_t) c * ring_frames + slot] | ((uint32_t) (uint16_t) ring_q[(size_t) c * ring_frames + slot] << 16);
  }
  reinterpret_cast<uint32_t *> (blocks)[(size_t) c * frames + k] = v;
}

int launch_ring_plan (uint32_t *d_ptr, const uint8_t *d_active, uint32_t channels, uint32_t ring_frames, bool is_write, bool is_out,
                      uint32_t frames, void *stream)
{
  ring_plan_kernel<<<(channels + 127) / 128, 128, 0, (cudaStream_t) stream>>> (d_ptr, d_active, channels, ring_frames, is_write, is_out, frames);
  return (int) cudaGetLastError ();
}
int launch_ring_write_pc (const int16_t *d_blocks, uint32_t stride, int16_t *ri, int16_t *rq, uint32_t channels, uint32_t ring_frames,
                          const uint32_t *d_ptr, uint32_t frames, void *stream, const uint32_t *d_src_off)
{
  dim3 grid ((frames + 1 + 127) / 128, channels);
  ring_write_pc_kernel<<<grid, 128, 0, (cudaStream_t) stream>>> (d_blocks, stride, ri, rq, ring_frames, d_ptr, frames, d_src_off);
  return (int) cudaGetLastError ();
}

// per-channel cadence with a super-block chain behind the RX ring: an active channel's block joins ITS super-block at ITS fill level
namespace {
__global__ void acc_store_pc_kernel (uint32_t *__restrict__ acc, uint32_t acc_stride_frames, const uint32_t *__restrict__ blocks, uint32_t frames,
                                     const uint32_t *__restrict__ off, const uint32_t *__restrict__ ptr)
{
  const uint32_t c = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= frames || ptr[4 * (size_t) c + 3] == kRingSkip) return;
  acc[(size_t) c * acc_stride_frames + off[c] + k] = blocks[(size_t) c * frames + k];
}
__global__ void fill_u32_kernel (uint32_t *p, uint32_t n, uint32_t v) { const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
}  // namespace
int launch_acc_store_pc (int16_t *d_acc, uint32_t acc_stride_frames, const int16_t *d_blocks, uint32_t channels, uint32_t frames, const uint32_t *d_off,
                         const uint32_t *d_ptr, void *stream)
{
  dim3 grid ((frames + 127) / 128, channels);
  acc_store_pc_kernel<<<grid, 128, 0, (cudaStream_t) stream>>> (reinterpret_cast<uint32_t *> (d_acc), acc_stride_frames, reinterpret_cast<const uint32_t *> (d_blocks), frames, d_off, d_ptr);
  return (int) cudaGetLastError ();
}
int launch_fill_u32 (uint32_t *d, uint32_t n, uint32_t v, void *stream)
{
  if (n == 0) return 0;
  fill_u32_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>> (d, n, v);
  return (int) cudaGetLastError ();
}
int launch_ring_read_pc (int16_t *d_blocks, const int16_t *ri, const int16_t *rq, uint32_t channels, uint32_t ring_frames,
                         const uint32_t *d_ptr, uint32_t frames, void *stream)
{
  dim3 grid ((frames + 127) / 128, channels);
  ring_read_pc_kernel<<<grid, 128, 0, (cudaStream_t) stream>>> (d_blocks, ri, rq, ring_frames, d_ptr, frames);
  return (int) cudaGetLastError ();
}

// ---- stream feeder: many 1 ms ticks of one ring in ONE launch --------------------------------------
// The pointer evolution of a ring depends on the call sequence only, never on the samples, so the host plans every tick
// (plan[2 t] = first slot written, plan[2 t + 1] = first slot read, both from sl::RingPtrs) and this kernel replays the
// sample movement: one CTA per channel keeps the ring (DSP_BUFF_SIZE frames, one u32 per frame) in shared memory and
// walks the ticks in order. kWriteFirst = true : RX ring, a tick is DSP_In_Buff_Write then DSP_In_Buff_Read
//                          kWriteFirst = false: TX ring, a tick is DSP_Out_Buff_Read (the I2S callback's first statement,
//                                               dsp_if.c:52) then DSP_Out_Buff_Write.
// The block written at tick t comes from `src_a` while t < ticks_a (the chain's carried super-block) and from `src_b`
// afterwards (PASS: ticks_a = 0).
template <bool kWriteFirst>
__global__ void ring_replay_kernel (const uint32_t *__restrict__ src_a, uint32_t stride_a, uint32_t ticks_a, const uint32_t *__restrict__ src_b,
                                    uint32_t stride_b, uint32_t *__restrict__ dst, int16_t *__restrict__ ring_i, int16_t *__restrict__ ring_q,
                                    uint32_t R, const uint32_t *__restrict__ plan, uint32_t ticks, uint32_t B)
{
  extern __shared__ uint32_t sRing[];
  const uint32_t c = blockIdx.x, k = threadIdx.x;
  for (uint32_t i = k; i < R; i += blockDim.x)
    sRing[i] = (uint32_t) (uint16_t) ring_i[(size_t) c * R + i] | ((uint32_t) (uint16_t) ring_q[(size_t) c * R + i] << 16);
  __syncthreads ();
  for (uint32_t t = 0; t < ticks; t++)
  {
    const uint32_t wr0 = plan[2 * t], rd0 = plan[2 * t + 1];
    const uint32_t *src = (t < ticks_a) ? src_a + (size_t) c * stride_a + (size_t) t * B : src_b + (size_t) c * stride_b + (size_t) (t - ticks_a) * B;
    if (kWriteFirst)
    {
      for (uint32_t j = k; j <= B; j += blockDim.x) sRing[(wr0 + j) % R] = src[j < B ? j : B - 1u];   // dsp_if.c:286-300
      __syncthreads ();
      for (uint32_t j = k; j < B; j += blockDim.x) dst[((size_t) c * ticks + t) * B + j] = sRing[(rd0 + j) % R];   // dsp_if.c:328-339
      __syncthreads ();
    }
    else
    {
      for (uint32_t j = k; j < B; j += blockDim.x) dst[((size_t) c * ticks + t) * B + j] = sRing[(rd0 + j) % R];   // dsp_if.c:206-217
      __syncthreads ();
      for (uint32_t j = k; j <= B; j += blockDim.x) sRing[(wr0 + j) % R] = src[j < B ? j : B - 1u];   // dsp_if.c:165-179
      __syncthreads ();
    }
  }
  for (uint32_t i = k; i < R; i += blockDim.x)
  {
    ring_i[(size_t) c * R + i] = (int16_t) (sRing[i] & 0xFFFFu);
    ring_q[(size_t) c * R + i] = (int16_t) (sRing[i] >> 16);
  }
}

int launch_ring_replay (bool write_first, const int16_t *src_a, uint32_t stride_a, uint32_t ticks_a, const int16_t *src_b, uint32_t stride_b,
                        int16_t *dst, int16_t *ri, int16_t *rq, uint32_t channels, uint32_t ring_frames, const uint32_t *d_plan, uint32_t ticks,
                        uint32_t block_frames, void *stream)
{
  const unsigned threads = block_frames + 1 <= 64 ? 64 : (block_frames + 1 <= 128 ? 128 : 256);
  const size_t smem = (size_t) ring_frames * 4;
  const uint32_t *a = reinterpret_cast<const uint32_t *> (src_a), *b = reinterpret_cast<const uint32_t *> (src_b);
  if (write_first)
    ring_replay_kernel<true><<<channels, threads, smem, (cudaStream_t) stream>>> (a, stride_a, ticks_a, b, stride_b, reinterpret_cast<uint32_t *> (dst), ri, rq, ring_frames, d_plan, ticks, block_frames);
  else
    ring_replay_kernel<false><<<channels, threads, smem, (cudaStream_t) stream>>> (a, stride_a, ticks_a, b, stride_b, reinterpret_cast<uint32_t *> (dst), ri, rq, ring_frames, d_plan, ticks, block_frames);
  return (int) cudaGetLastError ();
}

// PASS chain, bulk path: the firmware's steady-state behaviour is the identity on int16 frames.
__global__ void copy_iq_kernel (const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n16, const uint32_t *__restrict__ in_tail,
                                uint32_t *__restrict__ out_tail, size_t ntail)
{
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
  for (size_t k = i; k < n16; k += stride) out[k] = __ldcs (in + k);
  for (size_t k = i; k < ntail; k += stride) out_tail[k] = in_tail[k];
}

int launch_ring_write (const int16_t *d_blocks, uint32_t stride, int16_t *ri, int16_t *rq, uint32_t channels, uint32_t ring_frames,
                       uint32_t wr0, uint32_t frames, void *stream)
{
  dim3 grid ((frames + 1 + 127) / 128, channels);
  ring_write_kernel<<<grid, 128, 0, (cudaStream_t) stream>>> (d_blocks, stride, ri, rq, ring_frames, wr0, frames);
  return (int) cudaGetLastError ();
}
int launch_ring_read (int16_t *d_blocks, const int16_t *ri, const int16_t *rq, uint32_t channels, uint32_t ring_frames, uint32_t rd0,
                      uint32_t frames, void *stream)
{
  dim3 grid ((frames + 127) / 128, channels);
  ring_read_kernel<<<grid, 128, 0, (cudaStream_t) stream>>> (d_blocks, ri, rq, ring_frames, rd0, frames);
  return (int) cudaGetLastError ();
}
int launch_copy_iq (const int16_t *d_in, int16_t *d_out, size_t n_frames, void *stream)
{
  const size_t n16 = n_frames / 4, ntail = n_frames % 4;
  const uint32_t *tin = reinterpret_cast<const uint32_t *> (d_in) + n16 * 4;
  uint32_t *tout = reinterpret_cast<uint32_t *> (d_out) + n16 * 4;
  size_t want = (n16 + 255) / 256; if (want < 1) want = 1;
  int blocks = (int) (want < 148u * 16u ? want : 148u * 16u);
  copy_iq_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>> (reinterpret_cast<const uint4 *> (d_in), reinterpret_cast<uint4 *> (d_out), n16, tin, tout, ntail);
  return (int) cudaGetLastError ();
}


// ---------------------------------------------------------------------------------------------------------
// CW side-tone, mixed where the firmware marks it (dsp_if.c:218, end of DSP_Out_Buff_Read). OUR composition of reference
// stages: arm_sin_f32 (FastMathFunctions/arm_sin_f32.c:72, table + linear interpolation) -> arm_scale_f32 -> arm_float_to_q15
// (truncating, :147) -> arm_add_q15 (saturating, :115) onto L and R. The phase is an integer counter k = (k0 + n f) mod fs, so
// the tone is a pure function of k: one second of it is tabulated once per setting and the mix is one add per word.
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void tone_table_kernel (int16_t *__restrict__ tone, uint32_t fs, float w, float level, const float *__restrict__ tab)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= fs) return;
  // arm_sin_f32.c:72-119 for x = k w >= 0
  const float x = __fmul_rn ((float) k, w);
  float in = __fmul_rn (x, 0.159154943092f);
  const int n = (int) in;
  in = __fsub_rn (in, (float) n);
  const float findex = __fmul_rn (512.0f, in);
  const unsigned idx = ((unsigned) (uint16_t) (int) findex) & 0x1ffu;
  const float fract = __fsub_rn (findex, (float) idx);
  const float sv = __fadd_rn (__fmul_rn (__fsub_rn (1.0f, fract), tab[idx]), __fmul_rn (fract, tab[idx + 1]));
  const float sc = __fmul_rn (sv, level);                                                    // arm_scale_f32.c:77
  const int q = __float2int_rz (__fmul_rn (sc, 32768.0f));                                   // arm_float_to_q15.c:147
  tone[k] = (int16_t) max (-32768, min (32767, q));
}
__global__ void sidetone_mix_kernel (int16_t *__restrict__ blk, uint32_t channels, uint32_t frames, const uint8_t *__restrict__ key,
                                     uint32_t *__restrict__ cnt, const uint32_t *__restrict__ first, const int16_t *__restrict__ tone, uint32_t f, uint32_t fs)
{
  const uint32_t c = blockIdx.x;
  if (first && first[4 * c + 3] == 0xFFFFFFFFu) return;                                      // channel skipped this call (per-channel cadence)
  const uint32_t k0 = cnt[c];
  const bool down = key[c] != 0;
  if (down)
    for (uint32_t n = threadIdx.x; n < frames; n += blockDim.x)
    {
      const int t = tone[(uint32_t) (((uint64_t) k0 + (uint64_t) n * f) % fs)];
      int16_t *p = blk + ((size_t) c * frames + n) * 2;
      p[0] = (int16_t) max (-32768, min (32767, (int) p[0] + t));                             // arm_add_q15.c:115
      p[1] = (int16_t) max (-32768, min (32767, (int) p[1] + t));
    }
  __syncthreads ();
  if (threadIdx.x == 0) cnt[c] = down ? (uint32_t) (((uint64_t) k0 + (uint64_t) frames * f) % fs) : 0u;
}
}  // namespace

int launch_tone_table (int16_t *d_tone, uint32_t fs, float level, const float *d_sin513, void *stream)
{
  const float w = (float) (6.283185307179586 / (double) fs);
  tone_table_kernel<<<(fs + 255) / 256, 256, 0, (cudaStream_t) stream>>> (d_tone, fs, w, level, d_sin513);
  return (int) cudaGetLastError ();
}
int launch_sidetone_mix (int16_t *d_blocks, uint32_t channels, uint32_t frames, const uint8_t *d_key, uint32_t *d_cnt, const uint32_t *d_first,
                         const int16_t *d_tone, uint32_t freq_hz, uint32_t fs, void *stream)
{
  sidetone_mix_kernel<<<channels, 64, 0, (cudaStream_t) stream>>> (d_blocks, channels, frames, d_key, d_cnt, d_first, d_tone, freq_hz, fs);
  return (int) cudaGetLastError ();
}

}  // namespace sl
