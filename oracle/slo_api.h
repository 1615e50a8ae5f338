/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * One C API, two implementations:
 *   - oracle/_ref/libslo_ref.so   (prefix ref_)  : thin glue that CALLS the reference's own vendored
 *       CMSIS-DSP V1.5.3 routines, compiled in place from /root/reference/Drivers/CMSIS/DSP/Source
 *       (see oracle/Makefile). This is "the reference itself, run here".
 *   - oracle/_port/libslo_port.so (prefix port_) : a plain-C restatement of the same routines
 *       (oracle/port/*.c), each function citing the reference file:line it follows. It needs nothing
 *       from /root/reference, so it can be rebuilt on the GPU box.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * either library; the product (selenite_lite_b200/) never does.
 *
 * All stage functions work on ONE channel and take plain arrays. `block` is the blockSize the CMSIS
 * routine is invoked with (the firmware cadence is 48 frames at 48 kHz); `n` must be a
 * multiple of `block`. State arrays follow the CMSIS layouts quoted next to each prototype.
 */
#ifndef SLO_API_H
#define SLO_API_H
#include <stdint.h>

#ifndef SLO_PREFIX
#error "define SLO_PREFIX to ref_ or port_"
#endif
#define SLO_CAT2(a, b) a##b
#define SLO_CAT(a, b) SLO_CAT2 (a, b)
#define SLO(name) SLO_CAT (SLO_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* ---- conversions (SupportFunctions/arm_q15_to_float.c:65, arm_float_to_q15.c:64) ---- */
void SLO (q15_to_float) (const int16_t *src, float *dst, uint32_t n);
void SLO (float_to_q15) (const float *src, int16_t *dst, uint32_t n);

/* ---- FIR (FilteringFunctions/arm_fir_f32.c:553, arm_fir_q15.c:591, arm_fir_fast_q15.c:60, arm_fir_q31.c:60)
 * coeffs are time-reversed {b[T-1]..b[0]}; state has ntaps+block-1 entries (q15: ntaps+block). */
void SLO (fir_f32) (const float *coeffs, uint32_t ntaps, float *state, const float *src, float *dst, uint32_t n, uint32_t block);
void SLO (fir_q15) (const int16_t *coeffs, uint32_t ntaps, int16_t *state, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block);
void SLO (fir_fast_q15) (const int16_t *coeffs, uint32_t ntaps, int16_t *state, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block);
void SLO (fir_q31) (const int32_t *coeffs, uint32_t ntaps, int32_t *state, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block);
/* decimator: state ntaps+block-1, one output per M inputs (arm_fir_decimate_f32.c:129) */
void SLO (fir_decimate_f32) (const float *coeffs, uint32_t ntaps, uint32_t M, float *state, const float *src, float *dst, uint32_t n, uint32_t block);
void SLO (fir_decimate_q15) (const int16_t *coeffs, uint32_t ntaps, uint32_t M, int16_t *state, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block);
/* interpolator: ntaps % L == 0, state block+ntaps/L-1, L outputs per input (arm_fir_interpolate_f32.c:470) */
void SLO (fir_interpolate_f32) (const float *coeffs, uint32_t ntaps, uint32_t L, float *state, const float *src, float *dst, uint32_t n, uint32_t block);
void SLO (fir_interpolate_q15) (const int16_t *coeffs, uint32_t ntaps, uint32_t L, int16_t *state, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block);
/* CW side-tone mixed into one call's worth of DAC frames (the hook of dsp_if.c:218). OUR composition: k = (*counter + n f) mod fs,
 * arm_sin_f32 ((float) k * (float) (2 pi / fs)) -> arm_scale_f32 (level) -> arm_float_to_q15 -> arm_add_q15 onto L and R (saturating);
 * key up: nothing is mixed and the counter returns to 0. lr = int16 [frames][2] in place. */
void SLO (sidetone_mix) (int16_t *lr, uint32_t frames, uint32_t *counter, int key_down, uint32_t freq_hz, uint32_t fs, float level);
/* normalised LMS adaptive FIR (arm_lms_norm_f32.c:161): per sample y = sum state*coeffs (oldest first), e = ref - y,
 * coeffs += (e mu / (energy + 1.19e-7)) * state; energy is the running sum of squares over the tap window. coeffs[ntaps] and
 * en_x0[2] = {energy, x0} are updated in place; state = ntaps-1 previous samples + block (as arm_lms_norm_init_f32 lays it out). */
void SLO (lms_norm_f32) (float *coeffs, uint32_t ntaps, float mu, float *state, float *en_x0, const float *src, const float *ref,
                         float *out, float *err, uint32_t n, uint32_t block);
/* q31 polyphase stages (arm_fir_decimate_q31.c:60, arm_fir_interpolate_q31.c:62): q63 accumulator, result (q31) (acc >> 31) */
void SLO (fir_decimate_q31) (const int32_t *coeffs, uint32_t ntaps, uint32_t M, int32_t *state, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block);
void SLO (fir_interpolate_q31) (const int32_t *coeffs, uint32_t ntaps, uint32_t L, int32_t *state, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block);

/* ---- biquad cascades: coeffs {b0,b1,b2,a1,a2} per stage, feedback sign +a1,+a2
 * (arm_biquad_cascade_df1_f32.c:52). df2T state 2/stage, stereo df2T 4/stage, df1 4/stage.
 * q15 coeffs are {b0,0,b1,b2,a1,a2} (arm_biquad_cascade_df1_q15.c:318). */
void SLO (biquad_df2T_f32) (const float *coeffs, uint32_t nstages, float *state, const float *src, float *dst, uint32_t n, uint32_t block);
void SLO (biquad_stereo_df2T_f32) (const float *coeffs, uint32_t nstages, float *state, const float *src, float *dst, uint32_t nframes, uint32_t block);
void SLO (biquad_df1_f32) (const float *coeffs, uint32_t nstages, float *state, const float *src, float *dst, uint32_t n, uint32_t block);
void SLO (biquad_df1_q15) (const int16_t *coeffs, uint32_t nstages, int32_t postshift, int16_t *state, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block);
void SLO (biquad_df1_q31) (const int32_t *coeffs, uint32_t nstages, int32_t postshift, int32_t *state, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block);

/* ---- transforms: in-place interleaved re/im (arm_cfft_f32.c:562, arm_cfft_q15.c:77, arm_cfft_q31.c:77,
 * arm_rfft_fast_f32.c:288). N in {16..4096}. */
void SLO (cfft_f32) (float *data, uint32_t N, int ifft, int bitrev);
void SLO (cfft_q15) (int16_t *data, uint32_t N, int ifft, int bitrev);
void SLO (cfft_q31) (int32_t *data, uint32_t N, int ifft, int bitrev);
void SLO (rfft_fast_f32) (float *in_destroyed, float *out, uint32_t N, int ifft);

/* ---- complex math (ComplexMathFunctions/arm_cmplx_*.c) ---- */
void SLO (cmplx_mult_cmplx_f32) (const float *a, const float *b, float *dst, uint32_t n);
void SLO (cmplx_mult_real_f32) (const float *a, const float *r, float *dst, uint32_t n);
void SLO (cmplx_conj_f32) (const float *a, float *dst, uint32_t n);
void SLO (cmplx_mag_f32) (const float *a, float *dst, uint32_t n);
void SLO (cmplx_mag_squared_f32) (const float *a, float *dst, uint32_t n);
void SLO (cmplx_mag_q15) (const int16_t *a, int16_t *dst, uint32_t n);

/* ---- statistics (StatisticsFunctions/arm_max_f32.c:58, arm_rms_f32.c:64, arm_power_f32.c:64, ...) ---- */
float SLO (max_f32) (const float *src, uint32_t n, uint32_t *idx);
float SLO (rms_f32) (const float *src, uint32_t n);
float SLO (power_f32) (const float *src, uint32_t n);
float SLO (mean_f32) (const float *src, uint32_t n);
int16_t SLO (max_q15) (const int16_t *src, uint32_t n, uint32_t *idx);
int16_t SLO (rms_q15) (const int16_t *src, uint32_t n);

/* ---- basic math (BasicMathFunctions/arm_scale_f32.c:77, arm_scale_q15.c:56, ...) ---- */
void SLO (scale_f32) (const float *src, float scale, float *dst, uint32_t n);
void SLO (mult_f32) (const float *a, const float *b, float *dst, uint32_t n);
void SLO (add_f32) (const float *a, const float *b, float *dst, uint32_t n);
void SLO (sub_f32) (const float *a, const float *b, float *dst, uint32_t n);
void SLO (abs_f32) (const float *a, float *dst, uint32_t n);
void SLO (scale_q15) (const int16_t *src, int16_t scale_fract, int32_t shift, int16_t *dst, uint32_t n);
void SLO (add_q15) (const int16_t *a, const int16_t *b, int16_t *dst, uint32_t n);
void SLO (sub_q15) (const int16_t *a, const int16_t *b, int16_t *dst, uint32_t n);
void SLO (abs_q15) (const int16_t *a, int16_t *dst, uint32_t n);
void SLO (shift_q15) (const int16_t *a, int32_t shift, int16_t *dst, uint32_t n);

/* ---- NCO (FastMathFunctions/arm_sin_f32.c:72, arm_cos_f32.c) : element-wise over an array ---- */
void SLO (sin_f32) (const float *x, float *dst, uint32_t n);
void SLO (cos_f32) (const float *x, float *dst, uint32_t n);

/* =====================================================================================
 * Chains. The COMPOSITION is ours (the reference has no chain); every box is one of
 * the stage routines above, so ref_ chains are straight-line CMSIS calls.
 * ===================================================================================== */
#define SLO_MAX_STAGES 4
#define SLO_MAX_FFT 4096

/* RX-SSB-f32: q15_to_float -> overlap-save [cfft fwd, cmplx_mult(mask), cfft inv, keep last
 * `hop`, real part] -> biquad_df2T (mono audio) -> per-`agc_block` AGC [abs, max, env/gain recurrence,
 * scale] -> float_to_q15, written L = R. */
typedef struct
{
  uint32_t fft_len;   /* 512 */
  uint32_t hop;       /* 384; overlap = fft_len - hop */
  uint32_t agc_block; /* 48 */
  uint32_t n_stages;  /* 2 */
  float biquad[5 * SLO_MAX_STAGES];
  float agc_target, agc_decay, agc_floor, agc_gmax;
  const float *mask;  /* 2*fft_len, interleaved re/im, unscaled */
  uint32_t envelope;  /* detector: 0: product detector (real part), 1: AM envelope detector [cmplx_mag_f32],
                         2: FM limiter-discriminator [cmplx_conj, cmplx_mult_cmplx, cmplx_mag; see chains.inc.c] */
} slo_rx_f32_params;

typedef struct
{
  int16_t ovl[2 * SLO_MAX_FFT];       /* last (fft_len-hop) raw input frames, interleaved I,Q */
  float bq[2 * SLO_MAX_STAGES];       /* df2T d1,d2 per stage */
  float env;                          /* AGC envelope */
  float zlast[2];                     /* FM: the last filtered baseband sample of the previous super-block (re, im) */
} slo_rx_f32_state;
#define SLO_FM_FLOOR 9.765625e-4f     /* = 2^-10. FM soft squelch: |z[n] conj z[n-1]| below this (|z| < 2^-5 = -30 dBFS: weak or no carrier, filter start-up)
                                         divides by this instead, so the limiter never magnifies float32 rounding of the filter: with the floor at 1e-8 the
                                         reference build and the port disagreed by 85x the 1e-5 bar inside the first super-block, at 2^-10 by 0.23x */

/* frames % hop == 0. audio_dbg (optional) receives the post-biquad, pre-AGC float audio;
 * gain_dbg (optional) the per-agc_block gain. */
void SLO (rx_ssb_f32) (const slo_rx_f32_params *p, slo_rx_f32_state *st, const int16_t *in_iq, int16_t *out_lr,
                       float *audio_dbg, float *gain_dbg, uint32_t frames);
/* C channels, [C][frames][2] layouts, one state per channel, channels split over nthreads pthreads. */
void SLO (rx_ssb_f32_batch) (const slo_rx_f32_params *p, slo_rx_f32_state *st, const int16_t *in_iq, int16_t *out_lr,
                             uint32_t channels, uint32_t frames, uint32_t nthreads);

/* TX-SSB-f32: mic = L of each L=R frame -> q15_to_float -> overlap-save [cfft fwd of (mic, 0),
 * cmplx_mult(mask), cfft inv, keep last `hop`] -> per-`alc_block` ALC [cmplx_mag, max, env/gain recurrence, scale]
 * -> float_to_q15, written interleaved I,Q. The one-sided mask is band-pass and Hilbert pair in one. */
typedef struct
{
  uint32_t fft_len, hop, alc_block;
  float alc_target, alc_decay, alc_floor, alc_gmax;
  const float *mask;
} slo_tx_f32_params;

typedef struct
{
  int16_t ovl[SLO_MAX_FFT];           /* last (fft_len-hop) mic samples */
  float env;
} slo_tx_f32_state;

/* frames % hop == 0. iq_dbg (optional) receives the pre-ALC complex baseband [frames][2]; gain_dbg the per-block gain. */
void SLO (tx_ssb_f32) (const slo_tx_f32_params *p, slo_tx_f32_state *st, const int16_t *in_lr, int16_t *out_iq,
                       float *iq_dbg, float *gain_dbg, uint32_t frames);

/* CHAN-64-f32 (BASELINE config 4): one wideband stream -> q15_to_float -> `bins`-branch polyphase
 * filter [branch r = fir_f32 with taps e_r[p] = proto[bins*p + bins-1-r] on the commutated input x[bins*m + r], I and Q
 * separately] -> cfft_f32(len bins, forward) per hop -> per bin: real part (or cmplx_mag when `envelope`) ->
 * per-`agc_block` AGC [abs, max, gain law, scale] -> float_to_q15, written channel-major [bin][hop] L = R. */
#define SLO_CHAN_MAX_BINS 64
#define SLO_CHAN_MAX_TAPS 8
typedef struct
{
  uint32_t bins, taps_per_branch, agc_block, envelope;
  float agc_target, agc_decay, agc_floor, agc_gmax;
  const float *proto;                                   /* bins * taps_per_branch */
} slo_chan_params;

typedef struct
{
  float fir_i[SLO_CHAN_MAX_BINS][SLO_CHAN_MAX_TAPS];    /* arm_fir_f32 history per branch (numTaps-1 used) */
  float fir_q[SLO_CHAN_MAX_BINS][SLO_CHAN_MAX_TAPS];
  float env[SLO_CHAN_MAX_BINS];
} slo_chan_state;

/* frames % (bins * agc_block) == 0. out_lr: [bins][frames/bins][2]; audio_dbg (optional): [bins][frames/bins];
 * gain_dbg (optional): [bins][frames/bins/agc_block]. */
void SLO (chan_f32) (const slo_chan_params *p, slo_chan_state *st, const int16_t *in_iq, int16_t *out_lr,
                     float *audio_dbg, float *gain_dbg, uint32_t frames);

/* RX-SSB-q15 (SURVEY.md Appendix B "the fully-integer variant"): phasing-method SSB demodulator, every
 * box an integer CMSIS routine, so the GPU must match bit for bit. Per `agc_block` (the firmware block):
 * de-interleave I, Q -> fir_q15(I, taps_i), fir_q15(Q, taps_q) [a +-45 degree band-pass Hilbert pair]
 * -> add_q15 (USB) or sub_q15 (LSB), saturating -> abs_q15 + max_q15 = block peak p_b
 * -> AGC law (ours): env_b = max(p_b, max_{j=1..agc_window-1} (p_{b-j} * rel[j]) >> 15)   [finite release window]
 *    q = min((agc_target << 15) / max(env_b, agc_floor), agc_gmax_q15)                     [gain in Q15, integer divide]
 *    shift s = smallest s >= 0 with (q >> s) <= 32767, scaleFract = q >> s
 * -> scale_q15(audio, scaleFract, s) -> written L = R. */
#define SLO_Q15_TAPS 64
#define SLO_Q15_WIN 32
#define SLO_Q15_MAX_BLOCK 192
typedef struct
{
  uint32_t ntaps;                                        /* even (arm_fir_init_q15.c:88-128), <= SLO_Q15_TAPS */
  uint32_t agc_block;                                    /* 48 */
  uint32_t agc_window;                                   /* blocks of peak history the envelope sees, 1..SLO_Q15_WIN */
  uint32_t lsb;                                          /* 0: I' + Q' (upper sideband), 1: I' - Q' */
  int16_t taps_i[SLO_Q15_TAPS], taps_q[SLO_Q15_TAPS];    /* b[0..ntaps-1] in natural order */
  int16_t rel[SLO_Q15_WIN];                              /* q15 release weight by block age; rel[0] unused */
  int16_t agc_target, agc_floor;                         /* q15; floor >= 1 */
  uint32_t agc_gmax_q15;                                 /* gain limit in Q15 (1.0 = 32768) */
  /* optional audio filter between the mixer and the AGC detector: arm_biquad_cascade_df1_q15 on the mixed audio,
   * bq_stages = 0 switches it off; coefficients {b0, 0, b1, b2, a1, a2} per stage as arm_biquad_cascade_df1_init_q15 lays them out */
  uint32_t bq_stages; int32_t bq_postshift; int16_t bq_coeffs[6 * 4];
} slo_rx_q15_params;

typedef struct
{
  int16_t fir_i[SLO_Q15_TAPS + SLO_Q15_MAX_BLOCK], fir_q[SLO_Q15_TAPS + SLO_Q15_MAX_BLOCK];   /* arm_fir_q15 pState */
  int16_t peaks[SLO_Q15_WIN];                            /* peaks[j] = block peak j+1 blocks ago */
  int16_t bq[4 * 4];                                     /* arm_biquad_cascade_df1_q15 pState: {x[n-1], x[n-2], y[n-1], y[n-2]} per stage */
} slo_rx_q15_state;

/* frames % agc_block == 0. audio_dbg (optional): the pre-AGC q15 audio [frames]; gain_dbg (optional): q per block. */
void SLO (rx_ssb_q15) (const slo_rx_q15_params *p, slo_rx_q15_state *st, const int16_t *in_iq, int16_t *out_lr,
                       int16_t *audio_dbg, uint32_t *gain_dbg, uint32_t frames);

#ifdef __cplusplus
}
#endif
#endif /* SLO_API_H */
