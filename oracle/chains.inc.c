/* ORACLE / TEST INFRASTRUCTURE ONLY — never on the product path.
 *
 * Chain glue shared by both oracle libraries: compiled once with SLO_PREFIX=ref_ (every SLO(stage) call lands in
 * the reference's CMSIS-DSP routine) and once with SLO_PREFIX=port_ (the plain-C restatement).
 * THE COMPOSITION IS OURS: the reference firmware copies samples and contains no chain (dsp_if.c:367
 * DSP_Set_Mode is empty). What the reference pins is each stage's arithmetic; the chain spec lives in DESIGN.md §3.
 */
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include "slo_api.h"

/* RX-SSB-f32: see slo_api.h. Stage by stage this is what a firmware author would write inside
 * DSP_In_Buff_Write (dsp_if.c:286-289) with arm_math.h. */
void SLO (rx_ssb_f32) (const slo_rx_f32_params *p, slo_rx_f32_state *st, const int16_t *in_iq, int16_t *out_lr,
                       float *audio_dbg, float *gain_dbg, uint32_t frames)
{
  const uint32_t N = p->fft_len, hop = p->hop, ovl = N - hop, B = p->agc_block;
  int16_t *raw = (int16_t *) malloc (sizeof (int16_t) * 2 * N);
  int16_t *mono = (int16_t *) malloc (sizeof (int16_t) * hop);
  float *frame = (float *) malloc (sizeof (float) * 2 * N);
  float *prod = (float *) malloc (sizeof (float) * 2 * N);
  float *audio = (float *) malloc (sizeof (float) * hop);
  float *scaled = (float *) malloc (sizeof (float) * hop);
  float *absb = (float *) malloc (sizeof (float) * B);
  float *prev = (float *) malloc (sizeof (float) * 2 * hop);
  float *w = (float *) malloc (sizeof (float) * 2 * hop);

  for (uint32_t o = 0; o < frames; o += hop)
  {
    /* overlap-save frame = [previous ovl frames | hop new frames] */
    memcpy (raw, st->ovl, sizeof (int16_t) * 2 * ovl);
    memcpy (raw + 2 * ovl, in_iq + 2 * (size_t) o, sizeof (int16_t) * 2 * hop);
    memcpy (st->ovl, raw + 2 * hop, sizeof (int16_t) * 2 * ovl);

    SLO (q15_to_float) (raw, frame, 2 * N);
    SLO (cfft_f32) (frame, N, 0, 1);
    SLO (cmplx_mult_cmplx_f32) (frame, p->mask, prod, N);
    SLO (cfft_f32) (prod, N, 1, 1);
    if (p->envelope == 1) SLO (cmplx_mag_f32) (prod + 2 * ovl, audio, hop);   /* AM: keep last hop, envelope */
    else if (p->envelope == 2)
    {
      /* FM (mode byte 0x08, rxtx_if.h:40). CMSIS-DSP V1.5.3 has no arctangent, so the discriminator is the limiter +
       * quadrature form that needs none: w[n] = z[n] conj (z[n-1]) has the phase step of the carrier as its angle, and
       * d[n] = Im w[n] / |w[n]| = sin (phase step) — the amplitude is divided out (the limiter), the sine is within 2 % of
       * the angle up to +-2.7 kHz of deviation at 48 kHz and monotonic up to +-12 kHz. OURS, like every composition here. */
      const float *z = prod + 2 * ovl;                                   /* keep last hop, both rails */
      prev[0] = st->zlast[0]; prev[1] = st->zlast[1];
      memcpy (prev + 2, z, sizeof (float) * 2 * (hop - 1));              /* z delayed by one sample */
      st->zlast[0] = z[2 * (hop - 1)]; st->zlast[1] = z[2 * (hop - 1) + 1];
      SLO (cmplx_conj_f32) (prev, prev, hop);
      SLO (cmplx_mult_cmplx_f32) (z, prev, w, hop);
      SLO (cmplx_mag_f32) (w, audio, hop);
      for (uint32_t k = 0; k < hop; k++) audio[k] = w[2 * k + 1] / (audio[k] > SLO_FM_FLOOR ? audio[k] : SLO_FM_FLOOR);
    }
    else for (uint32_t k = 0; k < hop; k++) audio[k] = prod[2 * (ovl + k)];   /* keep last hop, real part */

    SLO (biquad_df2T_f32) (p->biquad, p->n_stages, st->bq, audio, audio, hop, B);
    if (audio_dbg) memcpy (audio_dbg + o, audio, sizeof (float) * hop);

    for (uint32_t b = 0; b < hop; b += B)
    {
      SLO (abs_f32) (audio + b, absb, B);
      float peak = SLO (max_f32) (absb, B, 0);
      /* AGC law (ours): instant attack, exponential release, bounded gain */
      float rel = st->env * p->agc_decay;
      float env = peak > rel ? peak : rel;
      float den = env > p->agc_floor ? env : p->agc_floor;
      float g = p->agc_target / den;
      if (g > p->agc_gmax) g = p->agc_gmax;
      st->env = env;
      if (gain_dbg) gain_dbg[(o + b) / B] = g;
      SLO (scale_f32) (audio + b, g, scaled + b, B);
    }
    SLO (float_to_q15) (scaled, mono, hop);
    for (uint32_t k = 0; k < hop; k++)                                   /* USB IN endpoint is stereo: L = R */
    {
      out_lr[2 * (size_t) (o + k)] = mono[k];
      out_lr[2 * (size_t) (o + k) + 1] = mono[k];
    }
  }
  free (raw); free (mono); free (frame); free (prod); free (audio); free (scaled); free (absb); free (prev); free (w);
}

typedef struct
{
  const slo_rx_f32_params *p; slo_rx_f32_state *st; const int16_t *in; int16_t *out;
  uint32_t c0, c1, frames;
} slo_rx_job;

static void *slo_rx_worker (void *arg)
{
  slo_rx_job *j = (slo_rx_job *) arg;
  for (uint32_t c = j->c0; c < j->c1; c++)
    SLO (rx_ssb_f32) (j->p, j->st + c, j->in + 2 * (size_t) c * j->frames, j->out + 2 * (size_t) c * j->frames, 0, 0, j->frames);
  return 0;
}

void SLO (rx_ssb_f32_batch) (const slo_rx_f32_params *p, slo_rx_f32_state *st, const int16_t *in_iq, int16_t *out_lr,
                             uint32_t channels, uint32_t frames, uint32_t nthreads)
{
  if (nthreads < 1) nthreads = 1;
  if (nthreads > channels) nthreads = channels;
  pthread_t *th = (pthread_t *) malloc (sizeof (pthread_t) * nthreads);
  slo_rx_job *jobs = (slo_rx_job *) malloc (sizeof (slo_rx_job) * nthreads);
  for (uint32_t t = 0; t < nthreads; t++)
  {
    jobs[t].p = p; jobs[t].st = st; jobs[t].in = in_iq; jobs[t].out = out_lr; jobs[t].frames = frames;
    jobs[t].c0 = (uint32_t) ((uint64_t) channels * t / nthreads);
    jobs[t].c1 = (uint32_t) ((uint64_t) channels * (t + 1) / nthreads);
    pthread_create (&th[t], 0, slo_rx_worker, &jobs[t]);
  }
  for (uint32_t t = 0; t < nthreads; t++) pthread_join (th[t], 0);
  free (th); free (jobs);
}

/* TX-SSB-f32: see slo_api.h. What a firmware author would write for the transmit direction with arm_math.h: the
 * codec delivers the microphone on both ADC channels (codec_if.c:304-306), the modulated I/Q goes to the DAC. */
void SLO (tx_ssb_f32) (const slo_tx_f32_params *p, slo_tx_f32_state *st, const int16_t *in_lr, int16_t *out_iq,
                       float *iq_dbg, float *gain_dbg, uint32_t frames)
{
  const uint32_t N = p->fft_len, hop = p->hop, ovl = N - hop, B = p->alc_block;
  int16_t *raw = (int16_t *) malloc (sizeof (int16_t) * N);
  float *mic = (float *) malloc (sizeof (float) * N);
  float *frame = (float *) malloc (sizeof (float) * 2 * N);
  float *prod = (float *) malloc (sizeof (float) * 2 * N);
  float *scaled = (float *) malloc (sizeof (float) * 2 * hop);
  float *mag = (float *) malloc (sizeof (float) * B);

  for (uint32_t o = 0; o < frames; o += hop)
  {
    memcpy (raw, st->ovl, sizeof (int16_t) * ovl);
    for (uint32_t k = 0; k < hop; k++) raw[ovl + k] = in_lr[2 * (size_t) (o + k)];   /* L channel */
    memcpy (st->ovl, raw + hop, sizeof (int16_t) * ovl);

    SLO (q15_to_float) (raw, mic, N);
    for (uint32_t k = 0; k < N; k++) { frame[2 * k] = mic[k]; frame[2 * k + 1] = 0.0f; }
    SLO (cfft_f32) (frame, N, 0, 1);
    SLO (cmplx_mult_cmplx_f32) (frame, p->mask, prod, N);
    SLO (cfft_f32) (prod, N, 1, 1);
    const float *iq = prod + 2 * ovl;                                      /* keep last hop */
    if (iq_dbg) memcpy (iq_dbg + 2 * (size_t) o, iq, sizeof (float) * 2 * hop);

    for (uint32_t b = 0; b < hop; b += B)
    {
      SLO (cmplx_mag_f32) (iq + 2 * b, mag, B);
      float peak = SLO (max_f32) (mag, B, 0);
      /* same gain law as the RX AGC (ours) */
      float rel = st->env * p->alc_decay;
      float env = peak > rel ? peak : rel;
      float den = env > p->alc_floor ? env : p->alc_floor;
      float g = p->alc_target / den;
      if (g > p->alc_gmax) g = p->alc_gmax;
      st->env = env;
      if (gain_dbg) gain_dbg[(o + b) / B] = g;
      SLO (scale_f32) (iq + 2 * b, g, scaled + 2 * b, 2 * B);
    }
    SLO (float_to_q15) (scaled, out_iq + 2 * (size_t) o, 2 * hop);
  }
  free (raw); free (mic); free (frame); free (prod); free (scaled); free (mag);
}

/* CHAN-64-f32: see slo_api.h. Straight-line CMSIS calls: one arm_fir_f32 instance per branch and rail over the whole
 * call, one arm_cfft_f32 per hop, then the per-bin detector and AGC at the firmware's 1 ms cadence. */
void SLO (chan_f32) (const slo_chan_params *p, slo_chan_state *st, const int16_t *in_iq, int16_t *out_lr,
                     float *audio_dbg, float *gain_dbg, uint32_t frames)
{
  const uint32_t M = p->bins, P = p->taps_per_branch, B = p->agc_block, hops = frames / M;
  float *xf = (float *) malloc (sizeof (float) * 2 * (size_t) frames);
  float *bi = (float *) malloc (sizeof (float) * hops), *bq = (float *) malloc (sizeof (float) * hops);
  float *vi = (float *) malloc (sizeof (float) * (size_t) M * hops), *vq = (float *) malloc (sizeof (float) * (size_t) M * hops);
  float *fstate = (float *) malloc (sizeof (float) * (P + hops));
  float *coef = (float *) malloc (sizeof (float) * P);
  float *spec = (float *) malloc (sizeof (float) * 2 * M), *magb = (float *) malloc (sizeof (float) * M);
  float *audio = (float *) malloc (sizeof (float) * (size_t) M * hops);
  float *absb = (float *) malloc (sizeof (float) * B), *scaled = (float *) malloc (sizeof (float) * B);
  int16_t *mono = (int16_t *) malloc (sizeof (int16_t) * B);

  SLO (q15_to_float) (in_iq, xf, 2 * frames);
  for (uint32_t r = 0; r < M; r++)
  {
    /* branch taps e_r[p] = h[M p + M-1-r], handed to CMSIS time-reversed (arm_fir_f32.c:54-66) */
    for (uint32_t k = 0; k < P; k++) coef[k] = p->proto[M * (P - 1 - k) + (M - 1 - r)];
    for (uint32_t m = 0; m < hops; m++) { bi[m] = xf[2 * ((size_t) M * m + r)]; bq[m] = xf[2 * ((size_t) M * m + r) + 1]; }
    memset (fstate, 0, sizeof (float) * (P + hops)); memcpy (fstate, st->fir_i[r], sizeof (float) * (P - 1));
    SLO (fir_f32) (coef, P, fstate, bi, vi + (size_t) r * hops, hops, B);
    memcpy (st->fir_i[r], fstate, sizeof (float) * (P - 1));
    memset (fstate, 0, sizeof (float) * (P + hops)); memcpy (fstate, st->fir_q[r], sizeof (float) * (P - 1));
    SLO (fir_f32) (coef, P, fstate, bq, vq + (size_t) r * hops, hops, B);
    memcpy (st->fir_q[r], fstate, sizeof (float) * (P - 1));
  }
  for (uint32_t m = 0; m < hops; m++)
  {
    for (uint32_t r = 0; r < M; r++) { spec[2 * r] = vi[(size_t) r * hops + m]; spec[2 * r + 1] = vq[(size_t) r * hops + m]; }
    SLO (cfft_f32) (spec, M, 0, 1);
    if (p->envelope) { SLO (cmplx_mag_f32) (spec, magb, M); for (uint32_t k = 0; k < M; k++) audio[(size_t) k * hops + m] = magb[k]; }
    else for (uint32_t k = 0; k < M; k++) audio[(size_t) k * hops + m] = spec[2 * k];
  }
  if (audio_dbg) memcpy (audio_dbg, audio, sizeof (float) * (size_t) M * hops);
  for (uint32_t k = 0; k < M; k++)
    for (uint32_t b = 0; b < hops; b += B)
    {
      const float *a = audio + (size_t) k * hops + b;
      SLO (abs_f32) (a, absb, B);
      float peak = SLO (max_f32) (absb, B, 0);
      float rel = st->env[k] * p->agc_decay;
      float env = peak > rel ? peak : rel;
      float den = env > p->agc_floor ? env : p->agc_floor;
      float g = p->agc_target / den;
      if (g > p->agc_gmax) g = p->agc_gmax;
      st->env[k] = env;
      if (gain_dbg) gain_dbg[(size_t) k * (hops / B) + b / B] = g;
      SLO (scale_f32) (a, g, scaled, B);
      SLO (float_to_q15) (scaled, mono, B);
      for (uint32_t i = 0; i < B; i++) { out_lr[2 * ((size_t) k * hops + b + i)] = mono[i]; out_lr[2 * ((size_t) k * hops + b + i) + 1] = mono[i]; }
    }
  free (xf); free (bi); free (bq); free (vi); free (vq); free (fstate); free (coef); free (spec); free (magb); free (audio); free (absb); free (scaled); free (mono);
}

/* RX-SSB-q15: see slo_api.h. All-integer phasing demodulator, one firmware block per iteration — what a firmware
 * author would write inside DSP_In_Buff_Write (dsp_if.c:286-289) with the q15 half of arm_math.h. */
void SLO (rx_ssb_q15) (const slo_rx_q15_params *p, slo_rx_q15_state *st, const int16_t *in_iq, int16_t *out_lr,
                       int16_t *audio_dbg, uint32_t *gain_dbg, uint32_t frames)
{
  const uint32_t T = p->ntaps, B = p->agc_block, K = p->agc_window;
  int16_t ci[SLO_Q15_TAPS], cq[SLO_Q15_TAPS];
  int16_t xi[SLO_Q15_MAX_BLOCK], xq[SLO_Q15_MAX_BLOCK], fi[SLO_Q15_MAX_BLOCK], fq[SLO_Q15_MAX_BLOCK];
  int16_t a[SLO_Q15_MAX_BLOCK], ab[SLO_Q15_MAX_BLOCK], y[SLO_Q15_MAX_BLOCK];
  for (uint32_t k = 0; k < T; k++) { ci[k] = p->taps_i[T - 1 - k]; cq[k] = p->taps_q[T - 1 - k]; }   /* arm_fir_q15.c: pCoeffs time-reversed */

  for (uint32_t o = 0; o < frames; o += B)
  {
    for (uint32_t k = 0; k < B; k++) { xi[k] = in_iq[2 * (size_t) (o + k)]; xq[k] = in_iq[2 * (size_t) (o + k) + 1]; }   /* dsp_if.c:227-238 */
    SLO (fir_q15) (ci, T, st->fir_i, xi, fi, B, B);
    SLO (fir_q15) (cq, T, st->fir_q, xq, fq, B, B);
    if (p->lsb) SLO (sub_q15) (fi, fq, a, B); else SLO (add_q15) (fi, fq, a, B);
    if (p->bq_stages) SLO (biquad_df1_q15) (p->bq_coeffs, p->bq_stages, p->bq_postshift, st->bq, a, a, B, B);   /* optional audio filter (Appendix B) */
    if (audio_dbg) memcpy (audio_dbg + o, a, sizeof (int16_t) * B);
    SLO (abs_q15) (a, ab, B);
    const int32_t peak = SLO (max_q15) (ab, B, 0);
    /* AGC law (ours): finite-window peak hold with a q15 release table */
    int32_t env = peak;
    for (uint32_t j = 1; j < K; j++)
    {
      const int32_t v = ((int32_t) st->peaks[j - 1] * (int32_t) p->rel[j]) >> 15;
      if (v > env) env = v;
    }
    for (uint32_t j = SLO_Q15_WIN - 1; j > 0; j--) st->peaks[j] = st->peaks[j - 1];
    st->peaks[0] = (int16_t) peak;
    const uint32_t den = (uint32_t) (env > p->agc_floor ? env : p->agc_floor);
    uint32_t q = ((uint32_t) p->agc_target << 15) / den;
    if (q > p->agc_gmax_q15) q = p->agc_gmax_q15;
    uint32_t s = 0;
    while ((q >> s) > 32767u) s++;
    if (gain_dbg) gain_dbg[o / B] = q;
    SLO (scale_q15) (a, (int16_t) (q >> s), (int32_t) s, y, B);
    for (uint32_t k = 0; k < B; k++) { out_lr[2 * (size_t) (o + k)] = y[k]; out_lr[2 * (size_t) (o + k) + 1] = y[k]; }   /* L = R */
  }
}


/* CW side-tone at the firmware's hook (dsp_if.c:218 "mix CW tone to speaker signal here"): see slo_api.h. */
void SLO (sidetone_mix) (int16_t *lr, uint32_t frames, uint32_t *counter, int key_down, uint32_t freq_hz, uint32_t fs, float level)
{
  if (!key_down) { *counter = 0; return; }
  const float w = (float) (6.283185307179586 / (double) fs);
  float *x = (float *) malloc (sizeof (float) * frames), *s = (float *) malloc (sizeof (float) * frames), *sc = (float *) malloc (sizeof (float) * frames);
  int16_t *t = (int16_t *) malloc (sizeof (int16_t) * frames), *l = (int16_t *) malloc (sizeof (int16_t) * frames), *r = (int16_t *) malloc (sizeof (int16_t) * frames);
  for (uint32_t n = 0; n < frames; n++)
  {
    const uint32_t k = (uint32_t) (((uint64_t) *counter + (uint64_t) n * freq_hz) % fs);
    x[n] = (float) k * w;
    l[n] = lr[2 * n]; r[n] = lr[2 * n + 1];
  }
  SLO (sin_f32) (x, s, frames);
  SLO (scale_f32) (s, level, sc, frames);
  SLO (float_to_q15) (sc, t, frames);
  SLO (add_q15) (l, t, l, frames);
  SLO (add_q15) (r, t, r, frames);
  for (uint32_t n = 0; n < frames; n++) { lr[2 * n] = l[n]; lr[2 * n + 1] = r[n]; }
  *counter = (uint32_t) (((uint64_t) *counter + (uint64_t) frames * freq_hz) % fs);
  free (x); free (s); free (sc); free (t); free (l); free (r);
}
