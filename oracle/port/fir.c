/* ORACLE / TEST INFRASTRUCTURE ONLY. FIR family. State layout (all variants): state[0..ntaps-2] holds the previous
 * samples oldest first, the new block is appended, and after the block the last ntaps-1 samples move to the front
 * (Source/FilteringFunctions/arm_fir_f32.c:54-70 doc, copy-back :947 ff.). coeffs[k] = b[ntaps-1-k]. */
#include "port_common.h"

/* arm_fir_f32.c:553 ff. (CM3/CM4 branch): acc over k = 0..ntaps-1 of state[n+k]*coeffs[k], oldest first. */
void port_fir_f32 (const float *c, uint32_t nt, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (float) * block);
    for (uint32_t i = 0; i < block; i++)
    {
      float acc = 0.0f;
      for (uint32_t k = 0; k < nt; k++) acc += st[i + k] * c[k];
      dst[o + i] = acc;
    }
    memmove (st, st + block, sizeof (float) * (nt - 1));
  }
}

/* arm_fir_q15.c:591-642 : q63 accumulator, >>15, truncated to q31 by the __SSAT prototype, saturate 16. */
void port_fir_q15 (const int16_t *c, uint32_t nt, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (int16_t) * block);
    for (uint32_t i = 0; i < block; i++)
    {
      int64_t acc = 0;
      for (uint32_t k = 0; k < nt; k++) acc += (int32_t) st[i + k] * c[k];
      dst[o + i] = (int16_t) slo_ssat16 ((int32_t) (acc >> 15));
    }
    memmove (st, st + block, sizeof (int16_t) * (nt - 1));
  }
}

/* arm_fir_fast_q15.c:60 ff. : 32-bit wrapping accumulator (no guard bits), >>15, saturate 16. */
void port_fir_fast_q15 (const int16_t *c, uint32_t nt, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (int16_t) * block);
    for (uint32_t i = 0; i < block; i++)
    {
      uint32_t acc = 0;
      for (uint32_t k = 0; k < nt; k++) acc += (uint32_t) ((int32_t) st[i + k] * c[k]);
      dst[o + i] = (int16_t) slo_ssat16 ((int32_t) acc >> 15);
    }
    memmove (st, st + block, sizeof (int16_t) * (nt - 1));
  }
}

/* arm_fir_q31.c:60 ff. : q63 accumulator of q31*q31, result (q31)(acc >> 31) (no saturation, doc :50-56). */
void port_fir_q31 (const int32_t *c, uint32_t nt, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (int32_t) * block);
    for (uint32_t i = 0; i < block; i++)
    {
      int64_t acc = 0;
      for (uint32_t k = 0; k < nt; k++) acc += (int64_t) st[i + k] * c[k];
      dst[o + i] = (int32_t) (acc >> 31);
    }
    memmove (st, st + block, sizeof (int32_t) * (nt - 1));
  }
}

/* arm_fir_decimate_f32.c:129 ff. (generic loop :428-506): M new samples enter the state per output; the output is
 * the FIR evaluated at every M-th position. */
void port_fir_decimate_f32 (const float *c, uint32_t nt, uint32_t M, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (float) * block);
    for (uint32_t i = 0; i < block / M; i++)
    {
      float acc = 0.0f;
      for (uint32_t k = 0; k < nt; k++) acc += st[i * M + k] * c[k];
      dst[o / M + i] = acc;
    }
    memmove (st, st + block, sizeof (float) * (nt - 1));
  }
}
/* arm_fir_decimate_q15.c (plain-C branch): q63 acc, >>15, sat16 */
void port_fir_decimate_q15 (const int16_t *c, uint32_t nt, uint32_t M, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (int16_t) * block);
    for (uint32_t i = 0; i < block / M; i++)
    {
      int64_t acc = 0;
      for (uint32_t k = 0; k < nt; k++) acc += (int32_t) st[i * M + k] * c[k];
      dst[o / M + i] = (int16_t) slo_ssat16 ((int32_t) (acc >> 15));
    }
    memmove (st, st + block, sizeof (int16_t) * (nt - 1));
  }
}

/* arm_fir_interpolate_f32.c:470-563 : phaseLength = ntaps/L; state holds phaseLength-1 old inputs + block.
 * For each input sample, outputs j = 0..L-1 use taps coeffs[(L-1-j) ... step L] — the loop walks
 * pCoeffs + (i-1) for i = L..1 (:508-526), i.e. output j=0 uses offset L-1, against the state oldest first. */
void port_fir_interpolate_f32 (const float *c, uint32_t nt, uint32_t L, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  uint32_t P = nt / L;
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + P - 1, src + o, sizeof (float) * block);
    for (uint32_t i = 0; i < block; i++)
      for (uint32_t j = 0; j < L; j++)
      {
        float acc = 0.0f;
        for (uint32_t k = 0; k < P; k++) acc += st[i + k] * c[(L - 1 - j) + k * L];
        dst[(size_t) (o + i) * L + j] = acc;
      }
    memmove (st, st + block, sizeof (float) * (P - 1));
  }
}
/* arm_fir_interpolate_q15.c (plain-C branch): q63 acc, >>15, sat16 */
void port_fir_interpolate_q15 (const int16_t *c, uint32_t nt, uint32_t L, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  uint32_t P = nt / L;
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + P - 1, src + o, sizeof (int16_t) * block);
    for (uint32_t i = 0; i < block; i++)
      for (uint32_t j = 0; j < L; j++)
      {
        int64_t acc = 0;
        for (uint32_t k = 0; k < P; k++) acc += (int32_t) st[i + k] * c[(L - 1 - j) + k * L];
        dst[(size_t) (o + i) * L + j] = (int16_t) slo_ssat16 ((int32_t) (acc >> 15));
      }
    memmove (st, st + block, sizeof (int16_t) * (P - 1));
  }
}

/* arm_fir_decimate_q31.c:60 ff. (generic loop :216-293; the ARM_MATH_DSP branch :77-215 accumulates in the same order): q63
 * accumulator of q31 x q31 products, one output per M inputs, result (q31) (acc >> 31) (:178 / :269), no saturation. */
void port_fir_decimate_q31 (const int32_t *c, uint32_t nt, uint32_t M, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + nt - 1, src + o, sizeof (int32_t) * block);
    for (uint32_t i = 0; i < block / M; i++)
    {
      int64_t acc = 0;
      for (uint32_t k = 0; k < nt; k++) acc += (int64_t) st[i * M + k] * c[k];
      dst[o / M + i] = (int32_t) (acc >> 31);
    }
    memmove (st, st + block, sizeof (int32_t) * (nt - 1));
  }
}
/* arm_fir_interpolate_q31.c:62 ff. (generic loop :385-488): phaseLength = ntaps / L, output j of an input uses the coefficients
 * (L - 1 - j) + k L against the state oldest first (:423-445), q63 accumulator, result (q31) (acc >> 31) (:454). */
void port_fir_interpolate_q31 (const int32_t *c, uint32_t nt, uint32_t L, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  uint32_t P = nt / L;
  for (uint32_t o = 0; o < n; o += block)
  {
    memcpy (st + P - 1, src + o, sizeof (int32_t) * block);
    for (uint32_t i = 0; i < block; i++)
      for (uint32_t j = 0; j < L; j++)
      {
        int64_t acc = 0;
        for (uint32_t k = 0; k < P; k++) acc += (int64_t) st[i + k] * c[(L - 1 - j) + k * L];
        dst[(size_t) (o + i) * L + j] = (int32_t) (acc >> 31);
      }
    memmove (st, st + block, sizeof (int32_t) * (P - 1));
  }
}

/* arm_lms_norm_f32.c:161 ff. (generic loop :337-404; the ARM_MATH_DSP branch :200-330 unrolls the same sequential sums by four):
 * the new sample enters the state, the window energy drops the sample that left (x0) and takes the new one (:349-350), the
 * filter output is the oldest-first sum (:359-366), e = ref - y (:372-374), w = e mu / (energy + epsilon) (:378),
 * every coefficient takes w * its state sample (:389-398), x0 = the oldest sample of this window (:400). */
void port_lms_norm_f32 (float *coeffs, uint32_t nt, float mu, float *st, float *en_x0, const float *src, const float *ref,
                         float *out, float *err, uint32_t n, uint32_t block)
{
  float energy = en_x0[0], x0 = en_x0[1];
  for (uint32_t o = 0; o < n; o += block)
  {
    for (uint32_t i = 0; i < block; i++)
    {
      const float in = src[o + i];
      st[nt - 1 + i] = in;
      energy -= x0 * x0;
      energy += in * in;
      float sum = 0.0f;
      for (uint32_t k = 0; k < nt; k++) sum += st[i + k] * coeffs[k];
      out[o + i] = sum;
      const float e = ref[o + i] - sum;
      err[o + i] = e;
      const float w = (e * mu) / (energy + 0.000000119209289f);
      for (uint32_t k = 0; k < nt; k++) coeffs[k] += w * st[i + k];
      x0 = st[i];
    }
    memmove (st, st + block, sizeof (float) * (nt - 1));
  }
  en_x0[0] = energy; en_x0[1] = x0;
}
