/* ORACLE / TEST INFRASTRUCTURE ONLY. Biquad cascades. All are stage-major (each stage filters the whole block before
 * the next stage starts, e.g. Source/FilteringFunctions/arm_biquad_cascade_df1_f32.c:393-399); since every stage is a
 * causal sample-sequential filter the results do not depend on the blocking. Feedback sign is +a1,+a2
 * (arm_biquad_cascade_df1_f32.c:52,:58-63). */
#include "port_common.h"

/* arm_biquad_cascade_df2T_f32.c:459-583 (CM3/CM4 branch), exact association :551-562:
 *   y = b0*x + d1;  d1 = (b1*x + a1*y) + d2;  d2 = b2*x + a2*y */
void port_biquad_df2T_f32 (const float *c, uint32_t ns, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  (void) block;
  const float *in = src;
  for (uint32_t s = 0; s < ns; s++)
  {
    float b0 = c[5 * s], b1 = c[5 * s + 1], b2 = c[5 * s + 2], a1 = c[5 * s + 3], a2 = c[5 * s + 4];
    float d1 = st[2 * s], d2 = st[2 * s + 1];
    for (uint32_t i = 0; i < n; i++)
    {
      float x = in[i];
      float p0 = b0 * x, p1 = b1 * x;
      float y = p0 + d1;
      float p3 = a1 * y, p2 = b2 * x;
      float A1 = p1 + p3;
      float p4 = a2 * y;
      d1 = A1 + d2;
      d2 = p2 + p4;
      dst[i] = y;
    }
    st[2 * s] = d1; st[2 * s + 1] = d2;
    in = dst;
  }
}

/* arm_biquad_cascade_stereo_df2T_f32.c:142 ff., state {d1a,d2a,d1b,d2b} per stage (:428-457), interleaved L/R frames. */
void port_biquad_stereo_df2T_f32 (const float *c, uint32_t ns, float *st, const float *src, float *dst, uint32_t nf, uint32_t block)
{
  (void) block;
  const float *in = src;
  for (uint32_t s = 0; s < ns; s++)
  {
    float b0 = c[5 * s], b1 = c[5 * s + 1], b2 = c[5 * s + 2], a1 = c[5 * s + 3], a2 = c[5 * s + 4];
    float d1a = st[4 * s], d2a = st[4 * s + 1], d1b = st[4 * s + 2], d2b = st[4 * s + 3];
    for (uint32_t i = 0; i < nf; i++)
    {
      float xa = in[2 * i], xb = in[2 * i + 1];
      float ya = (b0 * xa) + d1a, yb = (b0 * xb) + d1b;
      d1a = ((b1 * xa) + (a1 * ya)) + d2a; d1b = ((b1 * xb) + (a1 * yb)) + d2b;
      d2a = (b2 * xa) + (a2 * ya);         d2b = (b2 * xb) + (a2 * yb);
      dst[2 * i] = ya; dst[2 * i + 1] = yb;
    }
    st[4 * s] = d1a; st[4 * s + 1] = d2a; st[4 * s + 2] = d1b; st[4 * s + 3] = d2b;
    in = dst;
  }
}

/* arm_biquad_cascade_df1_f32.c:165 ff., expression :367: acc = ((((b0*x)+(b1*x1))+(b2*x2))+(a1*y1))+(a2*y2); state {x1,x2,y1,y2} */
void port_biquad_df1_f32 (const float *c, uint32_t ns, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  (void) block;
  const float *in = src;
  for (uint32_t s = 0; s < ns; s++)
  {
    float b0 = c[5 * s], b1 = c[5 * s + 1], b2 = c[5 * s + 2], a1 = c[5 * s + 3], a2 = c[5 * s + 4];
    float x1 = st[4 * s], x2 = st[4 * s + 1], y1 = st[4 * s + 2], y2 = st[4 * s + 3];
    for (uint32_t i = 0; i < n; i++)
    {
      float x = in[i];
      float acc = ((((b0 * x) + (b1 * x1)) + (b2 * x2)) + (a1 * y1)) + (a2 * y2);
      x2 = x1; x1 = x; y2 = y1; y1 = acc;
      dst[i] = acc;
    }
    st[4 * s] = x1; st[4 * s + 1] = x2; st[4 * s + 2] = y1; st[4 * s + 3] = y2;
    in = dst;
  }
}

/* arm_biquad_cascade_df1_q15.c:300-400 (plain-C branch): coeffs {b0,0,b1,b2,a1,a2} (:318-323), q63 accumulator,
 * acc >> (15 - postShift) truncated to q31 by __SSAT's prototype, saturate 16. */
void port_biquad_df1_q15 (const int16_t *c, uint32_t ns, int32_t ps, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  (void) block;
  const int16_t *in = src;
  int32_t shift = 15 - ps;
  for (uint32_t s = 0; s < ns; s++)
  {
    int16_t b0 = c[6 * s], b1 = c[6 * s + 2], b2 = c[6 * s + 3], a1 = c[6 * s + 4], a2 = c[6 * s + 5];
    int16_t x1 = st[4 * s], x2 = st[4 * s + 1], y1 = st[4 * s + 2], y2 = st[4 * s + 3];
    for (uint32_t i = 0; i < n; i++)
    {
      int16_t x = in[i];
      int64_t acc = (int32_t) b0 * x;
      acc += (int32_t) b1 * x1; acc += (int32_t) b2 * x2; acc += (int32_t) a1 * y1; acc += (int32_t) a2 * y2;
      int32_t y = slo_ssat16 ((int32_t) (acc >> shift));
      x2 = x1; x1 = x; y2 = y1; y1 = (int16_t) y;
      dst[i] = (int16_t) y;
    }
    st[4 * s] = x1; st[4 * s + 1] = x2; st[4 * s + 2] = y1; st[4 * s + 3] = y2;
    in = dst;
  }
}

/* arm_biquad_cascade_df1_q31.c:60 ff. (plain-C branch :100-200): q63 accumulator of q31*q31 products (wraps, doc :50-56),
 * y = (q31)(acc >> (31 - postShift)) assembled from the two halves (:131-138) — i.e. plain 64-bit shift, low 32 bits. */
void port_biquad_df1_q31 (const int32_t *c, uint32_t ns, int32_t ps, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  (void) block;
  const int32_t *in = src;
  uint32_t sh = 31u - (uint32_t) ps;
  for (uint32_t s = 0; s < ns; s++)
  {
    int32_t b0 = c[5 * s], b1 = c[5 * s + 1], b2 = c[5 * s + 2], a1 = c[5 * s + 3], a2 = c[5 * s + 4];
    int32_t x1 = st[4 * s], x2 = st[4 * s + 1], y1 = st[4 * s + 2], y2 = st[4 * s + 3];
    for (uint32_t i = 0; i < n; i++)
    {
      int32_t x = in[i];
      uint64_t acc = (uint64_t) ((int64_t) b0 * x);
      acc += (uint64_t) ((int64_t) b1 * x1); acc += (uint64_t) ((int64_t) b2 * x2);
      acc += (uint64_t) ((int64_t) a1 * y1); acc += (uint64_t) ((int64_t) a2 * y2);
      int32_t y = (int32_t) (uint32_t) (acc >> sh);
      x2 = x1; x1 = x; y2 = y1; y1 = y;
      dst[i] = y;
    }
    st[4 * s] = x1; st[4 * s + 1] = x2; st[4 * s + 2] = y1; st[4 * s + 3] = y2;
    in = dst;
  }
}
