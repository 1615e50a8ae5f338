/* ORACLE / TEST INFRASTRUCTURE ONLY. Complex math, statistics, fast math. */
#include "port_common.h"
#include <stdio.h>

/* Source/ComplexMathFunctions/arm_cmplx_mult_cmplx_f32.c:72 : (ac - bd, ad + bc) */
void port_cmplx_mult_cmplx_f32 (const float *a, const float *b, float *d, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++)
  {
    float ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
    d[2 * i] = (ar * br) - (ai * bi);
    d[2 * i + 1] = (ar * bi) + (ai * br);
  }
}
/* arm_cmplx_mult_real_f32.c:73 */
void port_cmplx_mult_real_f32 (const float *a, const float *r, float *d, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) { d[2 * i] = a[2 * i] * r[i]; d[2 * i + 1] = a[2 * i + 1] * r[i]; } }
/* arm_cmplx_conj_f32.c:71 */
void port_cmplx_conj_f32 (const float *a, float *d, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) { d[2 * i] = a[2 * i]; d[2 * i + 1] = -a[2 * i + 1]; } }
/* arm_cmplx_mag_f32.c:72 : arm_sqrt_f32(re*re + im*im) ; arm_sqrt_f32 (Include/arm_math.h:5726-5751) is sqrtf, negative -> 0 */
void port_cmplx_mag_f32 (const float *a, float *d, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++)
  {
    float v = (a[2 * i] * a[2 * i]) + (a[2 * i + 1] * a[2 * i + 1]);
    d[i] = v >= 0.0f ? sqrtf (v) : 0.0f;
  }
}
/* arm_cmplx_mag_squared_f32.c */
void port_cmplx_mag_squared_f32 (const float *a, float *d, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) d[i] = (a[2 * i] * a[2 * i]) + (a[2 * i + 1] * a[2 * i + 1]); }

/* Source/FastMathFunctions/arm_sqrt_q15.c:50-140 : fast inverse-sqrt seed through the float bit pattern, three
 * fixed-point Newton steps with q15 truncations, multiply back. Every cast below is one the reference makes. */
static int16_t slo_sqrt_q15 (int16_t in)
{
  int16_t number = in, temp1, var1, signBits1, half;
  int32_t bits;
  float tf;
  if (number <= 0) return 0;
  signBits1 = (int16_t) (__builtin_clz ((uint32_t) number) - 17);
  if ((signBits1 % 2) == 0) number = (int16_t) (number << signBits1);
  else number = (int16_t) (number << (signBits1 - 1));
  half = (int16_t) (number >> 1);
  temp1 = number;
  tf = number * 3.051757812500000e-005f;
  memcpy (&bits, &tf, 4);
  bits = 0x5f3759df - (bits >> 1);
  memcpy (&tf, &bits, 4);
  var1 = (int16_t) (int32_t) (tf * 16384);
  for (int it = 0; it < 3; it++)
  {
    int16_t sq = (int16_t) (((int32_t) var1 * var1) >> 15);
    int16_t hs = (int16_t) (((int32_t) sq * (int32_t) half) >> 15);
    var1 = (int16_t) (((int16_t) (((int32_t) var1 * (0x3000 - hs)) >> 15)) << 2);
  }
  var1 = (int16_t) (((int16_t) (((int32_t) temp1 * var1) >> 15)) << 1);
  if ((signBits1 % 2) == 0) var1 = (int16_t) (var1 >> (signBits1 / 2));
  else var1 = (int16_t) (var1 >> ((signBits1 - 1) / 2));
  return var1;
}

/* arm_cmplx_mag_q15.c:120-130 : result is 2.14 (:50): sqrt_q15((q15)(((q63)re*re + im*im) >> 17)) */
void port_cmplx_mag_q15 (const int16_t *a, int16_t *d, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++)
  {
    int32_t acc0 = (int32_t) a[2 * i] * a[2 * i], acc1 = (int32_t) a[2 * i + 1] * a[2 * i + 1];
    d[i] = slo_sqrt_q15 ((int16_t) (((int64_t) acc0 + acc1) >> 17));
  }
}

/* Source/StatisticsFunctions/arm_max_f32.c:58 : first maximum wins (strict <) */
float port_max_f32 (const float *s, uint32_t n, uint32_t *idx)
{
  float out = s[0]; uint32_t oi = 0;
  for (uint32_t i = 1; i < n; i++) if (out < s[i]) { out = s[i]; oi = i; }
  if (idx) *idx = oi;
  return out;
}
/* arm_rms_f32.c:64,:122 : sequential sum of squares, sqrt(sum / N) */
float port_rms_f32 (const float *s, uint32_t n)
{
  float sum = 0.0f;
  for (uint32_t i = 0; i < n; i++) sum += s[i] * s[i];
  float v = sum / (float) n;
  return v >= 0.0f ? sqrtf (v) : 0.0f;
}
/* arm_power_f32.c:64, arm_mean_f32.c */
float port_power_f32 (const float *s, uint32_t n) { float sum = 0.0f; for (uint32_t i = 0; i < n; i++) sum += s[i] * s[i]; return sum; }
float port_mean_f32 (const float *s, uint32_t n) { float sum = 0.0f; for (uint32_t i = 0; i < n; i++) sum += s[i]; return sum / (float) n; }
/* arm_max_q15.c */
int16_t port_max_q15 (const int16_t *s, uint32_t n, uint32_t *idx)
{
  int16_t out = s[0]; uint32_t oi = 0;
  for (uint32_t i = 1; i < n; i++) if (out < s[i]) { out = s[i]; oi = i; }
  if (idx) *idx = oi;
  return out;
}
/* arm_rms_q15.c:58,:107-131 : q63 sum of squares, sat16((sum / N) >> 15), arm_sqrt_q15 */
int16_t port_rms_q15 (const int16_t *s, uint32_t n)
{
  int64_t sum = 0;
  for (uint32_t i = 0; i < n; i++) sum += (int32_t) s[i] * s[i];
  return slo_sqrt_q15 ((int16_t) slo_ssat16 ((int32_t) ((sum / (int64_t) n) >> 15)));
}

/* Source/FastMathFunctions/arm_sin_f32.c:72-119, arm_cos_f32.c: 512-entry table + linear interpolation.
 * The table (CommonTables/arm_common_tables.c:21895) is sin(2*pi*n/512) written with 8 decimals; the same values are
 * regenerated here by formatting to 8 decimals and parsing as float. */
static float slo_sin_table[513];
static int slo_sin_table_ready = 0;
static void slo_sin_table_init (void)
{
  char buf[32];
  for (int i = 0; i <= 512; i++)
  {
    snprintf (buf, sizeof buf, "%.8f", sin (2.0 * M_PI * (double) i / 512.0));
    slo_sin_table[i] = strtof (buf, 0);
  }
  slo_sin_table_ready = 1;
}
static float slo_sin1 (float x, int is_cos)
{
  float in, fract, findex, a, b;
  int32_t n; uint16_t index;
  if (!slo_sin_table_ready) slo_sin_table_init ();
  if (!is_cos)
  {
    if ((x < 0.0f) && (x >= -1.9e-7f)) return x;
    in = x * 0.159154943092f;
    n = (int32_t) in;
    if (x < 0.0f) n--;
  }
  else
  {
    in = x * 0.159154943092f + 0.25f;
    n = (int32_t) in;
    if (in < 0.0f) n--;
  }
  in = in - (float) n;
  findex = (float) 512 * in;
  index = ((uint16_t) findex) & 0x1ff;
  fract = findex - (float) index;
  a = slo_sin_table[index]; b = slo_sin_table[index + 1];
  return (1.0f - fract) * a + fract * b;
}
void port_sin_f32 (const float *x, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = slo_sin1 (x[i], 0); }
void port_cos_f32 (const float *x, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = slo_sin1 (x[i], 1); }
