/* ORACLE / TEST INFRASTRUCTURE ONLY. Conversions and element-wise math. */
#include "port_common.h"

/* Source/SupportFunctions/arm_q15_to_float.c:87 : (float32_t) x / 32768.0f */
void port_q15_to_float (const int16_t *src, float *dst, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) dst[i] = (float) src[i] / 32768.0f; }

/* Source/SupportFunctions/arm_float_to_q15.c:117-120,:147 (ARM_MATH_ROUNDING not defined, .cproject:43-46):
 * multiply by 32768, C cast to q31 (truncation toward zero), saturate to 16 bits. */
void port_float_to_q15 (const float *src, int16_t *dst, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) dst[i] = (int16_t) slo_ssat16 ((int32_t) (src[i] * 32768.0f)); }

/* Source/BasicMathFunctions/arm_scale_f32.c:77, arm_mult_f32.c, arm_add_f32.c, arm_sub_f32.c, arm_abs_f32.c:63 (fabsf) */
void port_scale_f32 (const float *s, float k, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = s[i] * k; }
void port_mult_f32 (const float *a, const float *b, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = a[i] * b[i]; }
void port_add_f32 (const float *a, const float *b, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = a[i] + b[i]; }
void port_sub_f32 (const float *a, const float *b, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = a[i] - b[i]; }
void port_abs_f32 (const float *a, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = fabsf (a[i]); }

/* Source/BasicMathFunctions/arm_scale_q15.c:56,:138 : kShift = 15 - shift; sat16((x * scaleFract) >> kShift) */
void port_scale_q15 (const int16_t *s, int16_t k, int32_t shift, int16_t *d, uint32_t n)
{
  int32_t ksh = 15 - shift;
  for (uint32_t i = 0; i < n; i++) d[i] = (int16_t) slo_ssat16 (((int32_t) s[i] * k) >> ksh);
}
/* arm_add_q15.c:115, arm_sub_q15.c:115 : saturating */
void port_add_q15 (const int16_t *a, const int16_t *b, int16_t *d, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) d[i] = (int16_t) slo_ssat16 ((int32_t) a[i] + b[i]); }
void port_sub_q15 (const int16_t *a, const int16_t *b, int16_t *d, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) d[i] = (int16_t) slo_ssat16 ((int32_t) a[i] - b[i]); }
/* arm_abs_q15.c:140-146 : (in > 0) ? in : ((in == 0x8000) ? 0x7fff : -in) */
void port_abs_q15 (const int16_t *a, int16_t *d, uint32_t n)
{ for (uint32_t i = 0; i < n; i++) d[i] = (int16_t) (a[i] > 0 ? a[i] : (a[i] == (int16_t) 0x8000 ? 0x7fff : -a[i])); }
/* arm_shift_q15.c:190-230 : left shift saturates, right shift is arithmetic */
void port_shift_q15 (const int16_t *a, int32_t sh, int16_t *d, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++)
    d[i] = (sh >= 0) ? (int16_t) slo_ssat16 ((int32_t) a[i] << sh) : (int16_t) (a[i] >> (-sh));
}
