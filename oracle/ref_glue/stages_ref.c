/* ORACLE / TEST INFRASTRUCTURE ONLY — never on the product path.
 *
 * ref_* stage functions: every body is a call to the reference's own vendored CMSIS-DSP V1.5.3 routine
 * of the same name (compiled in place from /root/reference/Drivers/CMSIS/DSP/Source by oracle/Makefile).
 * The glue only (1) fills the CMSIS instance struct WITHOUT calling arm_*_init_* (those zero the state,
 * and here the caller owns the state across calls) and (2) walks `n` samples in `block`-sized calls,
 * the way the firmware's 1 ms cadence would.
 */
#define SLO_PREFIX ref_
#include "../slo_api.h"
#include "arm_math.h"
#include "arm_const_structs.h"

void ref_q15_to_float (const int16_t *src, float *dst, uint32_t n) { arm_q15_to_float ((q15_t *) src, dst, n); }
void ref_float_to_q15 (const float *src, int16_t *dst, uint32_t n) { arm_float_to_q15 ((float *) src, dst, n); }

void ref_fir_f32 (const float *c, uint32_t nt, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  arm_fir_instance_f32 S = { (uint16_t) nt, st, (float *) c };
  for (uint32_t o = 0; o < n; o += block) arm_fir_f32 (&S, (float *) src + o, dst + o, block);
}
void ref_fir_q15 (const int16_t *c, uint32_t nt, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_instance_q15 S = { (uint16_t) nt, st, (q15_t *) c };
  for (uint32_t o = 0; o < n; o += block) arm_fir_q15 (&S, (q15_t *) src + o, dst + o, block);
}
void ref_fir_fast_q15 (const int16_t *c, uint32_t nt, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_instance_q15 S = { (uint16_t) nt, st, (q15_t *) c };
  for (uint32_t o = 0; o < n; o += block) arm_fir_fast_q15 (&S, (q15_t *) src + o, dst + o, block);
}
void ref_fir_q31 (const int32_t *c, uint32_t nt, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_instance_q31 S = { (uint16_t) nt, st, (q31_t *) c };
  for (uint32_t o = 0; o < n; o += block) arm_fir_q31 (&S, (q31_t *) src + o, dst + o, block);
}
void ref_fir_decimate_f32 (const float *c, uint32_t nt, uint32_t M, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  arm_fir_decimate_instance_f32 S = { (uint8_t) M, (uint16_t) nt, (float *) c, st };
  for (uint32_t o = 0; o < n; o += block) arm_fir_decimate_f32 (&S, (float *) src + o, dst + o / M, block);
}
void ref_fir_decimate_q15 (const int16_t *c, uint32_t nt, uint32_t M, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_decimate_instance_q15 S = { (uint8_t) M, (uint16_t) nt, (q15_t *) c, st };
  for (uint32_t o = 0; o < n; o += block) arm_fir_decimate_q15 (&S, (q15_t *) src + o, dst + o / M, block);
}
void ref_fir_interpolate_f32 (const float *c, uint32_t nt, uint32_t L, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  arm_fir_interpolate_instance_f32 S = { (uint8_t) L, (uint16_t) (nt / L), (float *) c, st };
  for (uint32_t o = 0; o < n; o += block) arm_fir_interpolate_f32 (&S, (float *) src + o, dst + o * L, block);
}
void ref_lms_norm_f32 (float *coeffs, uint32_t nt, float mu, float *st, float *en_x0, const float *src, const float *ref,
                        float *out, float *err, uint32_t n, uint32_t block)
{
  arm_lms_norm_instance_f32 S = { (uint16_t) nt, st, coeffs, mu, en_x0[0], en_x0[1] };
  for (uint32_t o = 0; o < n; o += block) arm_lms_norm_f32 (&S, (float *) src + o, (float *) ref + o, out + o, err + o, block);
  en_x0[0] = S.energy; en_x0[1] = S.x0;
}
void ref_fir_decimate_q31 (const int32_t *c, uint32_t nt, uint32_t M, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_decimate_instance_q31 S = { (uint8_t) M, (uint16_t) nt, (q31_t *) c, st };
  for (uint32_t o = 0; o < n; o += block) arm_fir_decimate_q31 (&S, (q31_t *) src + o, dst + o / M, block);
}
void ref_fir_interpolate_q31 (const int32_t *c, uint32_t nt, uint32_t L, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_interpolate_instance_q31 S = { (uint8_t) L, (uint16_t) (nt / L), (q31_t *) c, st };
  for (uint32_t o = 0; o < n; o += block) arm_fir_interpolate_q31 (&S, (q31_t *) src + o, dst + o * L, block);
}
void ref_fir_interpolate_q15 (const int16_t *c, uint32_t nt, uint32_t L, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  arm_fir_interpolate_instance_q15 S = { (uint8_t) L, (uint16_t) (nt / L), (q15_t *) c, st };
  for (uint32_t o = 0; o < n; o += block) arm_fir_interpolate_q15 (&S, (q15_t *) src + o, dst + o * L, block);
}

void ref_biquad_df2T_f32 (const float *c, uint32_t ns, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  arm_biquad_cascade_df2T_instance_f32 S = { (uint8_t) ns, st, (float *) c };
  for (uint32_t o = 0; o < n; o += block) arm_biquad_cascade_df2T_f32 (&S, (float *) src + o, dst + o, block);
}
void ref_biquad_stereo_df2T_f32 (const float *c, uint32_t ns, float *st, const float *src, float *dst, uint32_t nframes, uint32_t block)
{
  arm_biquad_cascade_stereo_df2T_instance_f32 S = { (uint8_t) ns, st, (float *) c };
  for (uint32_t o = 0; o < nframes; o += block) arm_biquad_cascade_stereo_df2T_f32 (&S, (float *) src + 2 * o, dst + 2 * o, block);
}
void ref_biquad_df1_f32 (const float *c, uint32_t ns, float *st, const float *src, float *dst, uint32_t n, uint32_t block)
{
  arm_biquad_casd_df1_inst_f32 S = { ns, st, (float *) c };
  for (uint32_t o = 0; o < n; o += block) arm_biquad_cascade_df1_f32 (&S, (float *) src + o, dst + o, block);
}
void ref_biquad_df1_q15 (const int16_t *c, uint32_t ns, int32_t ps, int16_t *st, const int16_t *src, int16_t *dst, uint32_t n, uint32_t block)
{
  arm_biquad_casd_df1_inst_q15 S = { (int8_t) ns, st, (q15_t *) c, (int8_t) ps };
  for (uint32_t o = 0; o < n; o += block) arm_biquad_cascade_df1_q15 (&S, (q15_t *) src + o, dst + o, block);
}
void ref_biquad_df1_q31 (const int32_t *c, uint32_t ns, int32_t ps, int32_t *st, const int32_t *src, int32_t *dst, uint32_t n, uint32_t block)
{
  arm_biquad_casd_df1_inst_q31 S = { ns, st, (q31_t *) c, (uint8_t) ps };
  for (uint32_t o = 0; o < n; o += block) arm_biquad_cascade_df1_q31 (&S, (q31_t *) src + o, dst + o, block);
}

static const arm_cfft_instance_f32 *cfft_f32_inst (uint32_t N)
{
  switch (N) {
    case 16: return &arm_cfft_sR_f32_len16;     case 32: return &arm_cfft_sR_f32_len32;
    case 64: return &arm_cfft_sR_f32_len64;     case 128: return &arm_cfft_sR_f32_len128;
    case 256: return &arm_cfft_sR_f32_len256;   case 512: return &arm_cfft_sR_f32_len512;
    case 1024: return &arm_cfft_sR_f32_len1024; case 2048: return &arm_cfft_sR_f32_len2048;
    case 4096: return &arm_cfft_sR_f32_len4096; default: return 0;
  }
}
static const arm_cfft_instance_q15 *cfft_q15_inst (uint32_t N)
{
  switch (N) {
    case 16: return &arm_cfft_sR_q15_len16;     case 32: return &arm_cfft_sR_q15_len32;
    case 64: return &arm_cfft_sR_q15_len64;     case 128: return &arm_cfft_sR_q15_len128;
    case 256: return &arm_cfft_sR_q15_len256;   case 512: return &arm_cfft_sR_q15_len512;
    case 1024: return &arm_cfft_sR_q15_len1024; case 2048: return &arm_cfft_sR_q15_len2048;
    case 4096: return &arm_cfft_sR_q15_len4096; default: return 0;
  }
}
static const arm_cfft_instance_q31 *cfft_q31_inst (uint32_t N)
{
  switch (N) {
    case 16: return &arm_cfft_sR_q31_len16;     case 32: return &arm_cfft_sR_q31_len32;
    case 64: return &arm_cfft_sR_q31_len64;     case 128: return &arm_cfft_sR_q31_len128;
    case 256: return &arm_cfft_sR_q31_len256;   case 512: return &arm_cfft_sR_q31_len512;
    case 1024: return &arm_cfft_sR_q31_len1024; case 2048: return &arm_cfft_sR_q31_len2048;
    case 4096: return &arm_cfft_sR_q31_len4096; default: return 0;
  }
}
void ref_cfft_f32 (float *d, uint32_t N, int ifft, int bitrev) { arm_cfft_f32 (cfft_f32_inst (N), d, (uint8_t) ifft, (uint8_t) bitrev); }
void ref_cfft_q15 (int16_t *d, uint32_t N, int ifft, int bitrev) { arm_cfft_q15 (cfft_q15_inst (N), d, (uint8_t) ifft, (uint8_t) bitrev); }
void ref_cfft_q31 (int32_t *d, uint32_t N, int ifft, int bitrev) { arm_cfft_q31 (cfft_q31_inst (N), d, (uint8_t) ifft, (uint8_t) bitrev); }
void ref_rfft_fast_f32 (float *in, float *out, uint32_t N, int ifft)
{
  arm_rfft_fast_instance_f32 S;
  arm_rfft_fast_init_f32 (&S, (uint16_t) N);
  arm_rfft_fast_f32 (&S, in, out, (uint8_t) ifft);
}

void ref_cmplx_mult_cmplx_f32 (const float *a, const float *b, float *dst, uint32_t n) { arm_cmplx_mult_cmplx_f32 ((float *) a, (float *) b, dst, n); }
void ref_cmplx_mult_real_f32 (const float *a, const float *r, float *dst, uint32_t n) { arm_cmplx_mult_real_f32 ((float *) a, (float *) r, dst, n); }
void ref_cmplx_conj_f32 (const float *a, float *dst, uint32_t n) { arm_cmplx_conj_f32 ((float *) a, dst, n); }
void ref_cmplx_mag_f32 (const float *a, float *dst, uint32_t n) { arm_cmplx_mag_f32 ((float *) a, dst, n); }
void ref_cmplx_mag_squared_f32 (const float *a, float *dst, uint32_t n) { arm_cmplx_mag_squared_f32 ((float *) a, dst, n); }
void ref_cmplx_mag_q15 (const int16_t *a, int16_t *dst, uint32_t n) { arm_cmplx_mag_q15 ((q15_t *) a, dst, n); }

float ref_max_f32 (const float *s, uint32_t n, uint32_t *idx) { float r; uint32_t i; arm_max_f32 ((float *) s, n, &r, &i); if (idx) *idx = i; return r; }
float ref_rms_f32 (const float *s, uint32_t n) { float r; arm_rms_f32 ((float *) s, n, &r); return r; }
float ref_power_f32 (const float *s, uint32_t n) { float r; arm_power_f32 ((float *) s, n, &r); return r; }
float ref_mean_f32 (const float *s, uint32_t n) { float r; arm_mean_f32 ((float *) s, n, &r); return r; }
int16_t ref_max_q15 (const int16_t *s, uint32_t n, uint32_t *idx) { q15_t r; uint32_t i; arm_max_q15 ((q15_t *) s, n, &r, &i); if (idx) *idx = i; return r; }
int16_t ref_rms_q15 (const int16_t *s, uint32_t n) { q15_t r; arm_rms_q15 ((q15_t *) s, n, &r); return r; }

void ref_scale_f32 (const float *s, float k, float *d, uint32_t n) { arm_scale_f32 ((float *) s, k, d, n); }
void ref_mult_f32 (const float *a, const float *b, float *d, uint32_t n) { arm_mult_f32 ((float *) a, (float *) b, d, n); }
void ref_add_f32 (const float *a, const float *b, float *d, uint32_t n) { arm_add_f32 ((float *) a, (float *) b, d, n); }
void ref_sub_f32 (const float *a, const float *b, float *d, uint32_t n) { arm_sub_f32 ((float *) a, (float *) b, d, n); }
void ref_abs_f32 (const float *a, float *d, uint32_t n) { arm_abs_f32 ((float *) a, d, n); }
void ref_scale_q15 (const int16_t *s, int16_t k, int32_t sh, int16_t *d, uint32_t n) { arm_scale_q15 ((q15_t *) s, k, (int8_t) sh, d, n); }
void ref_add_q15 (const int16_t *a, const int16_t *b, int16_t *d, uint32_t n) { arm_add_q15 ((q15_t *) a, (q15_t *) b, d, n); }
void ref_sub_q15 (const int16_t *a, const int16_t *b, int16_t *d, uint32_t n) { arm_sub_q15 ((q15_t *) a, (q15_t *) b, d, n); }
void ref_abs_q15 (const int16_t *a, int16_t *d, uint32_t n) { arm_abs_q15 ((q15_t *) a, d, n); }
void ref_shift_q15 (const int16_t *a, int32_t sh, int16_t *d, uint32_t n) { arm_shift_q15 ((q15_t *) a, (int8_t) sh, d, n); }

void ref_sin_f32 (const float *x, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = arm_sin_f32 (x[i]); }
void ref_cos_f32 (const float *x, float *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = arm_cos_f32 (x[i]); }
