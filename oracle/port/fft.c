/* ORACLE / TEST INFRASTRUCTURE ONLY. Transforms.
 * port_cfft_f32 restates the CONTRACT of Source/TransformFunctions/arm_cfft_f32.c:562-615 — in-place interleaved
 * re/im, forward unscaled, inverse = conj-in / forward / conj-and-scale-by-1/N out (:571-580, :604-614), natural-order
 * output when bitrev=1 — not its radix-8 butterfly schedule (arm_cfft_radix8_f32.c:45). The butterflies here are a
 * radix-2 decimation-in-time evaluated in double and rounded once to float, so the port is the correctly-rounded
 * answer; it agrees with the reference build to a few 1e-7 of the output RMS (tests/test_oracle_pinning.py).
 * bitrev=0 (digit-reversed output order of the radix-8 schedule) is not restated: ref_ only.
 * arm_cfft_q15 / arm_cfft_q31 (fixed-point radix-4 with per-stage scaling, arm_cfft_radix4_q15.c:69-76) are ref_ only. */
#include "port_common.h"

static void slo_fft_double (double *re, double *im, uint32_t N, int inverse)
{
  /* bit-reversal permutation */
  for (uint32_t i = 1, j = 0; i < N; i++)
  {
    uint32_t bit = N >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (uint32_t len = 2; len <= N; len <<= 1)
  {
    double ang = (inverse ? 2.0 : -2.0) * M_PI / (double) len;
    for (uint32_t i = 0; i < N; i += len)
      for (uint32_t k = 0; k < len / 2; k++)
      {
        double wr = cos (ang * k), wi = sin (ang * k);
        uint32_t a = i + k, b = i + k + len / 2;
        double tr = re[b] * wr - im[b] * wi, ti = re[b] * wi + im[b] * wr;
        re[b] = re[a] - tr; im[b] = im[a] - ti;
        re[a] += tr; im[a] += ti;
      }
  }
}

void port_cfft_f32 (float *d, uint32_t N, int ifft, int bitrev)
{
  (void) bitrev;
  double *re = (double *) malloc (sizeof (double) * N), *im = (double *) malloc (sizeof (double) * N);
  for (uint32_t i = 0; i < N; i++) { re[i] = d[2 * i]; im[i] = d[2 * i + 1]; }
  slo_fft_double (re, im, N, ifft);
  double s = ifft ? 1.0 / (double) N : 1.0;
  for (uint32_t i = 0; i < N; i++) { d[2 * i] = (float) (re[i] * s); d[2 * i + 1] = (float) (im[i] * s); }
  free (re); free (im);
}

/* arm_rfft_fast_f32.c:288-312 : forward DESTROYS its input (:300-306 runs the N/2 complex FFT in place) and packs
 * {Re X0, Re X(N/2), Re X1, Im X1, ...} (:47-66); inverse takes that packing back to N real samples. */
void port_rfft_fast_f32 (float *in, float *out, uint32_t N, int ifft)
{
  double *re = (double *) calloc (N, sizeof (double)), *im = (double *) calloc (N, sizeof (double));
  if (!ifft)
  {
    for (uint32_t i = 0; i < N; i++) re[i] = in[i];
    slo_fft_double (re, im, N, 0);
    out[0] = (float) re[0]; out[1] = (float) re[N / 2];
    for (uint32_t k = 1; k < N / 2; k++) { out[2 * k] = (float) re[k]; out[2 * k + 1] = (float) im[k]; }
  }
  else
  {
    re[0] = in[0]; re[N / 2] = in[1];
    for (uint32_t k = 1; k < N / 2; k++) { re[k] = in[2 * k]; im[k] = in[2 * k + 1]; re[N - k] = in[2 * k]; im[N - k] = -in[2 * k + 1]; }
    slo_fft_double (re, im, N, 1);
    for (uint32_t i = 0; i < N; i++) out[i] = (float) (re[i] / (double) N);
  }
  free (re); free (im);
}
