/* ORACLE / TEST INFRASTRUCTURE ONLY. Restatement of the firmware ring, /root/reference/Core/Src/dsp_if.c, as an
 * instance (the reference keeps one global per direction, dsp_if.c:32-35) with the geometry as a run-time field
 * (compile-time in the reference, Core/Inc/dsp_if.h:69-85). Used to check the product's host-side pointer logic on
 * the GPU box, where the reference build may be absent. */
#include <stdint.h>
#include <string.h>

#define SLO_RING_MAX 4096
typedef struct
{
  uint32_t size;                  /* DSP_BUFF_SIZE = fs/1000 * 8  (dsp_if.h:81-85) */
  int16_t i[SLO_RING_MAX], q[SLO_RING_MAX];
  uint32_t enable, rd, wr;
} slo_ring;

void port_ring_init (slo_ring *r, uint32_t fs) { memset (r, 0, sizeof *r); r->size = fs / 1000u * 8u; }

static void ring_put (slo_ring *r, int16_t i, int16_t q)    /* dsp_if.c:93-104 / :227-238 */
{
  r->i[r->wr] = i; r->q[r->wr] = q;
  r->wr++;
  if (r->wr == r->size) r->wr = 0;
}

/* Producer with drift slip. is_out=0: DSP_In_Buff_Write (dsp_if.c:250-301) — gap is only computed once the reader has
 * armed buff_enable (:254-264), so before that gap = 0 < N/4 and wr slips forward every call.
 * is_out=1: DSP_Out_Buff_Write (dsp_if.c:116-180) — the writer itself arms buff_enable and places wr half a ring
 * ahead of rd (:124-134), gap always computed. nhw = half-words (2 per frame). */
void port_ring_write (slo_ring *r, int is_out, const int16_t *buf, uint32_t nhw)
{
  uint32_t N = r->size, gap = 0;
  if (is_out)
  {
    if (!r->enable)
    {
      r->wr = r->rd + N / 2;
      if (r->wr >= N) r->wr -= N;
      r->enable = 1;
    }
    gap = r->wr; if (r->rd > r->wr) gap += N; gap -= r->rd;
  }
  else if (r->enable)
  {
    gap = r->wr; if (r->rd > r->wr) gap += N; gap -= r->rd;
  }
  gap &= 0xFFFFu;                                                    /* uint16_t gap in the reference */
  if (gap > 3u * N / 4u) { if (r->wr < 1u) r->wr += N; r->wr--; }     /* wr is faster: step back */
  if (gap < N / 4u) { r->wr++; if (r->wr >= N) r->wr -= N; }          /* rd is faster: step forward */
  for (uint32_t k = 0; k < nhw; k += 2) ring_put (r, buf[k], buf[k + 1]);
  ring_put (r, buf[nhw - 2], buf[nhw - 1]);                           /* repeat last frame ... */
  if (r->wr < 1u) r->wr += N;
  r->wr--;                                                            /* ... and step back over it */
}

/* Consumer. is_out=0: DSP_In_Buff_Read (dsp_if.c:310-340) — first call sets rd = wr + N/2, and on overflow RESETS TO 0
 * (:320-323, not rd - N). is_out=1: DSP_Out_Buff_Read (dsp_if.c:204-219). nhw = half-words. */
void port_ring_read (slo_ring *r, int is_out, int16_t *buf, uint32_t nhw)
{
  uint32_t N = r->size;
  if (!is_out && !r->enable)
  {
    r->rd = r->wr + N / 2;
    if (r->rd >= N) r->rd = 0;
    r->enable = 1;
  }
  for (uint32_t k = 0; k < nhw; k += 2)
  {
    buf[k] = r->i[r->rd]; buf[k + 1] = r->q[r->rd];
    r->rd++;
    if (r->rd >= N) r->rd = 0;
  }
}

void port_ring_mute (slo_ring *r) { memset (r->i, 0, sizeof r->i); memset (r->q, 0, sizeof r->q); }  /* dsp_if.c:188-195 */
uint32_t port_ring_sizeof (void) { return (uint32_t) sizeof (slo_ring); }
void port_ring_get_ptrs (const slo_ring *r, uint32_t *out) { out[0] = r->enable; out[1] = r->rd; out[2] = r->wr; }
void port_ring_get_iq (const slo_ring *r, int16_t *i, int16_t *q) { memcpy (i, r->i, 2 * r->size); memcpy (q, r->q, 2 * r->size); }
