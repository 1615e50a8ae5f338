/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * arm_bitreversal_32 / arm_bitreversal_16 exist in the reference only as Cortex-M assembly
 * (Drivers/CMSIS/DSP/Source/TransformFunctions/arm_bitreversal2.S:142-212), which cannot be assembled for
 * x86. This is the C statement of what that assembly does, needed so arm_cfft_* / arm_rfft_* link on the host:
 * the table holds pairs of BYTE offsets; _32 swaps the two 32-bit words at each offset pair (re, then im at +4),
 * _16 halves the offsets (LSR #1) and swaps one 32-bit word (= one packed q15 re/im pair).
 */
#include <stdint.h>
#include <string.h>

void arm_bitreversal_32 (uint32_t *pSrc, const uint16_t bitRevLen, const uint16_t *pBitRevTab)
{
  uint8_t *base = (uint8_t *) pSrc;
  for (uint32_t i = 0; i < bitRevLen; i += 2)
  {
    uint32_t a = pBitRevTab[i], b = pBitRevTab[i + 1], t0[2], t1[2];
    memcpy (t0, base + a, 8); memcpy (t1, base + b, 8);
    memcpy (base + a, t1, 8); memcpy (base + b, t0, 8);
  }
}

void arm_bitreversal_16 (uint16_t *pSrc, const uint16_t bitRevLen, const uint16_t *pBitRevTab)
{
  uint8_t *base = (uint8_t *) pSrc;
  for (uint32_t i = 0; i < bitRevLen; i += 2)
  {
    uint32_t a = pBitRevTab[i] >> 1, b = pBitRevTab[i + 1] >> 1, t0, t1;
    memcpy (&t0, base + a, 4); memcpy (&t1, base + b, 4);
    memcpy (base + a, &t1, 4); memcpy (base + b, &t0, 4);
  }
}
