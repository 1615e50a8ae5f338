/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * The q15 FFT family as the FIRMWARE builds it: the firmware compiles CMSIS-DSP with ARM_MATH_CM4
 * (.cproject:44), which defines ARM_MATH_DSP (arm_math.h:321-323) and selects the SIMD-intrinsic branch of
 * arm_radix4_butterfly_q15 (TransformFunctions/arm_cfft_radix4_q15.c:156-560) and arm_cfft_radix4by2_q15
 * (arm_cfft_q15.c:134-236). That branch rounds differently from the shift-based C branch the default oracle build
 * (ARM_MATH_CM3) takes — the only routines on this path where the two differ — so for the q15 FFT this build is the authority.
 *
 * How it is built on a host without copying a line of the reference: arm_math.h is included FIRST while ARM_MATH_DSP is still
 * undefined, so it supplies its own C statements of __QADD16, __SHADD16, __SMUAD, ... (arm_math.h:675-1004, "C custom defined
 * intrinsic function for M3 and M0 processors"); THEN ARM_MATH_DSP is defined and the two reference sources are included from
 * where they lie (-I Drivers/CMSIS/DSP/Source), so their function bodies compile the branch the Cortex-M4 build compiles, on
 * top of those C intrinsics. The public names are prefixed so that they can live next to the CM3-path objects in one library. */
#include "arm_math.h"
#include "arm_const_structs.h"
#define ARM_MATH_DSP
#define arm_radix4_butterfly_q15          cm4_arm_radix4_butterfly_q15
#define arm_radix4_butterfly_inverse_q15  cm4_arm_radix4_butterfly_inverse_q15
#define arm_cfft_radix4_q15               cm4_arm_cfft_radix4_q15
#define arm_cfft_radix4by2_q15            cm4_arm_cfft_radix4by2_q15
#define arm_cfft_radix4by2_inverse_q15    cm4_arm_cfft_radix4by2_inverse_q15
#define arm_cfft_q15                      cm4_arm_cfft_q15
#include "TransformFunctions/arm_cfft_radix4_q15.c"
#include "TransformFunctions/arm_cfft_q15.c"
#undef arm_cfft_q15

static const arm_cfft_instance_q15 *cm4_inst (uint32_t N)
{
  switch (N)
  {
    case 16: return &arm_cfft_sR_q15_len16;     case 32: return &arm_cfft_sR_q15_len32;
    case 64: return &arm_cfft_sR_q15_len64;     case 128: return &arm_cfft_sR_q15_len128;
    case 256: return &arm_cfft_sR_q15_len256;   case 512: return &arm_cfft_sR_q15_len512;
    case 1024: return &arm_cfft_sR_q15_len1024; case 2048: return &arm_cfft_sR_q15_len2048;
    case 4096: return &arm_cfft_sR_q15_len4096; default: return 0;
  }
}
void ref_cfft_q15_cm4 (int16_t *d, uint32_t N, int ifft, int bitrev) { cm4_arm_cfft_q15 (cm4_inst (N), d, (uint8_t) ifft, (uint8_t) bitrev); }
