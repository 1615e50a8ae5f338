/* ORACLE / TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's CMSIS-DSP V1.5.3 routines.
 * Paths below are relative to /root/reference/Drivers/CMSIS/DSP/. Not shipped; the product never links this. */
#ifndef SLO_PORT_COMMON_H
#define SLO_PORT_COMMON_H
#define SLO_PREFIX port_
#include "../slo_api.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

/* __SSAT(x, 16) as the plain-C fallback defines it: Include/../../Include/cmsis_gcc.h:1299-1315.
 * The argument is an int32_t, so 64-bit accumulators are truncated to 32 bits BEFORE saturation. */
static inline int32_t slo_ssat16 (int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
static inline int32_t slo_clip_q63_to_q31 (int64_t x)
{ return ((int32_t) (x >> 32) != ((int32_t) x >> 31)) ? ((0x7FFFFFFF ^ ((int32_t) (x >> 63)))) : (int32_t) x; }
#endif
