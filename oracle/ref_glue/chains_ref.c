/* ORACLE / TEST INFRASTRUCTURE ONLY. ref_ instantiation of the shared chain glue. */
#define SLO_PREFIX ref_
#include "../chains.inc.c"
