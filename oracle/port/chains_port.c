/* ORACLE / TEST INFRASTRUCTURE ONLY. port_ instantiation of the shared chain glue. */
#include "port_common.h"
#include "../chains.inc.c"
