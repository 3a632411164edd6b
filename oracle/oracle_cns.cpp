// oracle/oracle_cns.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Consensus-flavour routines.
#include "oracle.h"
extern "C" {
int orc_cns_get_alignment(const char*, int, int, const char*, int, int, double, int, int32_t*, char*, char*, int) { return -1; }
int orc_normalize_gaps(const char*, const char*, int, int, char*, char*, int) { return -1; }
}
