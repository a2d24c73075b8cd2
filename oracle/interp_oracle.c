/* interp_oracle.c — CPU oracle for interp1 / interp2.  TEST INFRASTRUCTURE ONLY.
 * PARITY UNPINNED against real Armadillo (absent from the image); see oracle.h and the
 * algorithm statement at the top of interp_oracle_impl.inc. */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include "oracle.h"

/* pass order of interp2 (see two_pass in the .inc): 0 = along X then Y (default), 1 = along Y then X */
static int oracle_interp2_yfirst = 0;
void oracle_interp2_set_order(int y_first) { oracle_interp2_yfirst = y_first ? 1 : 0; }

#define REAL double
#define SUFFIX(name) name##_f64
#define REAL_NAN ((double)NAN)
#define REAL_INF ((double)INFINITY)
#include "interp_oracle_impl.inc"
#undef REAL
#undef SUFFIX
#undef REAL_NAN
#undef REAL_INF

#define REAL float
#define SUFFIX(name) name##_f32
#define REAL_NAN ((float)NAN)
#define REAL_INF ((float)INFINITY)
#include "interp_oracle_impl.inc"
#undef REAL
#undef SUFFIX
#undef REAL_NAN
#undef REAL_INF
