/* oracle/rt_oracle.c — ORACLE L1 (test infrastructure, NOT product code).
 *
 * CPU restatement of the reference's path-tracing hot path in plain C, compiled
 * twice from rt_oracle_impl.h: orc64_* (double, the arithmetic of
 * rt_in_one_weekend) and orc32_* (float, the arithmetic of the CUDA trees).
 * Pinned against the real reference (oracle/_ref/libref_l0.so, built from the
 * unmodified sources under /root/reference) by tests/test_oracle_pinning.py and
 * by the committed golden vectors in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this library; the product path (libb200rt.so) never links or loads it.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC rt_oracle.c -lm  (oracle/Makefile)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rt_capi.h"

#define REAL double
#define REAL_IS_FLOAT 0
#define FN(x) orc64_##x
#define V3 orc64_V3
#define SQRT sqrt
#define FABS fabs
#define FMIN fmin
#define FMAX fmax
#define POW pow
#define SIN sin
#define LOG log
#define FLOOR floor
#define ACOS acos
#define ATAN2 atan2
#define INFINITY_R ((double)INFINITY)
#include "rt_oracle_impl.h"
#undef REAL
#undef REAL_IS_FLOAT
#undef FN
#undef V3
#undef SQRT
#undef FABS
#undef FMIN
#undef FMAX
#undef POW
#undef SIN
#undef LOG
#undef FLOOR
#undef ACOS
#undef ATAN2
#undef INFINITY_R

#define REAL float
#define REAL_IS_FLOAT 1
#define FN(x) orc32_##x
#define V3 orc32_V3
#define SQRT sqrtf
#define FABS fabsf
#define FMIN fminf
#define FMAX fmaxf
#define POW powf
#define SIN sinf
#define LOG logf
#define FLOOR floorf
#define ACOS acosf
#define ATAN2 atan2f
#define INFINITY_R ((float)INFINITY)
#include "rt_oracle_impl.h"

void orc_srand(unsigned seed) { srand(seed); }
int orc_version(void) { return 2; }
