/*
 * drtk_oracle.c -- CPU oracle for the DRTK rasterisation hot path (f32 and f64 builds of
 * drtk_oracle_impl.h).  TEST INFRASTRUCTURE ONLY -- see the header of drtk_oracle_impl.h.
 *
 * Build (oracle/Makefile):  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC drtk_oracle.c -lm
 * -ffp-contract=off matters: the oracle decides itself where a fused multiply-add is used
 * (mode 1 of oracle_rasterize) and where it must not be.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUFFIX _f32
#define EPSVAL 1e-8f
#define FMA fmaf
#define FMIN fminf
#define FMAX fmaxf
#define SQRT sqrtf
#include "drtk_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef EPSVAL
#undef FMA
#undef FMIN
#undef FMAX
#undef SQRT

#define REAL double
#define SUFFIX _f64
#define EPSVAL 1e-16
#define FMA fma
#define FMIN fmin
#define FMAX fmax
#define SQRT sqrt
#include "drtk_oracle_impl.h"

int oracle_abi_version(void) { return 1; }
