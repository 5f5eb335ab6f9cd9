/* oracle/stubs/gsl/gsl_math.h -- TEST INFRASTRUCTURE ONLY: the reference uses <gsl/gsl_math.h> for M_PI & co. */
#ifndef ORACLE_STUB_GSL_MATH_H
#define ORACLE_STUB_GSL_MATH_H
#include <math.h>
#include <limits.h>
#include <float.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923132169164
#endif
#ifndef M_2_PI
#define M_2_PI 0.63661977236758134307553505349
#endif
#define gsl_finite(x) isfinite(x)
#endif
