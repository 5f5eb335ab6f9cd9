/* oracle/stubs/gsl_stub.c -- TEST INFRASTRUCTURE ONLY: the two GSL sort entry points the hot path's
 * host set-up uses (raytrace_utils.c:493, map_shuffle.c:133), as stable index/value sorts. */
#include <stdlib.h>
#include "gsl/gsl_sort_long.h"
static const long *g_keys; static size_t g_stride;
static int cmp_idx(const void *a, const void *b)
{
  size_t ia = *(const size_t*)a, ib = *(const size_t*)b;
  long ka = g_keys[ia * g_stride], kb = g_keys[ib * g_stride];
  if (ka < kb) return -1; if (ka > kb) return 1; return (ia < ib) ? -1 : (ia > ib);
}
static int cmp_long(const void *a, const void *b) { long x = *(const long*)a, y = *(const long*)b; return (x < y) ? -1 : (x > y); }
void gsl_sort_long_index(size_t *p, const long *data, const size_t stride, const size_t n)
{ for (size_t i = 0; i < n; ++i) p[i] = i; g_keys = data; g_stride = stride; qsort(p, n, sizeof(size_t), cmp_idx); }
void gsl_sort_long(long *data, const size_t stride, const size_t n)
{ if (stride != 1) abort(); qsort(data, n, sizeof(long), cmp_long); }
