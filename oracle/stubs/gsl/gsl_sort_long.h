/* oracle/stubs/gsl/gsl_sort_long.h -- TEST INFRASTRUCTURE ONLY (implemented in gsl_stub.c). */
#ifndef ORACLE_STUB_GSL_SORT_LONG_H
#define ORACLE_STUB_GSL_SORT_LONG_H
#include <stddef.h>
void gsl_sort_long(long *data, const size_t stride, const size_t n);
void gsl_sort_long_index(size_t *p, const long *data, const size_t stride, const size_t n);
#endif
