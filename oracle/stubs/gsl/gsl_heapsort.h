/* oracle/stubs/gsl/gsl_heapsort.h -- TEST INFRASTRUCTURE ONLY: included by the reference but no symbol of it is used. */
#ifndef ORACLE_STUB_GSL_HEAPSORT_H
#define ORACLE_STUB_GSL_HEAPSORT_H
#include <stddef.h>
#endif
