/* oracle/stubs/fitsio.h -- TEST INFRASTRUCTURE ONLY.  Minimal CFITSIO surface used by read_ring_weights
 * (healpix_shtrans.c:361-423); implemented in fits_stub.c for BINTABLE column 1 of big-endian doubles. */
#ifndef ORACLE_STUB_FITSIO_H
#define ORACLE_STUB_FITSIO_H
#include <stdio.h>
typedef struct { FILE *fp; long data_start; long nrows; long repeat; long rowbytes; } fitsfile;
typedef long long LONGLONG;
#define READONLY 0
#define TDOUBLE 82
int fits_open_file(fitsfile **fptr, const char *name, int mode, int *status);
int fits_close_file(fitsfile *fptr, int *status);
int fits_get_num_rows(fitsfile *fptr, long *nrows, int *status);
int fits_get_coltype(fitsfile *fptr, int colnum, int *typecode, long *repeat, long *width, int *status);
int fits_read_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem, LONGLONG nelem,
                  void *nulval, void *array, int *anynul, int *status);
void fits_report_error(FILE *stream, int status);
#endif
