/* oracle/stubs/fits_stub.c -- TEST INFRASTRUCTURE ONLY.  Just enough of a FITS reader for the HEALPix ring
 * weight files (first BINTABLE extension, column 1, big-endian IEEE doubles) that read_ring_weights
 * (healpix_shtrans.c:361-423) opens as "<path>/weight_ring_nNNNNN.fits[1]". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fitsio.h"

static long card_long(const char *block, const char *key)
{
  size_t kl = strlen(key);
  for (int c = 0; c < 36; ++c) {
    const char *card = block + 80 * c;
    if (strncmp(card, key, kl) == 0 && card[kl] == ' ' || (strncmp(card, key, kl) == 0 && card[kl] == '=')) {
      const char *eq = memchr(card, '=', 80);
      if (eq) return atol(eq + 1);
    }
  }
  return -1;
}
static int has_end(const char *block)
{
  for (int c = 0; c < 36; ++c) if (strncmp(block + 80 * c, "END     ", 8) == 0) return 1;
  return 0;
}
static int card_tform1(const char *block, long *repeat, char *code)
{
  for (int c = 0; c < 36; ++c) {
    const char *card = block + 80 * c;
    if (strncmp(card, "TFORM1  ", 8) == 0) {
      const char *q = memchr(card, '\'', 80);
      if (!q) return 0;
      ++q;
      long r = 0; int have = 0;
      while (*q >= '0' && *q <= '9') { r = 10 * r + (*q - '0'); ++q; have = 1; }
      *repeat = have ? r : 1; *code = *q;
      return 1;
    }
  }
  return 0;
}

int fits_open_file(fitsfile **fptr, const char *name, int mode, int *status)
{
  (void)mode;
  char path[4096]; strncpy(path, name, sizeof(path) - 1); path[sizeof(path) - 1] = 0;
  char *br = strrchr(path, '['); if (br) *br = 0;
  FILE *fp = fopen(path, "rb");
  if (!fp) { *status = 104; *fptr = NULL; return *status; }
  char block[2880];
  /* skip primary header (+ its data, which is empty for these files) */
  long naxis = -1, datasize = 0; int first = 1;
  for (;;) {
    if (fread(block, 1, 2880, fp) != 2880) { *status = 107; fclose(fp); return *status; }
    if (first) { naxis = card_long(block, "NAXIS"); first = 0; }
    if (has_end(block)) break;
  }
  if (naxis > 0) { *status = 999; fclose(fp); return *status; }
  (void)datasize;
  /* extension header */
  long naxis1 = -1, naxis2 = -1, repeat = 1; char code = 0; int got = 0;
  for (;;) {
    if (fread(block, 1, 2880, fp) != 2880) { *status = 107; fclose(fp); return *status; }
    if (naxis1 < 0) naxis1 = card_long(block, "NAXIS1");
    if (naxis2 < 0) naxis2 = card_long(block, "NAXIS2");
    if (!got) got = card_tform1(block, &repeat, &code);
    if (has_end(block)) break;
  }
  if (!got || (code != 'D' && code != 'E') || naxis1 < 0 || naxis2 < 0) { *status = 998; fclose(fp); return *status; }
  if (code != 'D') { *status = 997; fclose(fp); return *status; }
  fitsfile *f = (fitsfile*)malloc(sizeof(fitsfile));
  f->fp = fp; f->data_start = ftell(fp); f->nrows = naxis2; f->repeat = repeat; f->rowbytes = naxis1;
  *fptr = f; return 0;
}
int fits_close_file(fitsfile *f, int *status) { (void)status; if (f) { fclose(f->fp); free(f); } return 0; }
int fits_get_num_rows(fitsfile *f, long *nrows, int *status) { (void)status; *nrows = f->nrows; return 0; }
int fits_get_coltype(fitsfile *f, int colnum, int *typecode, long *repeat, long *width, int *status)
{ (void)colnum; (void)status; *typecode = TDOUBLE; *repeat = f->repeat; *width = 8; return 0; }
int fits_read_col(fitsfile *f, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem, LONGLONG nelem,
                  void *nulval, void *array, int *anynul, int *status)
{
  (void)datatype; (void)colnum; (void)nulval; (void)firstrow; (void)firstelem;
  double *out = (double*)array; if (anynul) *anynul = 0;
  LONGLONG k = 0;
  for (long row = 0; row < f->nrows && k < nelem; ++row) {
    fseek(f->fp, f->data_start + row * f->rowbytes, SEEK_SET);
    for (long e = 0; e < f->repeat && k < nelem; ++e) {
      unsigned char b[8], r[8];
      if (fread(b, 1, 8, f->fp) != 8) { *status = 108; return *status; }
      for (int i = 0; i < 8; ++i) r[i] = b[7 - i];
      memcpy(&out[k++], r, 8);
    }
  }
  return 0;
}
void fits_report_error(FILE *stream, int status) { if (status) fprintf(stream, "fits_stub: status %d\n", status); }
