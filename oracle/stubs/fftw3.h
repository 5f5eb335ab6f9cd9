/* oracle/stubs/fftw3.h -- TEST INFRASTRUCTURE ONLY.  The single-precision FFTW3 entry points CALCLENS calls
 * (healpix_shtrans.c:168-205 ring_synthesis, :549-571 ring_analysis).  FFTW3 itself is a third-party
 * dependency that is not under /root/reference and is not installed here (Makefile:159,166 links -lfftw3f,
 * version unpinned).  fft_shim.c restates its documented r2c/c2r definitions with FP64 arithmetic inside and
 * ONE rounding to float on output, so the oracle is "the reference with an exactly-rounded float FFT". */
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <stddef.h>
typedef float fftwf_complex[2];
typedef double fftw_complex[2];
struct oracle_fft_plan_s;
typedef struct oracle_fft_plan_s *fftwf_plan;
typedef struct oracle_fft_plan_s *fftw_plan;
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_MEASURE (0U)
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftwf_cleanup(void);
#endif
