/* oracle/stubs/fft_shim.c -- TEST INFRASTRUCTURE ONLY.
 * FP64-internal stand-in for the FFTW3 float API used by the reference's ring_analysis / ring_synthesis
 * (healpix_shtrans.c:549-571, :168-205).  Definitions restated from the FFTW3 manual ("What FFTW Really
 * Computes", 1d real data): r2c  Y_k = sum_j X_j exp(-2 pi i j k / n), k = 0..n/2;  c2r (unnormalised inverse)
 * X_j = sum_{k=0}^{n-1} Y_k exp(+2 pi i j k / n) with Y_{n-k} = conj(Y_k), imaginary parts of Y_0 and Y_{n/2}
 * ignored.  Arithmetic: complex FP64 FFT (radix-2 for powers of two, Bluestein otherwise), result rounded to
 * float exactly once.  In-place use (in == out) is supported because input is copied before any store.
 * Exactly-rounded means ties must not be left to FFT round-off: the outputs whose twiddle factors are rational
 * (r2c bins 0, n/4, n/2 and the real parts of bins n/6, n/3; c2r samples 0, n/4, n/2, 3n/4) are exact sums of floats
 * and land on float rounding ties with probability ~1/n, so they are evaluated by exact (long double) summation. */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fftw3.h"

struct oracle_fft_plan_s { int n; int kind; void *in; void *out; };

typedef struct { double re, im; } cplx;

/* exp(2 pi i num/den) with argument reduction in integers (num taken mod den, then octant folding) */
static cplx unit_root(long num, long den)
{
  num %= den; if (num < 0) num += den;
  /* angle = 2 pi num/den in [0, 2pi); fold to [0, pi/4] */
  long q = (8 * num) / den;            /* octant 0..7 */
  long r8 = 8 * num - q * den;         /* 8*num mod den, angle within octant = (pi/4) * r8/den */
  double a = (M_PI / 4.0) * ((double)r8 / (double)den);
  double c, s;
  if (q & 1) { a = (M_PI / 4.0) - a; }
  c = cos(a); s = sin(a);
  cplx w;
  switch (q) {
    case 0: w.re = c;  w.im = s;  break;
    case 1: w.re = s;  w.im = c;  break;
    case 2: w.re = -s; w.im = c;  break;
    case 3: w.re = -c; w.im = s;  break;
    case 4: w.re = -c; w.im = -s; break;
    case 5: w.re = -s; w.im = -c; break;
    case 6: w.re = s;  w.im = -c; break;
    default: w.re = c; w.im = -s; break;
  }
  return w;
}

/* cached per-size tables */
typedef struct fft_tab {
  int n;             /* transform length this entry serves */
  int m;             /* power-of-two work length (m == n for pow2 n, else >= 2n-1) */
  cplx *tw;          /* m/2 twiddles exp(-2 pi i k/m) */
  cplx *chirp;       /* Bluestein: n values exp(-i pi j^2/n) (forward sign) */
  cplx *chirp_fft;   /* Bluestein: FFT_m of the conjugate-chirp kernel */
  struct fft_tab *next;
} fft_tab;
static fft_tab *g_tabs = NULL;

static void fft_pow2(cplx *a, int m, const cplx *tw, int twm, int inverse)
{
  /* iterative radix-2 decimation in time; tw has twm/2 entries exp(-2 pi i k/twm), twm >= m a multiple */
  int i, j = 0;
  for (i = 1; i < m; ++i) {
    int bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
  }
  for (int len = 2; len <= m; len <<= 1) {
    int half = len >> 1, step = twm / len;
    for (i = 0; i < m; i += len)
      for (int k = 0; k < half; ++k) {
        cplx w = tw[k * step];
        if (inverse) w.im = -w.im;
        cplx u = a[i + k], v = a[i + k + half];
        cplx t; t.re = v.re * w.re - v.im * w.im; t.im = v.re * w.im + v.im * w.re;
        a[i + k].re = u.re + t.re; a[i + k].im = u.im + t.im;
        a[i + k + half].re = u.re - t.re; a[i + k + half].im = u.im - t.im;
      }
  }
}

static fft_tab *get_tab(int n)
{
  fft_tab *t;
  for (t = g_tabs; t; t = t->next) if (t->n == n) return t;
  t = (fft_tab*)calloc(1, sizeof(fft_tab));
  t->n = n;
  int pow2 = (n & (n - 1)) == 0;
  int m = n;
  if (!pow2) { m = 1; while (m < 2 * n - 1) m <<= 1; }
  t->m = m;
  t->tw = (cplx*)malloc(sizeof(cplx) * (size_t)(m / 2 > 0 ? m / 2 : 1));
  for (int k = 0; k < m / 2; ++k) t->tw[k] = unit_root(-(long)k, (long)m);
  if (!pow2) {
    t->chirp = (cplx*)malloc(sizeof(cplx) * (size_t)n);
    t->chirp_fft = (cplx*)calloc((size_t)m, sizeof(cplx));
    for (long j = 0; j < n; ++j) {
      long j2 = (j * j) % (2L * n);
      t->chirp[j] = unit_root(-j2, 2L * n);          /* exp(-i pi j^2/n) */
      cplx c = t->chirp[j]; c.im = -c.im;           /* conj: exp(+i pi j^2/n) */
      t->chirp_fft[j] = c;
      if (j) t->chirp_fft[m - j] = c;
    }
    fft_pow2(t->chirp_fft, m, t->tw, m, 0);
  }
  t->next = g_tabs; g_tabs = t;
  return t;
}

/* forward complex DFT of length n (sign -1), any n >= 1, in place on a[0..n) */
static void dft_forward(cplx *a, int n)
{
  if (n == 1) return;
  fft_tab *t = get_tab(n);
  if (t->m == n) { fft_pow2(a, n, t->tw, n, 0); return; }
  int m = t->m;
  cplx *w = (cplx*)calloc((size_t)m, sizeof(cplx));
  for (int j = 0; j < n; ++j) {
    cplx c = t->chirp[j];
    w[j].re = a[j].re * c.re - a[j].im * c.im;
    w[j].im = a[j].re * c.im + a[j].im * c.re;
  }
  fft_pow2(w, m, t->tw, m, 0);
  for (int k = 0; k < m; ++k) {
    cplx b = t->chirp_fft[k], v = w[k];
    w[k].re = v.re * b.re - v.im * b.im;
    w[k].im = v.re * b.im + v.im * b.re;
  }
  fft_pow2(w, m, t->tw, m, 1);
  double inv = 1.0 / m;
  for (int k = 0; k < n; ++k) {
    cplx c = t->chirp[k];
    double re = w[k].re * inv, im = w[k].im * inv;
    a[k].re = re * c.re - im * c.im;
    a[k].im = re * c.im + im * c.re;
  }
  free(w);
}

void *fftwf_malloc(size_t n) { return malloc(n); }
void fftwf_free(void *p) { free(p); }
void fftwf_cleanup(void) {}

static fftwf_plan mkplan(int n, int kind, void *in, void *out)
{
  fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
  p->n = n; p->kind = kind; p->in = in; p->out = out;
  return p;
}
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags) { (void)flags; return mkplan(n, 0, in, out); }
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags) { (void)flags; return mkplan(n, 1, in, out); }
void fftwf_destroy_plan(fftwf_plan p) { free(p); }

void fftwf_execute(const fftwf_plan p)
{
  int n = p->n;
  cplx *a = (cplx*)malloc(sizeof(cplx) * (size_t)n);
  if (p->kind == 0) {
    const float *x = (const float*)p->in;
    fftwf_complex *y = (fftwf_complex*)p->out;
    for (int j = 0; j < n; ++j) { a[j].re = (double)x[j]; a[j].im = 0.0; }
    dft_forward(a, n);
    {
      /* exact sums for the rational-twiddle bins; T[c] = sum of x_j over j = c (mod 12) */
      long double T[12]; for (int c = 0; c < 12; ++c) T[c] = 0.0L;
      for (int j = 0; j < n; ++j) T[j % 12] += (long double)x[j];
      long double s = 0; for (int c = 0; c < 12; ++c) s += T[c];
      a[0].re = (double)s; a[0].im = 0.0;
      if (n % 2 == 0) {
        s = 0; for (int c = 0; c < 12; ++c) s += (c & 1) ? -T[c] : T[c];
        a[n / 2].re = (double)s; a[n / 2].im = 0.0;
      }
      if (n % 4 == 0 && n >= 4) {
        a[n / 4].re = (double)((T[0] + T[4] + T[8]) - (T[2] + T[6] + T[10]));
        a[n / 4].im = (double)((T[3] + T[7] + T[11]) - (T[1] + T[5] + T[9]));
      }
      if (n % 12 == 0) {
        static const long double c6[6] = {1.0L, 0.5L, -0.5L, -1.0L, -0.5L, 0.5L};
        static const long double c3[3] = {1.0L, -0.5L, -0.5L};
        s = 0; for (int c = 0; c < 12; ++c) s += c6[c % 6] * T[c];
        a[n / 6].re = (double)s;
        s = 0; for (int c = 0; c < 12; ++c) s += c3[c % 3] * T[c];
        a[n / 3].re = (double)s;
      }
    }
    for (int k = 0; k <= n / 2; ++k) { y[k][0] = (float)a[k].re; y[k][1] = (float)a[k].im; }
  } else {
    const fftwf_complex *y = (const fftwf_complex*)p->in;
    float *x = (float*)p->out;
    /* X_j = sum_k Y_k e^{+2 pi i jk/n} = conj( DFT_forward( conj(Y) ) )_j ; build the Hermitian spectrum */
    for (int k = 0; k <= n / 2; ++k) {
      double re = (double)y[k][0], im = (double)y[k][1];
      if (k == 0 || 2 * k == n) im = 0.0;
      a[k].re = re; a[k].im = -im;
      if (k != 0 && 2 * k != n) { a[n - k].re = re; a[n - k].im = im; }
    }
    long double R[4] = {0, 0, 0, 0}, I[4] = {0, 0, 0, 0}, y0 = (long double)y[0][0], yh = (n % 2 == 0) ? (long double)y[n / 2][0] : 0.0L;
    for (int k = 1; 2 * k < n; ++k) { R[k & 3] += (long double)y[k][0]; I[k & 3] += (long double)y[k][1]; }
    dft_forward(a, n);
    for (int j = 0; j < n; ++j) x[j] = (float)a[j].re;
    if (n % 4 == 0 && n >= 4) {
      /* x_j = Y_0 + (-1)^j Y_{n/2} + 2 sum_{0<k<n/2} Re(Y_k e^{2 pi i jk/n}) at j = 0, n/4, n/2, 3n/4 */
      long double sr = ((n / 4) & 1) ? -1.0L : 1.0L;
      x[0] = (float)(double)(y0 + yh + 2.0L * (R[0] + R[1] + R[2] + R[3]));
      x[n / 2] = (float)(double)(y0 + yh + 2.0L * ((R[0] + R[2]) - (R[1] + R[3])));
      x[n / 4] = (float)(double)(y0 + sr * yh + 2.0L * ((R[0] - R[2]) - (I[1] - I[3])));
      x[3 * (n / 4)] = (float)(double)(y0 + sr * yh + 2.0L * ((R[0] - R[2]) + (I[1] - I[3])));
    }
  }
  free(a);
}
