/* oracle/port/calclens_port.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * CPU restatement of the CALCLENS SHTONLY lens-plane hot path, written from the reference's algorithm with each
 * function citing the reference lines it follows (paths relative to the CALCLENS tree).  It is the checker used when
 * the compiled reference (oracle/_ref) cannot be present, and it is itself pinned against oracle/_ref and against
 * the golden vectors in tests/golden/ (tests/test_oracle_*.py).  Plain C, straightforward loops, sized for
 * Nside <= 512.
 *
 * The reference's FFT is third-party (FFTW3 single precision, not under /root/reference).  Its documented r2c/c2r
 * definitions are restated here as a direct O(n^2) DFT with long-double accumulation and twiddles that are exact
 * at multiples of 30/45 degrees, rounded once to float -- an exactly-rounded float DFT, the same contract as
 * oracle/stubs/fft_shim.c.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define PI 3.14159265358979323846264338328
#define PI_2 1.57079632679489661923132169164
#define TWO_OVER_PI 0.63661977236758134307553505349

/* ============================================================ HEALPix geometry ============================== */
typedef struct { long startpix, ringpix, shifted; double cth, sth; } ringinfo;

/* healpix_utils.c:907-953 get_ring_info2 */
static ringinfo ring_info(long ring, long order)
{
  ringinfo q;
  long nside = 1L << order, npix = 12L * (1L << (2 * order)), npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1;
  double fact2 = 4. / npix, fact1 = (nside << 1) * fact2;
  long nr = (ring > 2 * nside) ? 4 * nside - ring : ring;
  if (nr < nside) {
    double tmp = nr * nr * fact2;
    q.cth = 1 - tmp; q.sth = sqrt(tmp * (2 - tmp)); q.ringpix = 4 * nr; q.shifted = 1; q.startpix = 2 * nr * (nr - 1);
  } else {
    q.cth = (2 * nside - nr) * fact1; q.sth = sqrt((1.0 - q.cth) * (1.0 + q.cth)); q.ringpix = 4 * nside;
    q.shifted = (((nr - nside) & 1) == 0) ? 1 : 0; q.startpix = ncap + (nr - nside) * q.ringpix;
  }
  if (nr != ring) { q.cth = -1.0 * q.cth; q.startpix = npix - q.startpix - q.ringpix; }
  return q;
}

static const long jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
static const long jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
static long isqrt_(long i) { return (long)sqrt(((double)i) + 0.5); }   /* healpix_utils.c:47-50 */

/* bit interleave, one bit at a time (healpix_utils.c:234-257 xyf2nest uses 8-bit tables for the same map) */
static long xyf2nest_(long ix, long iy, long face, long order)
{
  long r = 0;
  for (long b = 0; b < order; ++b) r |= (((ix >> b) & 1L) << (2 * b)) | (((iy >> b) & 1L) << (2 * b + 1));
  return (face << (2 * order)) + r;
}
static void nest2xyf_(long pix, long order, long *ix, long *iy, long *face)   /* healpix_utils.c:192-225 */
{
  long npface = 1L << (2 * order), p = pix & (npface - 1), x = 0, y = 0;
  *face = pix >> (2 * order);
  for (long b = 0; b < order; ++b) { x |= ((p >> (2 * b)) & 1L) << b; y |= ((p >> (2 * b + 1)) & 1L) << b; }
  *ix = x; *iy = y;
}
/* healpix_utils.c:271-363 ring2xyf */
static void ring2xyf_(long pix, long order, long *ix, long *iy, long *face)
{
  long nside = 1L << order, npix = 12L * (1L << (2 * order)), npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1, nl2 = 2 * nside, iring, iphi, kshift, nr, f;
  if (pix < ncap) {
    iring = (long)(0.5 * (1 + isqrt_(1 + 2 * pix))); iphi = (pix + 1) - 2 * iring * (iring - 1); kshift = 0; nr = iring;
    f = 0; long t = iphi - 1; if (t >= 2 * iring) { f = 2; t -= 2 * iring; } if (t >= iring) ++f;
  } else if (pix < npix - ncap) {
    long ip = pix - ncap;
    iring = (ip >> (order + 2)) + nside; iphi = (ip & (4 * nside - 1)) + 1; kshift = (iring + nside) & 1; nr = nside;
    long ire = iring - nside + 1, irm = nl2 + 2 - ire;
    long ifm = (iphi - ire / 2 + nside - 1) >> order, ifp = (iphi - irm / 2 + nside - 1) >> order;
    if (ifp == ifm) f = (ifp == 4) ? 4 : ifp + 4; else if (ifp < ifm) f = ifp; else f = ifm + 8;
  } else {
    long ip = npix - pix;
    iring = (long)(0.5 * (1 + isqrt_(2 * ip - 1))); iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1)); kshift = 0; nr = iring;
    iring = 2 * nl2 - iring;
    f = 8; long t = iphi - 1; if (t >= 2 * nr) { f = 10; t -= 2 * nr; } if (t >= nr) ++f;
  }
  long irt = iring - jrll[f] * nside + 1, ipt = 2 * iphi - jpll[f] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  *ix = (ipt - irt) >> 1; *iy = (-(ipt + irt)) >> 1; *face = f;
}
/* healpix_utils.c:365-411 xyf2ring */
static long xyf2ring_(long ix, long iy, long face, long order)
{
  long nside = 1L << order, npix = 12L * (1L << (2 * order)), npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1, nl4 = 4 * nside, jr = jrll[face] * nside - ix - iy - 1, nr, kshift, nb;
  if (jr < nside) { nr = jr; nb = 2 * nr * (nr - 1); kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; nb = npix - 2 * (nr + 1) * nr; kshift = 0; }
  else { nr = nside; nb = ncap + (jr - nside) * nl4; kshift = (jr - nside) & 1; }
  long jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4; else if (jp < 1) jp += nl4;
  return nb + jp - 1;
}
long port_ring2nest(long pix, long order) { long x, y, f; ring2xyf_(pix, order, &x, &y, &f); return xyf2nest_(x, y, f, order); }  /* :420-425 */
long port_nest2ring(long pix, long order) { long x, y, f; nest2xyf_(pix, order, &x, &y, &f); return xyf2ring_(x, y, f, order); }  /* :413-418 */

/* healpix_utils.c:548-622 ang2nest (order 29, then degraded) */
long port_ang2nest(double theta, double phi, long inorder)
{
  const long order = 29, nside = 1L << 29;
  long innside = 1L << inorder, face, ix, iy;
  double z = cos(theta), za = fabs(z), tt = phi;
  long ttl = (long)floor(tt / 2 / PI);
  tt = tt - ((double)ttl) * 2 * PI; tt *= TWO_OVER_PI;
  if (za <= 2.0 / 3.0) {
    double t1 = nside * (0.5 + tt), t2 = nside * (z * 0.75);
    long jp = (long)(t1 - t2), jm = (long)(t1 + t2), ifp = jp >> order, ifm = jm >> order;
    if (ifp == ifm) face = (ifp == 4) ? 4 : ifp + 4; else if (ifp < ifm) face = ifp; else face = ifm + 8;
    ix = jm & (nside - 1); iy = nside - (jp & (nside - 1)) - 1;
  } else {
    long ntt = (long)tt; if (ntt >= 4) ntt = 3;
    double tp = tt - ntt, tmp = nside * sqrt(3 * (1 - za));
    long jp = (long)(tp * tmp), jm = (long)((1.0 - tp) * tmp);
    if (jp >= nside) jp = nside - 1; if (jm >= nside) jm = nside - 1;
    if (z >= 0) { face = ntt; ix = nside - jm - 1; iy = nside - jp - 1; } else { face = ntt + 8; ix = jp; iy = jm; }
  }
  long opix = xyf2nest_(ix, iy, face, order);
  long ip = opix - nside * nside * face, diff = 1L << (2 * (order - inorder));
  return ip / diff + face * innside * innside;
}

/* healpix_utils.c:427-458 nest2peano */
long port_nest2peano(long pix, long order)
{
  static const unsigned char subpix[8][4] = {{0, 1, 3, 2}, {3, 0, 2, 1}, {2, 3, 1, 0}, {1, 2, 0, 3}, {0, 3, 1, 2}, {1, 0, 2, 3}, {2, 1, 3, 0}, {3, 2, 0, 1}};
  static const unsigned char subpath[8][4] = {{4, 0, 6, 0}, {7, 5, 1, 1}, {2, 4, 2, 6}, {3, 3, 7, 5}, {0, 2, 4, 4}, {5, 1, 5, 3}, {6, 6, 0, 2}, {1, 7, 3, 7}};
  static const unsigned char face2path[12] = {2, 5, 2, 5, 3, 6, 3, 6, 2, 3, 2, 3};
  static const unsigned char face2peanoface[12] = {0, 5, 6, 11, 10, 1, 4, 7, 2, 3, 8, 9};
  long face = pix >> (2 * order), result = 0;
  unsigned path = face2path[face];
  for (long shift = 2 * order - 2; shift >= 0; shift -= 2) {
    unsigned s = (unsigned)((pix >> shift) & 3);
    result = (result << 2) | subpix[path][s];
    path = subpath[path][s];
  }
  return result + (((long)face2peanoface[face]) << (2 * order));
}

/* healpix_utils.c:120-131 vec2ang */
static void vec2ang_(const double v[3], double *theta, double *phi)
{
  double norm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (v[0] == 0.0 && v[1] == 0.0) *phi = 0.0; else *phi = atan2(v[1], v[0]);
  if (*phi < 0.0) *phi = *phi + 2.0 * PI;
  *theta = acos(v[2] / norm);
}
/* healpix_utils.c:700-755 nest2ang + nest2vec, :133-141 ang2vec */
void port_nest2vec(long pix, double vec[3], long order)
{
  long nside = 1L << order, npix = 12L * (1L << (2 * order)), nl4 = 4 * nside, ix, iy, f, nr, kshift;
  double fact2 = 4. / npix, fact1 = (nside << 1) * fact2, z;
  nest2xyf_(pix, order, &ix, &iy, &f);
  long jr = (jrll[f] << order) - ix - iy - 1;
  if (jr < nside) { nr = jr; z = 1 - nr * nr * fact2; kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; z = nr * nr * fact2 - 1; kshift = 0; }
  else { nr = nside; z = (2 * nside - jr) * fact1; kshift = (jr - nside) & 1; }
  long jp = (jpll[f] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4; if (jp < 1) jp += nl4;
  double phi = (jp - (kshift + 1) * 0.5) * (PI_2 / nr), theta = acos(z);
  double ct = cos(theta), st = sqrt((1.0 + ct) * (1.0 - ct));
  vec[0] = st * cos(phi); vec[1] = st * sin(phi); vec[2] = ct;
}
/* healpix_utils.c:955-968 ring_above */
static long ring_above_(double z, long order)
{
  long nside = 1L << order; double az = fabs(z);
  if (az > 2.0 / 3.0) { long ir = (long)(nside * sqrt(3 * (1 - az))); return (z > 0) ? ir : 4 * nside - ir - 1; }
  return (long)(nside * (2 - 1.5 * z));
}
/* healpix_utils.c:971-1043 get_interpol */
void port_get_interpol(double theta, double phi, long pix[4], double wgt[4], long order)
{
  long nside = 1L << order, npix = 12L * (1L << (2 * order));
  double z = cos(theta), theta1 = 0.0, theta2 = 0.0, w1, tmp, dphi;
  long ir1 = ring_above_(z, order), ir2 = ir1 + 1, i1, i2;
  for (int s = 0; s < 2; ++s) {
    long ir = s ? ir2 : ir1;
    if ((s == 0 && ir1 > 0) || (s == 1 && ir2 < 4 * nside)) {
      ringinfo q = ring_info(ir, order);
      double th = atan2(q.sth, q.cth);
      if (s) theta2 = th; else theta1 = th;
      dphi = 2.0 * PI / q.ringpix;
      tmp = (phi / dphi - .5 * q.shifted);
      i1 = (tmp < 0) ? ((long)tmp) - 1 : (long)tmp;
      w1 = (phi - (i1 + .5 * q.shifted) * dphi) / dphi;
      i2 = i1 + 1;
      if (i1 < 0) i1 += q.ringpix; if (i2 >= q.ringpix) i2 -= q.ringpix;
      pix[2 * s] = q.startpix + i1; pix[2 * s + 1] = q.startpix + i2; wgt[2 * s] = 1 - w1; wgt[2 * s + 1] = w1;
    }
  }
  if (ir1 == 0) {
    double wt = theta / theta2; wgt[2] *= wt; wgt[3] *= wt;
    double fac = (1 - wt) * 0.25; wgt[0] = fac; wgt[1] = fac; wgt[2] += fac; wgt[3] += fac;
    pix[0] = (pix[2] + 2) % 4; pix[1] = (pix[3] + 2) % 4;
  } else if (ir2 == 4 * nside) {
    double wt = (theta - theta1) / (PI - theta1); wgt[0] *= (1 - wt); wgt[1] *= (1 - wt);
    double fac = wt * 0.25; wgt[0] += fac; wgt[1] += fac; wgt[2] = fac; wgt[3] = fac;
    pix[2] = ((pix[0] + 2) & 3) + npix - 4; pix[3] = ((pix[1] + 2) & 3) + npix - 4;
  } else {
    double wt = (theta - theta1) / (theta2 - theta1);
    wgt[0] *= (1 - wt); wgt[1] *= (1 - wt); wgt[2] *= wt; wgt[3] *= wt;
  }
}

/* ============================================================ exactly-rounded float DFT ===================== */
/* exp(-2 pi i num/den) in long double, exact at multiples of 30 and 45 degrees */
static void unit_root_l(long num, long den, long double *c, long double *s)
{
  num %= den; if (num < 0) num += den;
  if ((24 * num) % den == 0) {   /* multiple of 15 degrees: use exact values where they are rational */
    long k = (24 * num) / den;     /* angle = k * 15 deg */
    static const int cexact[24] = {2, 9, 9, 9, 1, 9, 0, 9, -1, 9, 9, 9, -2, 9, 9, 9, -1, 9, 0, 9, 1, 9, 9, 9};  /* 2cos, 9 = irrational */
    int kc = (int)k, ks = (int)((k + 18) % 24);   /* sin(a) = cos(a - 90deg) = cos(a + 270deg) */
    long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)num / (long double)den;
    *c = (cexact[kc] != 9) ? 0.5L * cexact[kc] : cosl(a);
    *s = (cexact[ks] != 9) ? -0.5L * cexact[ks] : -sinl(a);
    return;
  }
  long double a = 2.0L * 3.141592653589793238462643383279502884L * (long double)num / (long double)den;
  *c = cosl(a); *s = -sinl(a);
}
/* FFTW3 r2c definition: Y_k = sum_j X_j exp(-2 pi i jk/n), k = 0..n/2; one rounding to float.  healpix_shtrans.c:549-571 */
static void r2c_exact(const float *x, long n, float *yre, float *yim)
{
  long double *c = malloc(sizeof(long double) * n), *s = malloc(sizeof(long double) * n);
  for (long t = 0; t < n; ++t) unit_root_l(t, n, &c[t], &s[t]);
  for (long k = 0; k <= n / 2; ++k) {
    long double re = 0, im = 0;
    for (long j = 0; j < n; ++j) { long t = (j * k) % n; re += (long double)x[j] * c[t]; im += (long double)x[j] * s[t]; }
    yre[k] = (float)(double)re; yim[k] = (float)(double)im;
  }
  free(c); free(s);
}
/* FFTW3 c2r: X_j = sum_{k<n} Y_k exp(+2 pi i jk/n), Hermitian Y, imaginary parts of Y_0 and Y_{n/2} ignored.  healpix_shtrans.c:168-205 */
static void c2r_exact(const float *yre, const float *yim, long n, float *x)
{
  long double *c = malloc(sizeof(long double) * n), *s = malloc(sizeof(long double) * n);
  for (long t = 0; t < n; ++t) unit_root_l(t, n, &c[t], &s[t]);   /* c - i s' with s = -sin */
  for (long j = 0; j < n; ++j) {
    long double acc = (long double)yre[0];
    if (n % 2 == 0) acc += ((j & 1) ? -1.0L : 1.0L) * (long double)yre[n / 2];
    for (long k = 1; 2 * k < n; ++k) {
      long t = (j * k) % n;   /* exp(+i a) = c + i sin = c - i s */
      acc += 2.0L * ((long double)yre[k] * c[t] + (long double)yim[k] * s[t]);
    }
    x[j] = (float)(double)acc;
  }
  free(c); free(s);
}

/* ============================================================ lambda_lm generator =========================== */
typedef struct { long lmax; double *cf, *recfac, *mfac, *t1fac, *t2fac; long m_last; } plm_t;
/* healpix_plmgen.c:185-243 plmgen_init */
static plm_t *plm_new(long lmax)
{
  plm_t *p = malloc(sizeof(plm_t));
  p->lmax = lmax; p->m_last = -1;
  p->cf = malloc(sizeof(double) * 15); p->recfac = malloc(sizeof(double) * 2 * (lmax + 1));
  p->mfac = malloc(sizeof(double) * (lmax + 1)); p->t1fac = malloc(sizeof(double) * (lmax + 1)); p->t2fac = malloc(sizeof(double) * (2 * lmax + 1));
  double inv_sqrt4pi = 1.0 / sqrt(4.0 * PI), inv_ln2 = 1.0 / log(2.0);
  for (long m = 0; m < 15; ++m) p->cf[m] = ldexp(1.0, (int)((m - 4) * 90));
  p->mfac[0] = 1;
  for (long m = 1; m <= lmax; ++m) p->mfac[m] = p->mfac[m - 1] * sqrt((2 * m + 1.0) / (2 * m));
  for (long m = 0; m <= lmax; ++m) p->mfac[m] = inv_ln2 * log(inv_sqrt4pi * p->mfac[m]);
  for (long m = 0; m <= lmax; ++m) p->t1fac[m] = sqrt(4.0 * (m + 1) * (m + 1) - 1.0);
  for (long m = 0; m < 2 * lmax + 1; ++m) p->t2fac[m] = 1. / sqrt(m + 1.0);
  return p;
}
static void plm_free(plm_t *p) { free(p->cf); free(p->recfac); free(p->mfac); free(p->t1fac); free(p->t2fac); free(p); }
/* healpix_plmgen.c:73-183 plmgen (the m_crit/cth_crit shortcut only skips values below 1e-30 and is omitted) */
static long plm_gen(plm_t *p, double cth, double sth, long m, double *vec)
{
  const double eps = 1e-30, fsmall = ldexp(1.0, -90), fbig = ldexp(1.0, 90), inv_ln2 = 1.0 / log(2.0), ln2 = log(2.0);
  long lmax = p->lmax, l;
  if (m > 0 && sth == 0) return lmax + 1;
  if (p->m_last != m) {                                        /* :245-260 plmgen_recalc_recfac */
    double f_old = 1.0;
    for (l = m; l <= lmax; ++l) {
      p->recfac[2 * l] = p->t1fac[l] * p->t2fac[l + m] * p->t2fac[l - m];
      p->recfac[2 * l + 1] = p->recfac[2 * l] / f_old;
      f_old = p->recfac[2 * l];
    }
    p->m_last = m;
  }
  double logval = p->mfac[m];
  if (m > 0) logval += m * inv_ln2 * log(sth);
  long scale = (long)((logval / 90) - (-4));
  double corfac = (scale < 0) ? 0.0 : p->cf[scale];
  double lam_prev = 0, lam = exp(ln2 * (logval - (scale + (-4)) * 90));
  if (m & 1) lam = -lam;
  l = m;
  while (1) {
    if (fabs(lam * corfac) > eps) break;
    if (++l > lmax) break;
    double nxt = cth * lam * p->recfac[2 * (l - 1)] - lam_prev * p->recfac[2 * (l - 1) + 1];
    lam_prev = lam; lam = nxt;
    while (fabs(lam) > fbig) { lam_prev *= fsmall; lam *= fsmall; ++scale; corfac = (scale < 0) ? 0. : p->cf[scale]; }
  }
  if (l > lmax) return l;
  long firstl = l;
  lam_prev *= corfac; lam *= corfac;
  for (;; ) {
    vec[l] = lam;
    if (++l > lmax) break;
    double nxt = cth * lam * p->recfac[2 * (l - 1)] - lam_prev * p->recfac[2 * (l - 1) + 1];
    lam_prev = lam; lam = nxt;
  }
  return firstl;
}
long port_plmgen(long lmax, double cth, double sth, long m, double *vec) { plm_t *p = plm_new(lmax); long f = plm_gen(p, cth, sth, m, vec); plm_free(p); return f; }

/* healpix_shtrans.c:533-544 get_lmin_ylm */
static long lmin_ylm(long m, double sth) { long lmin = m, cut = (long)((m - 40) / 1.35 / sth); return cut > lmin ? cut : lmin; }

/* ============================================================ map2alm ======================================= */
/* map2alm_transpose_mpi.c:54-641, single rank.  ringmap: RING-ordered float map; alm m-major. */
void port_map2alm(long order, long lmax, const double *ring_weights, const float *ringmap, double *alm_re, double *alm_im)
{
  long nside = 1L << order, npix = 12L * nside * nside, nrp = 2 * nside, nslot = 4 * nside;
  double quadweight = 4.0 * PI / npix;
  /* g[m][slot], slot = 2*(ring-1) (+1 south) */
  double *gre = calloc((size_t)(lmax + 1) * nslot, sizeof(double)), *gim = calloc((size_t)(lmax + 1) * nslot, sizeof(double));
  float *buf = malloc(sizeof(float) * 4 * nside), *yre = malloc(sizeof(float) * (2 * nside + 1)), *yim = malloc(sizeof(float) * (2 * nside + 1));
  for (long ring = 1; ring <= nrp; ++ring) {
    ringinfo q = ring_info(ring, order);
    double w = ring_weights ? ring_weights[ring - 1] : 0.0;
    w += 1.0; w *= quadweight;                                                             /* :111-124 */
    for (int hemi = 0; hemi < 2; ++hemi) {
      if (hemi && ring == nrp) break;
      long start = hemi ? npix - q.startpix - q.ringpix : q.startpix, n = q.ringpix;
      for (long i = 0; i < n; ++i) buf[i] = (float)(ringmap[start + i] * w);                /* :161-162 */
      r2c_exact(buf, n, yre, yim);
      for (long m = 0; m <= lmax; ++m) {                                                    /* :227-315 */
        long mind = m % n; double vr, vi;
        if (mind > n / 2) { mind = n - mind; vr = yre[mind]; vi = -yim[mind]; } else { vr = yre[mind]; vi = yim[mind]; }
        if (q.shifted) {
          double p0 = cos(m * PI / n), p1 = -sin(m * PI / n);
          double t0 = vr * p0 - vi * p1, t1 = vr * p1 + vi * p0; vr = t0; vi = t1;
        }
        gre[m * nslot + 2 * (ring - 1) + hemi] = vr; gim[m * nslot + 2 * (ring - 1) + hemi] = vi;
      }
    }
  }
  plm_t *pd = plm_new(lmax);
  double *plm = malloc(sizeof(double) * (lmax + 1));
  long lmind = 0;
  for (long m = 0; m <= lmax; ++m) {                                                        /* :430-536 */
    for (long l = m; l <= lmax; ++l) { alm_re[lmind + l - m] = 0.0; alm_im[lmind + l - m] = 0.0; }
    for (long ring = 1; ring <= nrp; ++ring) {
      ringinfo q = ring_info(ring, order);
      if (lmin_ylm(m, q.sth) > lmax) continue;
      long firstl = plm_gen(pd, q.cth, q.sth, m, plm);
      if (firstl > lmax) continue;
      double nr_ = gre[m * nslot + 2 * (ring - 1)], ni = gim[m * nslot + 2 * (ring - 1)];
      double sr = gre[m * nslot + 2 * (ring - 1) + 1], si = gim[m * nslot + 2 * (ring - 1) + 1];
      if (ring < nrp) {
        double sfact = 1.0 - 2.0 * ((firstl + m) % 2);
        for (long l = firstl; l <= lmax; ++l) {
          alm_re[lmind + l - m] += nr_ * plm[l]; alm_im[lmind + l - m] += ni * plm[l];
          double fac1 = sfact * plm[l];
          alm_re[lmind + l - m] += sr * fac1; alm_im[lmind + l - m] += si * fac1;
          sfact = -sfact;
        }
      } else {
        for (long l = firstl; l <= lmax; ++l) { alm_re[lmind + l - m] += nr_ * plm[l]; alm_im[lmind + l - m] += ni * plm[l]; }
      }
    }
    lmind += lmax - m + 1;
  }
  /* NOTE on summation order: the reference adds all paired rings for every m first and the equator ring in a second
   * sweep (:430-498 then :500-534); adding the equator last per m as here is the same order per (l,m). */
  free(plm); plm_free(pd); free(gre); free(gim); free(buf); free(yre); free(yim);
}

/* shtpoissonsolve.c:526-550 */
void port_poisson_filter(long lmax, double *alm_re, double *alm_im)
{
  long i = 0;
  for (long m = 0; m <= lmax; ++m)
    for (long l = m; l <= lmax; ++l, ++i) {
      if (l == 0 && m == 0) { alm_re[i] = 0.0; alm_im[i] = 0.0; }
      else { double f = (double)(-1.0 / ((double)l) / (((double)l) + 1.0)); alm_re[i] *= f; alm_im[i] *= f; }
    }
}

/* ============================================================ alm2allmaps =================================== */
/* alm2allmaps_transpose_mpi.c:53-1240, single rank; maps[k*npix + pix], k in the reference's argument order */
void port_alm2allmaps(long order, long lmax, const double *alm_re, const double *alm_im, float *maps)
{
  long nside = 1L << order, npix = 12L * nside * nside, nrp = 2 * nside, nslot = 4 * nside, NM = lmax + 1;
  /* q[(slot*6 + map)*NM + m] */
  double *qr = calloc((size_t)nslot * 6 * NM, sizeof(double)), *qi = calloc((size_t)nslot * 6 * NM, sizeof(double));
  plm_t *pd = plm_new(lmax);
  double *plm = malloc(sizeof(double) * (lmax + 1));
  long lmind = 0;
  for (long m = 0; m <= lmax; ++m) {
    for (long ring = 1; ring <= nrp; ++ring) {
      ringinfo q = ring_info(ring, order);
      if (lmin_ylm(m, (float)q.sth) > lmax) continue;                                      /* :308 (float cast) */
      long firstl = plm_gen(pd, q.cth, q.sth, m, plm);
      if (firstl > lmax) continue;
      double an[6][2] = {{0}}, as[6][2] = {{0}};
      double sfact = 1.0 - (((firstl + m) % 2) << 1);
      for (long l = firstl; l <= lmax; ++l) {                                               /* :320-349 */
        double rval = alm_re[lmind + l - m] * plm[l], ival = alm_im[lmind + l - m] * plm[l];
        an[0][0] += rval; an[0][1] += ival; as[0][0] += sfact * rval; as[0][1] += sfact * ival;
        rval *= m; ival *= m;
        an[2][0] -= ival; an[2][1] += rval; as[2][0] -= sfact * ival; as[2][1] += sfact * rval;
        rval *= m; ival *= m;
        an[5][0] -= rval; an[5][1] -= ival; as[5][0] -= sfact * rval; as[5][1] -= sfact * ival;
        sfact = -sfact;
      }
      sfact = 1.0 - 2.0 * ((firstl + m) % 2);                                               /* :352-447 */
      double cs = q.cth / q.sth, cs2 = q.cth / q.sth / q.sth, c2s2 = cs * cs, gl1n, gl1s;
      long l = firstl;
      if (l > 0) {
        double ar = alm_re[lmind + l - m], ai = alm_im[lmind + l - m];
        double gn = ((double)l) * cs * plm[l], gs = -sfact * gn;
        an[1][0] += ar * gn; an[1][1] += ai * gn; as[1][0] += ar * gs; as[1][1] += ai * gs;
        double fac1 = ((double)l) * cs, fac2 = ((double)l) * plm[l] * (1.0 + c2s2);
        double qn = fac1 * gn - fac2, qs = -fac1 * gs - sfact * fac2;
        an[3][0] += ar * qn; an[3][1] += ai * qn; as[3][0] += ar * qs; as[3][1] += ai * qs;
        gl1n = gn; gl1s = gs;
        gn *= m; gs *= m;
        an[4][0] -= ai * gn; an[4][1] += ar * gn; as[4][0] -= ai * gs; as[4][1] += ar * gs;
      } else { gl1n = 0.0; gl1s = 0.0; }
      sfact = -sfact;
      for (l = firstl + 1; l <= lmax; ++l) {
        double ar = alm_re[lmind + l - m], ai = alm_im[lmind + l - m];
        double sv = sqrt((2.0 * l + 1.0) / (2.0 * l - 1.0) * ((double)(l * l - m * m)));   /* :247 svec */
        double gn = ((double)l) * cs * plm[l] - sv * plm[l - 1] / q.sth, gs = -sfact * gn;
        an[1][0] += ar * gn; an[1][1] += ai * gn; as[1][0] += ar * gs; as[1][1] += ai * gs;
        double fac1 = ((double)l) * cs, fac2 = ((double)l) * plm[l] * (1.0 + c2s2), fac3 = sv / q.sth;
        double qn = fac1 * gn - fac2 - fac3 * gl1n, qs = -fac1 * gs - sfact * fac2 - fac3 * gl1s;
        double gf = sv * plm[l - 1] * cs2;
        qn += gf; qs += gf * sfact;
        an[3][0] += ar * qn; an[3][1] += ai * qn; as[3][0] += ar * qs; as[3][1] += ai * qs;
        gl1n = gn; gl1s = gs;
        gn *= m; gs *= m;
        an[4][0] -= ai * gn; an[4][1] += ar * gn; as[4][0] -= ai * gs; as[4][1] += ar * gs;
        sfact = -sfact;
      }
      for (int k = 0; k < 6; ++k) {
        qr[((2 * (ring - 1)) * 6 + k) * NM + m] = an[k][0]; qi[((2 * (ring - 1)) * 6 + k) * NM + m] = an[k][1];
        if (ring < nrp) { qr[((2 * (ring - 1) + 1) * 6 + k) * NM + m] = as[k][0]; qi[((2 * (ring - 1) + 1) * 6 + k) * NM + m] = as[k][1]; }
      }
    }
    lmind += lmax - m + 1;
  }
  free(plm); plm_free(pd);
  float *yre = malloc(sizeof(float) * (2 * nside + 1)), *yim = malloc(sizeof(float) * (2 * nside + 1)), *x = malloc(sizeof(float) * 4 * nside);
  for (long ring = 1; ring <= nrp; ++ring) {
    ringinfo q = ring_info(ring, order);
    long n = q.ringpix, nc = n / 2 + 1;
    for (int hemi = 0; hemi < 2; ++hemi) {
      if (hemi && ring == nrp) break;
      long start = hemi ? npix - q.startpix - n : q.startpix, slot = 2 * (ring - 1) + hemi;
      for (int k = 0; k < 6; ++k) {
        for (long i = 0; i < nc; ++i) { yre[i] = 0.0f; yim[i] = 0.0f; }
        const double *br = qr + (slot * 6 + k) * NM, *bi = qi + (slot * 6 + k) * NM;
        for (long m = 0; m <= lmax; ++m) {                                                  /* :826-883 */
          long mp = m % n;
          if (mp < nc) {
            long l = (m - mp) / n; double sk = (q.shifted && (l % 2)) ? -1.0 : 1.0;
            yre[mp] += br[m] * sk; yim[mp] += bi[m] * sk;
          }
          if (m > 0) {
            mp = n - 1 - ((m - 1) % n);
            if (mp < nc) {
              long l = (-m - mp) / n; double sk = (q.shifted && (l % 2)) ? -1.0 : 1.0;
              yre[mp] += br[m] * sk; yim[mp] -= bi[m] * sk;
            }
          }
        }
        if (q.shifted)                                                                      /* healpix_shtrans.c:186-197 */
          for (long mp = 0; mp < nc; ++mp) {
            double c = cos(mp * PI / n), s = sin(mp * PI / n), t0 = (double)yre[mp], t1 = (double)yim[mp];
            yre[mp] = (float)(t0 * c - t1 * s); yim[mp] = (float)(t1 * c + t0 * s);
          }
        c2r_exact(yre, yim, n, x);
        if (k == 2 || k == 5 || k == 4) for (long i = 0; i < n; ++i) x[i] /= q.sth;         /* :1045-1051 */
        if (k == 5) for (long i = 0; i < n; ++i) x[i] /= q.sth;
        memcpy(maps + (size_t)k * npix + start, x, sizeof(float) * n);
      }
      float *mvt = maps + 1 * npix + start, *mvp = maps + 2 * npix + start, *mvtp = maps + 4 * npix + start, *mvpp = maps + 5 * npix + start;
      for (long i = 0; i < n; ++i) {                                                        /* :1097-1147 */
        if (!hemi) { mvtp[i] = (float)(mvtp[i] - q.cth / q.sth * mvp[i]); mvpp[i] = (float)(mvpp[i] + q.cth / q.sth * mvt[i]); }
        else { mvtp[i] = (float)(mvtp[i] + q.cth / q.sth * mvp[i]); mvpp[i] = (float)(mvpp[i] - q.cth / q.sth * mvt[i]); }
      }
    }
  }
  free(yre); free(yim); free(x); free(qr); free(qi);
}

/* ============================================================ rays ========================================== */
typedef struct { long nest; double n[3], beta[3], alpha[2], A[4], Aprev[4], U[4], phi; } ray_t;   /* raytrace.h:284-293 */

/* rot_paratrans.c:101-170 / :179-271 share this angle computation */
static void para_angle(const double *_v, const double *_r, double *cp, double *sp)
{
  double v[3], r[3], ax[3], p[3], re[3], et[3], ep[3];
  double nv = sqrt(_v[0] * _v[0] + _v[1] * _v[1] + _v[2] * _v[2]); v[0] = _v[0] / nv; v[1] = _v[1] / nv; v[2] = _v[2] / nv;
  double nr = sqrt(_r[0] * _r[0] + _r[1] * _r[1] + _r[2] * _r[2]); r[0] = _r[0] / nr; r[1] = _r[1] / nr; r[2] = _r[2] / nr;
  ax[0] = v[1] * r[2] - v[2] * r[1]; ax[1] = v[2] * r[0] - v[0] * r[2]; ax[2] = v[0] * r[1] - v[1] * r[0];
  double ca = v[0] * r[0] + v[1] * r[1] + v[2] * r[2], sa = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  if (sa != 0.0) { ax[0] /= sa; ax[1] /= sa; ax[2] /= sa; } else { ax[0] = 1.0; ax[1] = 0.0; ax[2] = 0.0; }
  p[0] = -v[1]; p[1] = v[0]; p[2] = 0.0;
  double ad = ax[0] * p[0] + ax[1] * p[1] + ax[2] * p[2];                                     /* :78-92 */
  double cx = ax[1] * p[2] - ax[2] * p[1], cy = ax[2] * p[0] - ax[0] * p[2], cz = ax[0] * p[1] - ax[1] * p[0];
  re[0] = p[0] * ca + ax[0] * ad * (1.0 - ca) + cx * sa; re[1] = p[1] * ca + ax[1] * ad * (1.0 - ca) + cy * sa; re[2] = p[2] * ca + ax[2] * ad * (1.0 - ca) + cz * sa;
  ep[0] = -r[1]; ep[1] = r[0]; ep[2] = 0.0;
  et[0] = r[2] * r[0]; et[1] = r[2] * r[1]; et[2] = -1.0 * (r[0] * r[0] + r[1] * r[1]);
  double norm = sqrt((1.0 - r[2]) * (1.0 + r[2]) * (1.0 - v[2]) * (1.0 + v[2]));
  *sp = (re[0] * et[0] + re[1] * et[1] + re[2] * et[2]) / norm;
  *cp = (re[0] * ep[0] + re[1] * ep[1] + re[2] * ep[2]) / norm;
}
static void para_tensor(const double T[2][2], double c, double s, double R[2][2])               /* :251-270 */
{
  double r[2][2] = {{c, -1.0 * s}, {s, c}}, rt[2][2] = {{c, s}, {-1.0 * s, c}}, t1[2][2];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t1[i][j] = T[i][0] * r[0][j] + T[i][1] * r[1][j];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) R[i][j] = rt[i][0] * t1[0][j] + rt[i][1] * t1[1][j];
}

/* shtpoissonsolve.c:1122-1204 shearinterp_comp + caller :666-702; maps RING-ordered (the reference looks the same
 * pixels up in NEST-ordered cells via ring2nest; the pixel centre comes from nest2vec of that NEST index) */
void port_shearinterp(long order, const float *maps, ray_t *rays, long nrays)
{
  long npix = 12L << (2 * order);
  for (long i = 0; i < nrays; ++i) {
    double theta, phi, wgt[4], pot = 0, gt = 0, gp = 0, ti[2][2] = {{0, 0}, {0, 0}};
    long pix[4];
    vec2ang_(rays[i].n, &theta, &phi);
    port_get_interpol(theta, phi, pix, wgt, order);
    for (int k = 0; k < 4; ++k) {
      long nest = port_ring2nest(pix[k], order);
      double vec[3], c, s, T[2][2], R[2][2];
      pot += maps[pix[k]] * wgt[k];
      port_nest2vec(nest, vec, order);
      para_angle(vec, rays[i].n, &c, &s);
      double t0 = maps[1 * npix + pix[k]], t1 = maps[2 * npix + pix[k]];
      gt += (t0 * c + t1 * s) * wgt[k]; gp += (-1.0 * t0 * s + t1 * c) * wgt[k];            /* rot_paratrans.c:168-169 */
      T[0][0] = maps[3 * npix + pix[k]]; T[0][1] = maps[4 * npix + pix[k]]; T[1][0] = T[0][1]; T[1][1] = maps[5 * npix + pix[k]];
      para_tensor(T, c, s, R);
      ti[0][0] += R[0][0] * wgt[k]; ti[0][1] += R[0][1] * wgt[k]; ti[1][0] += R[1][0] * wgt[k]; ti[1][1] += R[1][1] * wgt[k];
    }
    rays[i].phi = pot;
    rays[i].alpha[0] += -1.0 * gt; rays[i].alpha[1] += -1.0 * gp;
    rays[i].U[0] += ti[0][0]; rays[i].U[1] += ti[0][1]; rays[i].U[2] += ti[1][0]; rays[i].U[3] += ti[1][1];
  }
}

/* rayprop.c:18-189 rayprop_sphere, non-BORNAPPRX */
void port_rayprop(ray_t *rays, long nrays, double wp, double wpm1, double wpm2)
{
  for (long i = 0; i < nrays; ++i) {
    ray_t *r = &rays[i];
    double np[3], bp[3], Ap[4];
    double alpha = sqrt(r->alpha[0] * r->alpha[0] + r->alpha[1] * r->alpha[1]);
    if (alpha > 0.0) {
      double ph[3], th[3], a[3], x[3], R[3][3], norm;
      ph[0] = -1.0 * r->n[1]; ph[1] = r->n[0]; ph[2] = 0.0;
      norm = sqrt(ph[0] * ph[0] + ph[1] * ph[1]); ph[0] /= norm; ph[1] /= norm;
      th[0] = r->n[2] * r->n[0]; th[1] = r->n[2] * r->n[1]; th[2] = -1.0 * (r->n[0] * r->n[0] + r->n[1] * r->n[1]);
      norm = sqrt(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]); th[0] /= norm; th[1] /= norm; th[2] /= norm;
      for (int k = 0; k < 3; ++k) a[k] = r->alpha[0] * th[k] + r->alpha[1] * ph[k];
      x[0] = r->n[1] * a[2] - r->n[2] * a[1]; x[1] = r->n[2] * a[0] - r->n[0] * a[2]; x[2] = r->n[0] * a[1] - r->n[1] * a[0];
      norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); x[0] /= norm; x[1] /= norm; x[2] /= norm;
      double sa = sin(alpha), ca = cos(alpha);                                              /* rot_paratrans.c:17-45 */
      for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) R[p][q] = 0.0;
      R[0][0] = ca; R[1][1] = ca; R[2][2] = ca;
      for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) R[p][q] += x[p] * x[q] * (1.0 - ca);
      R[0][1] -= x[2] * sa; R[0][2] += x[1] * sa; R[1][2] -= x[0] * sa; R[1][0] += x[2] * sa; R[2][0] -= x[1] * sa; R[2][1] += x[0] * sa;
      for (int p = 0; p < 3; ++p) { bp[p] = R[p][0] * r->beta[0]; bp[p] += R[p][1] * r->beta[1]; bp[p] += R[p][2] * r->beta[2]; }
      double qa = 1.0, qb = 2.0 * (r->n[0] * bp[0] + r->n[1] * bp[1] + r->n[2] * bp[2]), qc = wpm1 * wpm1 - wp * wp;
      double q = -0.5 * (qb + qb / fabs(qb) * sqrt(qb * qb - 4.0 * qa * qc)), lambda = qc / q;
      if (lambda < 0.0) lambda = q / qa;
      for (int k = 0; k < 3; ++k) np[k] = r->n[k] + bp[k] * lambda;
    } else {
      for (int k = 0; k < 3; ++k) { bp[k] = r->beta[k]; np[k] = r->n[k] / wpm1 * wp; }
    }
    for (int n = 0; n < 2; ++n) for (int m = 0; m < 2; ++m)
      Ap[m + 2 * n] = (1.0 - wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2)) * r->Aprev[m + 2 * n] + (wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2)) * r->A[m + 2 * n]
                      - ((wp - wpm1) / wp) * (r->U[0 + 2 * n] * r->A[m + 2 * 0] + r->U[1 + 2 * n] * r->A[m + 2 * 1]);
    double c, s, T[2][2], RT[2][2];
    para_angle(r->n, np, &c, &s);
    T[0][0] = r->A[0]; T[0][1] = r->A[1]; T[1][0] = r->A[2]; T[1][1] = r->A[3];
    para_tensor(T, c, s, RT);
    r->Aprev[0] = RT[0][0]; r->Aprev[1] = RT[0][1]; r->Aprev[2] = RT[1][0]; r->Aprev[3] = RT[1][1];
    T[0][0] = Ap[0]; T[0][1] = Ap[1]; T[1][0] = Ap[2]; T[1][1] = Ap[3];
    para_tensor(T, c, s, RT);
    r->A[0] = RT[0][0]; r->A[1] = RT[0][1]; r->A[2] = RT[1][0]; r->A[3] = RT[1][1];
    for (int k = 0; k < 3; ++k) { r->n[k] = np[k]; r->beta[k] = bp[k]; }
    double rr = sqrt(r->n[0] * r->n[0] + r->n[1] * r->n[1] + r->n[2] * r->n[2]); rr = wp / rr;
    r->n[0] *= rr; r->n[1] *= rr; r->n[2] *= rr;
  }
}

/* raytrace_utils.c:302-347 init_rays for NEST pixels first..first+n-1 */
/* rayprop.c:40-62, the -DBORNAPPRX build of rayprop_sphere: rays move radially, A recursion without the U A product */
void port_rayprop_born(ray_t *rays, long nrays, double wp, double wpm1, double wpm2)
{
  for (long i = 0; i < nrays; ++i) {
    ray_t *r = &rays[i];
    double Ap[4];
    r->n[0] = r->n[0] / wpm1 * wp; r->n[1] = r->n[1] / wpm1 * wp; r->n[2] = r->n[2] / wpm1 * wp;
    for (int k = 0; k < 4; ++k)
      Ap[k] = (1.0 - wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2)) * r->Aprev[k] + (wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2)) * r->A[k]
              - ((wp - wpm1) / wp) * (r->U[k]);
    for (int k = 0; k < 4; ++k) { r->Aprev[k] = r->A[k]; r->A[k] = Ap[k]; }
    double rr = sqrt(r->n[0] * r->n[0] + r->n[1] * r->n[1] + r->n[2] * r->n[2]);   /* :183-187, both builds */
    rr = wp / rr;
    r->n[0] *= rr; r->n[1] *= rr; r->n[2] *= rr;
  }
}

void port_init_rays(ray_t *rays, long first, long n, long ray_order, double binL_2)
{
  memset(rays, 0, sizeof(ray_t) * n);
  for (long i = 0; i < n; ++i) {
    rays[i].nest = first + i;
    port_nest2vec(first + i, rays[i].beta, ray_order);
    for (int k = 0; k < 3; ++k) rays[i].n[k] = rays[i].beta[k] * binL_2;
    rays[i].A[0] = 1.0; rays[i].A[3] = 1.0; rays[i].Aprev[0] = 1.0; rays[i].Aprev[3] = 1.0;
  }
}
/* rayio.c:300-312: paratrans_ray_curr2obs (rot_paratrans.c:274-302) then rot_ray_ang2radec (:375-411) */
void port_ray_output(ray_t *rays, long nrays, long ray_order)
{
  for (long i = 0; i < nrays; ++i) {
    ray_t *r = &rays[i];
    double obs[3], c, s, T[2][2], R[2][2];
    port_nest2vec(r->nest, obs, ray_order);
    para_angle(r->n, obs, &c, &s);        /* the reference evaluates the same angle for Aprev and for A */
    T[0][0] = r->Aprev[0]; T[0][1] = r->Aprev[1]; T[1][0] = r->Aprev[2]; T[1][1] = r->Aprev[3];
    para_tensor(T, c, s, R);
    r->Aprev[0] = R[0][0]; r->Aprev[1] = R[0][1]; r->Aprev[2] = R[1][0]; r->Aprev[3] = R[1][1];
    T[0][0] = r->A[0]; T[0][1] = r->A[1]; T[1][0] = r->A[2]; T[1][1] = r->A[3];
    para_tensor(T, c, s, R);
    r->A[0] = R[0][0]; r->A[1] = R[0][1]; r->A[2] = R[1][0]; r->A[3] = R[1][1];
    /* (theta, phi) -> (ra, dec): alpha -> (alpha_phi, -alpha_theta); M -> [[M11, -M10], [-M01, M00]] */
    double a0 = r->alpha[0], a1 = r->alpha[1];
    r->alpha[0] = a1; r->alpha[1] = -1.0 * a0;
    double *M[3] = {r->A, r->Aprev, r->U};
    for (int k = 0; k < 3; ++k) {
      double m00 = M[k][0], m01 = M[k][1], m10 = M[k][2], m11 = M[k][3];
      M[k][0] = m11; M[k][2] = -1.0 * m01; M[k][1] = -1.0 * m10; M[k][3] = m00;
    }
  }
}

/* shtpoissonsolve.c:128-150 (NGPSHTDENS): val += (float)(mass/MASS_SCALE) at ang2nest(vec2ang(pos)); RING map out */
void port_deposit_ngp(const float *pos, const float *mass, long nparts, long order, float *ringmap)
{
  for (long k = 0; k < nparts; ++k) {
    double vec[3] = {(double)pos[3 * k], (double)pos[3 * k + 1], (double)pos[3 * k + 2]}, theta, phi;
    vec2ang_(vec, &theta, &phi);
    long nest = port_ang2nest(theta, phi, order);
    ringmap[port_nest2ring(nest, order)] += (float)(mass[k] / 1e10);
  }
}

long port_sizeof_ray(void) { return (long)sizeof(ray_t); }
