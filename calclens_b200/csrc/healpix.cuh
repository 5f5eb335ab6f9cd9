// calclens_b200/csrc/healpix.cuh
// HEALPix geometry and pixel indexing for host and device.  Integer results must be bit-exact with the
// reference (BASELINE north_star), so every floating-point expression below keeps the reference's operand
// order and rounding points; the reference lines are cited per function (paths relative to the CALCLENS tree).
// Written against the HEALPix definitions (Gorski et al. 2005): ring scheme, NESTED scheme (bit interleave of
// face-local x,y), 12 base faces with (jrll, jpll) = ring / phi offsets of the face corners.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CLB_HD __host__ __device__ __forceinline__
#else
#define CLB_HD static inline
#endif

namespace clb {

#ifndef CLB_PI
#define CLB_PI 3.14159265358979323846264338328      // M_PI as the reference's libm/gsl header spells it
#define CLB_PI_2 1.57079632679489661923132169164    // M_PI_2
#define CLB_2_PI 0.63661977236758134307553505349    // M_2_PI
#endif

struct RingInfo {
  long startpix;   // RING index of the first pixel
  long ringpix;    // pixels in the ring
  double costheta;
  double sintheta;
  long shifted;    // 1: pixel centres at (j+1/2) dphi
};

// ring = 1 .. 4*Nside-1 counted from the north pole.            [healpix_utils.c:907-953 get_ring_info2]
CLB_HD RingInfo ring_info(long ring, long order)
{
  RingInfo ri;
  long nside = 1L << order;
  long npix = 12L * (1L << (2 * order));
  long npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1;
  double fact2 = 4. / npix;
  double fact1 = (nside << 1) * fact2;
  long northring = (ring > 2 * nside) ? 4 * nside - ring : ring;
  if (northring < nside) {
    double tmp = northring * northring * fact2;
    ri.costheta = 1 - tmp;
    ri.sintheta = sqrt(tmp * (2 - tmp));
    ri.ringpix = 4 * northring;
    ri.shifted = 1;
    ri.startpix = 2 * northring * (northring - 1);
  } else {
    ri.costheta = (2 * nside - northring) * fact1;
    ri.sintheta = sqrt((1.0 - ri.costheta) * (1.0 + ri.costheta));
    ri.ringpix = 4 * nside;
    ri.shifted = (((northring - nside) & 1) == 0) ? 1 : 0;
    ri.startpix = ncap + (northring - nside) * ri.ringpix;
  }
  if (northring != ring) {
    ri.costheta = -1.0 * ri.costheta;
    ri.startpix = npix - ri.startpix - ri.ringpix;
  }
  return ri;
}

// integer sqrt as the reference takes it                               [healpix_utils.c:47-50 isqrt]
CLB_HD long isqrt_ref(long i) { return (long)sqrt(((double)i) + 0.5); }

// spread the low 32 bits of v onto the even bit positions (the reference's utab look-ups, healpix_utils.c:234-257)
CLB_HD uint64_t spread_bits(uint64_t v)
{
  v &= 0xffffffffull;
  v = (v | (v << 16)) & 0x0000ffff0000ffffull;
  v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
  v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
  v = (v | (v << 2)) & 0x3333333333333333ull;
  v = (v | (v << 1)) & 0x5555555555555555ull;
  return v;
}
// inverse: gather the even bit positions                               (ctab look-ups, healpix_utils.c:192-225)
CLB_HD uint64_t compress_bits(uint64_t v)
{
  v &= 0x5555555555555555ull;
  v = (v | (v >> 1)) & 0x3333333333333333ull;
  v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
  v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
  v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
  v = (v | (v >> 16)) & 0x00000000ffffffffull;
  return v;
}

CLB_HD long xyf2nest(long ix, long iy, long face, long order)
{
  return (face << (2 * order)) + (long)(spread_bits((uint64_t)ix) | (spread_bits((uint64_t)iy) << 1));
}
CLB_HD void nest2xyf(long pix, long order, long &ix, long &iy, long &face)
{
  long npface = 1L << (2 * order);
  face = pix >> (2 * order);
  uint64_t p = (uint64_t)(pix & (npface - 1));
  ix = (long)compress_bits(p);
  iy = (long)compress_bits(p >> 1);
}

#define CLB_JRLL(f) ((long)(2 + ((f) >> 2)))                                   // {2,2,2,2,3,3,3,3,4,4,4,4}
#define CLB_JPLL(f) ((long)((((f) >> 2) == 1) ? 2 * ((f)&3) : 2 * ((f)&3) + 1)) // {1,3,5,7,0,2,4,6,1,3,5,7}

// RING index -> (ix, iy, face)                                         [healpix_utils.c:271-363 ring2xyf]
CLB_HD void ring2xyf(long pix, long order, long &ix, long &iy, long &face)
{
  long nside = 1L << order;
  long npix = 12L * (1L << (2 * order));
  long npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1;
  long nl2 = 2 * nside;
  long iring, iphi, kshift, nr;
  if (pix < ncap) {
    iring = (long)(0.5 * (1 + isqrt_ref(1 + 2 * pix)));
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0;
    nr = iring;
    face = 0;
    long tmp = iphi - 1;
    if (tmp >= (2 * iring)) { face = 2; tmp -= 2 * iring; }
    if (tmp >= iring) face = face + 1;
  } else if (pix < (npix - ncap)) {
    long ip = pix - ncap;
    iring = (ip >> (order + 2)) + nside;
    iphi = (ip & (4 * nside - 1)) + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    long ire = iring - nside + 1;
    long irm = nl2 + 2 - ire;
    long ifm = (iphi - ire / 2 + nside - 1) >> order;
    long ifp = (iphi - irm / 2 + nside - 1) >> order;
    if (ifp == ifm) face = (ifp == 4) ? 4 : ifp + 4;
    else if (ifp < ifm) face = ifp;
    else face = ifm + 8;
  } else {
    long ip = npix - pix;
    iring = (long)(0.5 * (1 + isqrt_ref(2 * ip - 1)));
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0;
    nr = iring;
    iring = 2 * nl2 - iring;
    face = 8;
    long tmp = iphi - 1;
    if (tmp >= (2 * nr)) { face = 10; tmp -= 2 * nr; }
    if (tmp >= nr) face = face + 1;
  }
  long irt = iring - (CLB_JRLL(face) * nside) + 1;
  long ipt = 2 * iphi - CLB_JPLL(face) * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  ix = (ipt - irt) >> 1;
  iy = (-(ipt + irt)) >> 1;
}

// (ix, iy, face) -> RING index                                         [healpix_utils.c:365-411 xyf2ring]
CLB_HD long xyf2ring(long ix, long iy, long face, long order)
{
  long nside = 1L << order;
  long npix = 12L * (1L << (2 * order));
  long npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1;
  long nl4 = 4 * nside;
  long jr = (CLB_JRLL(face) * nside) - ix - iy - 1;
  long nr, kshift, n_before;
  if (jr < nside) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
  else if (jr > 3 * nside) { nr = nl4 - jr; n_before = npix - 2 * (nr + 1) * nr; kshift = 0; }
  else { nr = nside; n_before = ncap + (jr - nside) * nl4; kshift = (jr - nside) & 1; }
  long jp = (CLB_JPLL(face) * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4) jp -= nl4;
  else if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}

CLB_HD long ring2nest(long pix, long order)                            // [healpix_utils.c:420-425]
{
  long ix, iy, f;
  ring2xyf(pix, order, ix, iy, f);
  return xyf2nest(ix, iy, f, order);
}
CLB_HD long nest2ring(long pix, long order)                            // [healpix_utils.c:413-418]
{
  long ix, iy, f;
  nest2xyf(pix, order, ix, iy, f);
  return xyf2ring(ix, iy, f, order);
}

// (theta, phi) -> NESTED index, evaluated at order 29 and degraded.    [healpix_utils.c:548-622 ang2nest]
CLB_HD long ang2nest(double theta, double phi, long inorder)
{
  const long order = 29;
  const long nside = 1L << order;
  long innside = 1L << inorder;
  double z = cos(theta);
  double za = fabs(z);
  double tt = phi;
  long tt_long = (long)(floor(tt / 2 / CLB_PI));
  tt = tt - ((double)(tt_long)) * 2 * CLB_PI;
  tt *= CLB_2_PI;
  long face, ix, iy;
  if (za <= 2.0 / 3.0) {
    double temp1 = nside * (0.5 + tt);
    double temp2 = nside * (z * 0.75);
    long jp = (long)(temp1 - temp2);
    long jm = (long)(temp1 + temp2);
    long ifp = jp >> order;
    long ifm = jm >> order;
    if (ifp == ifm) face = (ifp == 4) ? 4 : ifp + 4;
    else if (ifp < ifm) face = ifp;
    else face = ifm + 8;
    ix = jm & (nside - 1);
    iy = nside - (jp & (nside - 1)) - 1;
  } else {
    long ntt = (long)(tt);
    if (ntt >= 4) ntt = 3;
    double tp = tt - ntt;
    double tmp = nside * sqrt(3 * (1 - za));
    long jp = (long)(tp * tmp);
    long jm = (long)((1.0 - tp) * tmp);
    if (jp >= nside) jp = nside - 1;
    if (jm >= nside) jm = nside - 1;
    if (z >= 0) { face = ntt; ix = nside - jm - 1; iy = nside - jp - 1; }
    else { face = ntt + 8; ix = jp; iy = jm; }
  }
  long opix = xyf2nest(ix, iy, face, order);
  long ip = opix - nside * nside * face;
  long difffac = 1L << (2 * (order - inorder));
  ip = ip / difffac;
  return ip + face * innside * innside;
}

// unit-sphere angles of a vector                                       [healpix_utils.c:120-131 vec2ang]
CLB_HD void vec2ang(const double vec[3], double &theta, double &phi)
{
  double norm = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  if (vec[0] == 0.0 && vec[1] == 0.0) phi = 0.0;
  else phi = atan2(vec[1], vec[0]);
  if (phi < 0.0) phi = phi + 2.0 * CLB_PI;
  theta = acos(vec[2] / norm);
}

// pixel-centre unit vector of a RING pixel.  The reference obtains it as nest2vec(ring2nest(pix)), i.e.
// nest2ang (z, phi from (jr, jp, nr, kshift)) followed by ang2vec.  (jr, jp) are functions of the ring position
// alone, so they are taken from the ring index directly and fed to the same expressions.
//                                        [healpix_utils.c:700-755 nest2ang/nest2vec, :133-141 ang2vec]
CLB_HD void ringpix2zphi(long pix, long order, double &z, double &phi)
{
  long nside = 1L << order;
  long npix = 12L * (1L << (2 * order));
  long npface = 1L << (2 * order);
  long ncap = (npface - nside) << 1;
  double fact2 = 4. / npix;
  double fact1 = (nside << 1) * fact2;
  long nr, kshift, jp;
  if (pix < ncap) {
    long iring = (long)(0.5 * (1 + isqrt_ref(1 + 2 * pix)));
    jp = (pix + 1) - 2 * iring * (iring - 1);
    nr = iring; kshift = 0;
    z = 1 - nr * nr * fact2;
  } else if (pix < (npix - ncap)) {
    long ip = pix - ncap;
    long jr = (ip >> (order + 2)) + nside;
    jp = (ip & (4 * nside - 1)) + 1;
    nr = nside; kshift = (jr - nside) & 1;
    z = (2 * nside - jr) * fact1;
  } else {
    long ip = npix - pix;
    long iring = (long)(0.5 * (1 + isqrt_ref(2 * ip - 1)));
    jp = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    nr = iring; kshift = 0;
    z = nr * nr * fact2 - 1;
  }
  phi = (jp - (kshift + 1) * 0.5) * (CLB_PI_2 / nr);
}
CLB_HD void zphi2vec(double z, double phi, double vec[3])
{
  // nest2ang returns theta = acos(z) and ang2vec takes costheta = cos(theta) again: keep that round trip
  double theta = acos(z);
  double costheta = cos(theta);
  double sintheta = sqrt((1.0 + costheta) * (1.0 - costheta));
  vec[0] = sintheta * cos(phi);
  vec[1] = sintheta * sin(phi);
  vec[2] = costheta;
}
CLB_HD void nest2vec(long pix, long order, double vec[3])
{
  double z, phi;
  ringpix2zphi(nest2ring(pix, order), order, z, phi);
  zphi2vec(z, phi, vec);
}

CLB_HD long ring_above(double z, long order)                           // [healpix_utils.c:955-968]
{
  long nside = 1L << order;
  double az = fabs(z);
  if (az > 2.0 / 3.0) {
    long iring = (long)(nside * sqrt(3 * (1 - az)));
    return (z > 0) ? iring : 4 * nside - iring - 1;
  }
  return (long)(nside * (2 - 1.5 * z));
}

// bilinear interpolation stencil: 4 RING pixels + weights             [healpix_utils.c:971-1043 get_interpol]
CLB_HD void get_interpol(double theta, double phi, long pix[4], double wgt[4], long order)
{
  long nside = 1L << order;
  long npix = 12L * (1L << (2 * order));
  double z = cos(theta);
  long ir1 = ring_above(z, order);
  long ir2 = ir1 + 1;
  double theta1 = 0.0, theta2 = 0.0, w1, tmp, dphi;
  long i1, i2;
  if (ir1 > 0) {
    RingInfo ri = ring_info(ir1, order);
    theta1 = atan2(ri.sintheta, ri.costheta);
    dphi = 2.0 * CLB_PI / ri.ringpix;
    tmp = (phi / dphi - .5 * ri.shifted);
    i1 = (tmp < 0) ? ((long)(tmp)) - 1 : (long)(tmp);
    w1 = (phi - (i1 + .5 * ri.shifted) * dphi) / dphi;
    i2 = i1 + 1;
    if (i1 < 0) i1 += ri.ringpix;
    if (i2 >= ri.ringpix) i2 -= ri.ringpix;
    pix[0] = ri.startpix + i1; pix[1] = ri.startpix + i2;
    wgt[0] = 1 - w1; wgt[1] = w1;
  }
  if (ir2 < (4 * nside)) {
    RingInfo ri = ring_info(ir2, order);
    theta2 = atan2(ri.sintheta, ri.costheta);
    dphi = 2.0 * CLB_PI / ri.ringpix;
    tmp = (phi / dphi - .5 * ri.shifted);
    i1 = (tmp < 0) ? ((long)(tmp)) - 1 : (long)(tmp);
    w1 = (phi - (i1 + .5 * ri.shifted) * dphi) / dphi;
    i2 = i1 + 1;
    if (i1 < 0) i1 += ri.ringpix;
    if (i2 >= ri.ringpix) i2 -= ri.ringpix;
    pix[2] = ri.startpix + i1; pix[3] = ri.startpix + i2;
    wgt[2] = 1 - w1; wgt[3] = w1;
  }
  if (ir1 == 0) {
    double wtheta = theta / theta2;
    wgt[2] *= wtheta; wgt[3] *= wtheta;
    double fac = (1 - wtheta) * 0.25;
    wgt[0] = fac; wgt[1] = fac; wgt[2] += fac; wgt[3] += fac;
    pix[0] = (pix[2] + 2) % 4;
    pix[1] = (pix[3] + 2) % 4;
  } else if (ir2 == 4 * nside) {
    double wtheta = (theta - theta1) / (CLB_PI - theta1);
    wgt[0] *= (1 - wtheta); wgt[1] *= (1 - wtheta);
    double fac = wtheta * 0.25;
    wgt[0] += fac; wgt[1] += fac; wgt[2] = fac; wgt[3] = fac;
    pix[2] = ((pix[0] + 2) & 3) + npix - 4;
    pix[3] = ((pix[1] + 2) & 3) + npix - 4;
  } else {
    double wtheta = (theta - theta1) / (theta2 - theta1);
    wgt[0] *= (1 - wtheta); wgt[1] *= (1 - wtheta);
    wgt[2] *= wtheta; wgt[3] *= wtheta;
  }
}

// Peano-Hilbert <-> NESTED (domain decomposition order of the reference) [healpix_utils.c:427-489]
// Tables are the HEALPix C++ library's (healpix_base.cc): sub-pixel permutation and next-path per Hilbert state.
CLB_HD long nest2peano(long pix, long order)
{
  const unsigned char subpix[8][4] = {{0, 1, 3, 2}, {3, 0, 2, 1}, {2, 3, 1, 0}, {1, 2, 0, 3},
                                      {0, 3, 1, 2}, {1, 0, 2, 3}, {2, 1, 3, 0}, {3, 2, 0, 1}};
  const unsigned char subpath[8][4] = {{4, 0, 6, 0}, {7, 5, 1, 1}, {2, 4, 2, 6}, {3, 3, 7, 5},
                                       {0, 2, 4, 4}, {5, 1, 5, 3}, {6, 6, 0, 2}, {1, 7, 3, 7}};
  const unsigned char face2path[12] = {2, 5, 2, 5, 3, 6, 3, 6, 2, 3, 2, 3};
  const unsigned char face2peanoface[12] = {0, 5, 6, 11, 10, 1, 4, 7, 2, 3, 8, 9};
  long face = pix >> (2 * order);
  unsigned path = face2path[face];
  long result = 0;
  for (long shift = 2 * order - 2; shift >= 0; shift -= 2) {
    unsigned spix = (unsigned)((pix >> shift) & 0x3);
    result <<= 2;
    result |= subpix[path][spix];
    path = subpath[path][spix];
  }
  return result + (((long)face2peanoface[face]) << (2 * order));
}

}  // namespace clb
