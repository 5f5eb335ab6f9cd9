// calclens_b200/csrc/raymath.cuh
// Per-ray arithmetic of the lens-plane step, host and device: parallel transport of tangent vectors/tensors
// along great circles, 4-point interpolation of the derivative maps at the ray position, and the plane-to-plane
// propagation (deflect beta, advance n, A-matrix recursion).  All FP64; the translation unit that includes this
// for the device is compiled with -fmad=false so that products and sums round exactly like the reference's
// plain C (oracle built without FP contraction).
#pragma once
#include "healpix.cuh"

namespace clb {

// HEALPixRay, 176 bytes, identical field order to raytrace.h:284-293 so host arrays can be copied verbatim.
struct Ray {
  long nest;
  double n[3];
  double beta[3];
  double alpha[2];
  double A[4];
  double Aprev[4];
  double U[4];
  double phi;
};
static_assert(sizeof(Ray) == 176, "Ray must match HEALPixRay");

// Rodrigues rotation of vec about a unit axis by (cos, sin)            [rot_paratrans.c:78-92]
CLB_HD void rot_vec_axis_trig(const double vec[3], double rvec[3], const double axis[3], double cosangle, double sinangle)
{
  double axisdotvec = axis[0] * vec[0] + axis[1] * vec[1] + axis[2] * vec[2];
  double c0 = axis[1] * vec[2] - axis[2] * vec[1];
  double c1 = axis[2] * vec[0] - axis[0] * vec[2];
  double c2 = axis[0] * vec[1] - axis[1] * vec[0];
  rvec[0] = vec[0] * cosangle + axis[0] * axisdotvec * (1.0 - cosangle) + c0 * sinangle;
  rvec[1] = vec[1] * cosangle + axis[1] * axisdotvec * (1.0 - cosangle) + c1 * sinangle;
  rvec[2] = vec[2] * cosangle + axis[2] * axisdotvec * (1.0 - cosangle) + c2 * sinangle;
}

// rotation angle psi of the (e_theta, e_phi) frame when transported from _vec to _rvec along their great circle.
// Shared front half of paratrans_tangvec / paratrans_tangtensor.        [rot_paratrans.c:101-150, :179-233]
CLB_HD void paratrans_angle(const double _vec[3], const double _rvec[3], double &cospsi, double &sinpsi)
{
  double vec[3], rvec[3], axis[3], p[3], rephi[3], ephi[3], etheta[3];
  double norm_vec = sqrt(_vec[0] * _vec[0] + _vec[1] * _vec[1] + _vec[2] * _vec[2]);
  vec[0] = _vec[0] / norm_vec; vec[1] = _vec[1] / norm_vec; vec[2] = _vec[2] / norm_vec;
  double norm_rvec = sqrt(_rvec[0] * _rvec[0] + _rvec[1] * _rvec[1] + _rvec[2] * _rvec[2]);
  rvec[0] = _rvec[0] / norm_rvec; rvec[1] = _rvec[1] / norm_rvec; rvec[2] = _rvec[2] / norm_rvec;
  axis[0] = vec[1] * rvec[2] - vec[2] * rvec[1];
  axis[1] = vec[2] * rvec[0] - vec[0] * rvec[2];
  axis[2] = vec[0] * rvec[1] - vec[1] * rvec[0];
  double cosangle = vec[0] * rvec[0] + vec[1] * rvec[1] + vec[2] * rvec[2];
  double sinangle = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  if (sinangle != 0.0) { axis[0] /= sinangle; axis[1] /= sinangle; axis[2] /= sinangle; }
  else { axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0; }
  p[0] = -vec[1]; p[1] = vec[0]; p[2] = 0.0;
  rot_vec_axis_trig(p, rephi, axis, cosangle, sinangle);
  ephi[0] = -rvec[1]; ephi[1] = rvec[0]; ephi[2] = 0.0;
  etheta[0] = rvec[2] * rvec[0]; etheta[1] = rvec[2] * rvec[1];
  etheta[2] = -1.0 * (rvec[0] * rvec[0] + rvec[1] * rvec[1]);
  double norm = sqrt((1.0 - rvec[2]) * (1.0 + rvec[2]) * (1.0 - vec[2]) * (1.0 + vec[2]));
  sinpsi = (rephi[0] * etheta[0] + rephi[1] * etheta[1] + rephi[2] * etheta[2]) / norm;
  cospsi = (rephi[0] * ephi[0] + rephi[1] * ephi[1] + rephi[2] * ephi[2]) / norm;
}

// T' = R^T T R with R = [[c, -s], [s, c]]                              [rot_paratrans.c:251-270]
CLB_HD void transport_tensor(const double T[2][2], double c, double s, double RT[2][2])
{
  double r[2][2] = {{c, -1.0 * s}, {s, c}};
  double rt[2][2] = {{r[0][0], r[1][0]}, {r[0][1], r[1][1]}};
  double t1[2][2];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) t1[i][j] = T[i][0] * r[0][j] + T[i][1] * r[1][j];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) RT[i][j] = rt[i][0] * t1[0][j] + rt[i][1] * t1[1][j];
}

// ---------------------------------------------------------------------------------------------------------------
// Device path of the interpolation (shtpoissonsolve.c:1122-1204 shearinterp_comp, :666-702 caller) with the per-ring
// quantities tabulated once per map resolution and the transport angle evaluated without normalising the rotation axis:
//   R e = e cos(a) + (u x e) + u (u.e) / (1 + cos(a)),   u = v x v'  (|u| = sin(a)),
// which is Rodrigues' formula with axis u/|u| (rot_paratrans.c:78-92) after cancelling |u|.  Pixel indices and
// interpolation weights are formed by exactly the reference's expressions (they must be bit-exact); the transported
// quantities agree with the reference (oracle) to ~1e-15.
// ---------------------------------------------------------------------------------------------------------------
struct RingTab {      // one entry per ring 1 .. 4*Nside-1 (entry 0 unused)
  double theta;       // atan2(sin, cos) of the ring as get_interpol takes it           [healpix_utils.c:986,1001]
  double cz, sz;      // cos/sin(theta) of the pixel centres as nest2ang+ang2vec give them [healpix_utils.c:133-141,700-748]
  double inv_sz;
  double cd, sd;      // cos/sin of the pixel spacing pi/2/nr
  double dphi;        // 2 pi / ringpix exactly as get_interpol forms it           [healpix_utils.c:987,1002]
  double inv_dphi;    // 1 / dphi (weights only; the pixel index keeps the reference's division)
  double inv_dtheta;  // 1 / (theta(ring+1) - theta(ring)) (weights only)
  double inv_ringpix; // 1 / ringpix (azimuth of pixel centres in units of pi)
};

// transport angle from unit vector v (with phi-hat numerator p = (-vy, vx, 0)) to unit vector r;
// inv_norm = 1 / (sin(theta_v) sin(theta_r))
CLB_HD void paratrans_angle_unit(const double v[3], const double r[3], double inv_norm, double &cospsi, double &sinpsi)
{
  // (this translation unit is built with -fmad=false for the bit-exact index arithmetic; the transport math, which has
  // no such requirement, asks for its fused multiply-adds explicitly)
  const double ux = fma(v[1], r[2], -(v[2] * r[1])), uy = fma(v[2], r[0], -(v[0] * r[2])), uz = fma(v[0], r[1], -(v[1] * r[0]));
  const double c = fma(v[0], r[0], fma(v[1], r[1], v[2] * r[2]));
  // 1/(1 + c): the two points are neighbours (c = 1 - eps, eps ~ 1e-8 .. 1e-3), so a short series in eps/2 replaces
  // the division (truncation error (eps/2)^5); far-apart points take the division
  const double eps = 1.0 - c;
  double inv1c;
  if (eps < 1e-2) { const double h = 0.5 * eps; inv1c = 0.5 * (1.0 + h * (1.0 + h * (1.0 + h * (1.0 + h)))); }
  else inv1c = 1.0 / (1.0 + c);
  const double up = (uy * v[0] - ux * v[1]) * inv1c;
  const double e0 = fma(ux, up, -fma(c, v[1], uz * v[0]));
  const double e1 = fma(uy, up, fma(c, v[0], -(uz * v[1])));
  const double e2 = fma(uz, up, fma(ux, v[0], uy * v[1]));
  const double rxy2 = fma(r[0], r[0], r[1] * r[1]);
  sinpsi = fma(r[2], fma(e0, r[0], e1 * r[1]), -(e2 * rxy2)) * inv_norm;
  cospsi = fma(e1, r[0], -(e0 * r[1])) * inv_norm;
}

#if defined(__CUDACC__)
// get_interpol (healpix_utils.c:971-1043) with the two ring colatitudes read from the table; additionally returns
// the two rings the stencil lives on.  Every expression that feeds an index or a weight is the reference's.
__device__ __forceinline__ void get_interpol_tab(double theta, double phi, long pix[4], double wgt[4], long order,
                                                 const RingTab *__restrict__ tab, long &ringA, long &ringB)
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
    theta1 = tab[ir1].theta;
    dphi = tab[ir1].dphi;
    tmp = (phi / dphi - .5 * ri.shifted);                          // index: the reference's division, bit for bit
    i1 = (tmp < 0) ? ((long)(tmp)) - 1 : (long)(tmp);
    w1 = (phi - (i1 + .5 * ri.shifted) * dphi) * tab[ir1].inv_dphi;   // weight: reciprocal (differs by <= 1 ulp)
    i2 = i1 + 1;
    if (i1 < 0) i1 += ri.ringpix;
    if (i2 >= ri.ringpix) i2 -= ri.ringpix;
    pix[0] = ri.startpix + i1; pix[1] = ri.startpix + i2;
    wgt[0] = 1 - w1; wgt[1] = w1;
  }
  if (ir2 < (4 * nside)) {
    RingInfo ri = ring_info(ir2, order);
    theta2 = tab[ir2].theta;
    dphi = tab[ir2].dphi;
    tmp = (phi / dphi - .5 * ri.shifted);
    i1 = (tmp < 0) ? ((long)(tmp)) - 1 : (long)(tmp);
    w1 = (phi - (i1 + .5 * ri.shifted) * dphi) * tab[ir2].inv_dphi;
    i2 = i1 + 1;
    if (i1 < 0) i1 += ri.ringpix;
    if (i2 >= ri.ringpix) i2 -= ri.ringpix;
    pix[2] = ri.startpix + i1; pix[3] = ri.startpix + i2;
    wgt[2] = 1 - w1; wgt[3] = w1;
  }
  ringA = ir1; ringB = ir2;
  if (ir1 == 0) {
    double wtheta = theta / theta2;
    wgt[2] *= wtheta; wgt[3] *= wtheta;
    double fac = (1 - wtheta) * 0.25;
    wgt[0] = fac; wgt[1] = fac; wgt[2] += fac; wgt[3] += fac;
    pix[0] = (pix[2] + 2) % 4;
    pix[1] = (pix[3] + 2) % 4;
    ringA = 1;
  } else if (ir2 == 4 * nside) {
    double wtheta = (theta - theta1) / (CLB_PI - theta1);
    wgt[0] *= (1 - wtheta); wgt[1] *= (1 - wtheta);
    double fac = wtheta * 0.25;
    wgt[0] += fac; wgt[1] += fac; wgt[2] = fac; wgt[3] = fac;
    pix[2] = ((pix[0] + 2) & 3) + npix - 4;
    pix[3] = ((pix[1] + 2) & 3) + npix - 4;
    ringB = 4 * nside - 1;
  } else {
    double wtheta = (theta - theta1) * tab[ir1].inv_dtheta;
    wgt[0] *= (1 - wtheta); wgt[1] *= (1 - wtheta);
    wgt[2] *= wtheta; wgt[3] *= wtheta;
  }
}

__device__ __forceinline__ void ray_interp_accumulate_fast(Ray &ray, long order, const RingTab *__restrict__ tab,
                                                           const float *__restrict__ m_phi, const float *__restrict__ m_gt,
                                                           const float *__restrict__ m_gp, const float *__restrict__ m_gtt,
                                                           const float *__restrict__ m_gtp, const float *__restrict__ m_gpp,
                                                           long pix[4])
{
  double theta, phi, wgt[4];
  long ring[2];
  vec2ang(ray.n, theta, phi);
  get_interpol_tab(theta, phi, pix, wgt, order, tab, ring[0], ring[1]);
  const long npix_map = 12L << (2 * order);
  // a ray with a non-finite position would index outside the maps; the reference aborts on a missing cell
  // (shtpoissonsolve.c:683-689) -- here the gather is kept in bounds and the NaNs stay visible in the ray
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (!(pix[k] >= 0 && pix[k] < npix_map)) pix[k] = 0;
  // issue the 24 gathers first: their latency hides behind the geometry below
  float f_phi[4], f_gt[4], f_gp[4], f_gtt[4], f_gtp[4], f_gpp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f_phi[k] = __ldg(m_phi + pix[k]); f_gt[k] = __ldg(m_gt + pix[k]); f_gp[k] = __ldg(m_gp + pix[k]);
    f_gtt[k] = __ldg(m_gtt + pix[k]); f_gtp[k] = __ldg(m_gtp + pix[k]); f_gpp[k] = __ldg(m_gpp + pix[k]);
  }
  // ray-side quantities shared by the four transports
  const double inv_r = 1.0 / sqrt(ray.n[0] * ray.n[0] + ray.n[1] * ray.n[1] + ray.n[2] * ray.n[2]);
  const double rv[3] = {ray.n[0] * inv_r, ray.n[1] * inv_r, ray.n[2] * inv_r};
  const double inv_sr = rsqrt((1.0 - rv[2]) * (1.0 + rv[2]));
  double pc[4], ps[4];   // transport angles of the four stencil pixels
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const RingTab rt = tab[ring[h]];
    const RingInfo ri = ring_info(ring[h], order);
    // azimuth of the first pixel: (j + shifted/2) * 2 pi / ringpix; the second one is one pixel spacing further
    // (modulo the ring), i.e. a rotation by the tabulated (cd, sd)
    const long j0 = pix[2 * h] - ri.startpix;
    double s0, c0;
    sincospi((double)(2 * j0 + ri.shifted) * rt.inv_ringpix, &s0, &c0);
    const double c1 = c0 * rt.cd - s0 * rt.sd, s1 = s0 * rt.cd + c0 * rt.sd;
    const double inv_norm = rt.inv_sz * inv_sr;
    const double va[3] = {rt.sz * c0, rt.sz * s0, rt.cz}, vb[3] = {rt.sz * c1, rt.sz * s1, rt.cz};
    paratrans_angle_unit(va, rv, inv_norm, pc[2 * h], ps[2 * h]);
    paratrans_angle_unit(vb, rv, inv_norm, pc[2 * h + 1], ps[2 * h + 1]);
  }
  double pot = 0.0, gtheta = 0.0, gphi = 0.0, t00 = 0.0, t01 = 0.0, t10 = 0.0, t11 = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double w = wgt[k], c = pc[k], s = ps[k];
    pot = fma((double)f_phi[k], w, pot);
    const double tv0 = f_gt[k], tv1 = f_gp[k];
    gtheta = fma(fma(tv0, c, tv1 * s), w, gtheta);
    gphi = fma(fma(tv1, c, -(tv0 * s)), w, gphi);
    const double T00 = f_gtt[k], T01 = f_gtp[k], T11 = f_gpp[k];
    // R^T T R with R = [[c, -s], [s, c]], T symmetric                       [rot_paratrans.c:251-270]
    const double a0 = fma(T00, c, T01 * s), a1 = fma(T01, c, -(T00 * s));   // row 0 of T R
    const double b0 = fma(T01, c, T11 * s), b1 = fma(T11, c, -(T01 * s));   // row 1 of T R
    t00 = fma(fma(c, a0, s * b0), w, t00); t01 = fma(fma(c, a1, s * b1), w, t01);
    t10 = fma(fma(c, b0, -(s * a0)), w, t10); t11 = fma(fma(c, b1, -(s * a1)), w, t11);
  }
  ray.phi = pot;
  ray.alpha[0] += -1.0 * gtheta;
  ray.alpha[1] += -1.0 * gphi;
  ray.U[0] += t00; ray.U[1] += t01; ray.U[2] += t10; ray.U[3] += t11;
}
#endif

#if defined(__CUDACC__)
// Plane constants of the A recursion, formed on the host with the reference's expressions (rayprop.c:134-139)
struct PlaneCoef { double cprev, ccur, cu; };   // (1 - c), c, (wp - wpm1)/wp with c = wpm1 (wp - wpm2) / wp / (wpm1 - wpm2)

// rayprop_sphere for one ray (rayprop.c:18-189, rot_paratrans.c:17-45): wp = w_{p+1}, wpm1 = w_p (the reference's
// argument names).  Same mathematics as the reference with the normalisations shared -- |n x a| = |n| |alpha|
// because a is tangent at n, |theta-hat numerator| = |n| |n_xy| -- and reciprocals multiplied instead of repeated
// divisions.  Agrees with the reference (oracle) to ~1e-15 relative.
__device__ __forceinline__ void ray_propagate_fast(Ray &ray, double wp, double wpm1, const PlaneCoef &pc)
{
  double np[3], betap[3], Ap[4];
  const double nx = ray.n[0], ny = ray.n[1], nz = ray.n[2];
  const double nxy2 = nx * nx + ny * ny;
  const double n2 = nxy2 + nz * nz;
  const double inv_n = rsqrt(n2);
  const double alpha2 = ray.alpha[0] * ray.alpha[0] + ray.alpha[1] * ray.alpha[1];
  if (alpha2 > 0.0) {
    const double alpha = sqrt(alpha2);
    const double inv_nxy = rsqrt(nxy2);
    // a = alpha_theta theta-hat + alpha_phi phi-hat (unit vectors at n)
    const double ct = ray.alpha[0] * inv_nxy * inv_n, cp = ray.alpha[1] * inv_nxy;
    const double a0 = fma(ct, nz * nx, -(cp * ny));
    const double a1 = fma(ct, nz * ny, cp * nx);
    const double a2 = -ct * nxy2;
    // unit rotation axis k = (n x a) / (|n| alpha)
    const double ik = inv_n / alpha;
    const double k0 = fma(ny, a2, -(nz * a1)) * ik, k1 = fma(nz, a0, -(nx * a2)) * ik, k2 = fma(nx, a1, -(ny * a0)) * ik;
    double sn, cs;
    sincos(alpha, &sn, &cs);
    const double omc = 1.0 - cs;
    // Rodrigues: beta' = beta cos + (k x beta) sin + k (k . beta)(1 - cos)     [rot_paratrans.c:17-45]
    const double kb = fma(k0, ray.beta[0], fma(k1, ray.beta[1], k2 * ray.beta[2])) * omc;
    betap[0] = fma(ray.beta[0], cs, fma(fma(k1, ray.beta[2], -(k2 * ray.beta[1])), sn, k0 * kb));
    betap[1] = fma(ray.beta[1], cs, fma(fma(k2, ray.beta[0], -(k0 * ray.beta[2])), sn, k1 * kb));
    betap[2] = fma(ray.beta[2], cs, fma(fma(k0, ray.beta[1], -(k1 * ray.beta[0])), sn, k2 * kb));
    const double qb = 2.0 * fma(nx, betap[0], fma(ny, betap[1], nz * betap[2]));
    const double qc = wpm1 * wpm1 - wp * wp;
    const double q = -0.5 * (qb + copysign(sqrt(fma(qb, qb, -4.0 * qc)), qb));
    double lambda = qc / q;
    if (lambda < 0.0) lambda = q;
    np[0] = fma(betap[0], lambda, nx); np[1] = fma(betap[1], lambda, ny); np[2] = fma(betap[2], lambda, nz);
  } else {
    betap[0] = ray.beta[0]; betap[1] = ray.beta[1]; betap[2] = ray.beta[2];
    const double f = wp / wpm1;
    np[0] = nx * f; np[1] = ny * f; np[2] = nz * f;
  }
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int m = 0; m < 2; ++m)
      Ap[m + 2 * n] = fma(pc.cprev, ray.Aprev[m + 2 * n], fma(pc.ccur, ray.A[m + 2 * n],
                          -(pc.cu * fma(ray.U[0 + 2 * n], ray.A[m + 2 * 0], ray.U[1 + 2 * n] * ray.A[m + 2 * 1]))));
  // transport A, Aprev from n to np; renormalise np to the shell radius
  const double inv_np = rsqrt(np[0] * np[0] + np[1] * np[1] + np[2] * np[2]);
  const double v0[3] = {nx * inv_n, ny * inv_n, nz * inv_n}, v1[3] = {np[0] * inv_np, np[1] * inv_np, np[2] * inv_np};
  const double inv_norm = rsqrt((1.0 - v1[2]) * (1.0 + v1[2]) * (1.0 - v0[2]) * (1.0 + v0[2]));
  double c, s;
  paratrans_angle_unit(v0, v1, inv_norm, c, s);
  {
    const double T00 = ray.A[0], T01 = ray.A[1], T10 = ray.A[2], T11 = ray.A[3];
    const double r00 = fma(T00, c, T01 * s), r01 = fma(T01, c, -(T00 * s)), r10 = fma(T10, c, T11 * s), r11 = fma(T11, c, -(T10 * s));
    ray.Aprev[0] = fma(c, r00, s * r10); ray.Aprev[1] = fma(c, r01, s * r11);
    ray.Aprev[2] = fma(c, r10, -(s * r00)); ray.Aprev[3] = fma(c, r11, -(s * r01));
  }
  {
    const double T00 = Ap[0], T01 = Ap[1], T10 = Ap[2], T11 = Ap[3];
    const double r00 = fma(T00, c, T01 * s), r01 = fma(T01, c, -(T00 * s)), r10 = fma(T10, c, T11 * s), r11 = fma(T11, c, -(T10 * s));
    ray.A[0] = fma(c, r00, s * r10); ray.A[1] = fma(c, r01, s * r11);
    ray.A[2] = fma(c, r10, -(s * r00)); ray.A[3] = fma(c, r11, -(s * r01));
  }
  const double rs = wp * inv_np;
  ray.n[0] = np[0] * rs; ray.n[1] = np[1] * rs; ray.n[2] = np[2] * rs;
  ray.beta[0] = betap[0]; ray.beta[1] = betap[1]; ray.beta[2] = betap[2];
}
#endif

// rayprop_sphere of a -DBORNAPPRX build (rayprop.c:40-62): the ray moves radially to the next shell and the A recursion
// uses U alone (no U A product, no deflection, no transport)
CLB_HD void ray_propagate_born(Ray &ray, double wp, double wpm1, double wpm2)
{
  double Ap[4];
  ray.n[0] = ray.n[0] / wpm1 * wp;
  ray.n[1] = ray.n[1] / wpm1 * wp;
  ray.n[2] = ray.n[2] / wpm1 * wp;
  for (int k = 0; k < 4; ++k)
    Ap[k] = (1.0 - wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2)) * ray.Aprev[k] + (wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2)) * ray.A[k]
            - ((wp - wpm1) / wp) * (ray.U[k]);
  for (int k = 0; k < 4; ++k) { ray.Aprev[k] = ray.A[k]; ray.A[k] = Ap[k]; }
  double r = sqrt(ray.n[0] * ray.n[0] + ray.n[1] * ray.n[1] + ray.n[2] * ray.n[2]);   // rayprop.c:183-187 (both builds)
  r = wp / r;
  ray.n[0] *= r; ray.n[1] *= r; ray.n[2] *= r;
}

}  // namespace clb
