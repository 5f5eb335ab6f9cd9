/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never part of the product library).
 *
 * Thin ctypes-callable entry points around the UNMODIFIED CALCLENS reference functions, which the Makefile
 * in this directory compiles straight from /root/reference into oracle/_ref/libcalclens_ref.so together with
 * the single-rank stubs in oracle/stubs/.  Nothing here re-implements reference arithmetic except the two
 * trivially small driver loops that are not callable functions in the reference:
 *   - the Poisson filter alm *= -1/(l(l+1))              (shtpoissonsolve.c:526-550)
 *   - the per-ray accumulation loop around shearinterp_comp (shtpoissonsolve.c:666-702)
 * shearinterp_comp is `static` in shtpoissonsolve.c, so that file is #included textually below
 * (the build reads it from /root/reference; it is not copied into this repository).
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <assert.h>

#include "shtpoissonsolve.c"   /* from -I/root/reference: brings in raytrace.h and static shearinterp_comp */

/* build a single-rank plan, optionally overriding lmax (the reference hard-wires 3*Nside-1) */
static HEALPixSHTPlan make_plan(long order, long lmax, const double *ring_weights)
{
  HEALPixSHTPlan plan = healpixsht_plan(order);
  if (lmax > 0) {
    assert(lmax <= order2lmax(order));
    plan.lmax = lmax;
    plan.lastMTasks[0] = lmax;
    plan.Nlm = num_lms(lmax);
  }
  if (ring_weights) {
    long n = 2 * order2nside(order);
    plan.ring_weights = (double*)malloc(sizeof(double) * n);
    memcpy(plan.ring_weights, ring_weights, sizeof(double) * n);
  }
  return plan;
}

long ref_nmapvec(long order) { HEALPixSHTPlan p = healpixsht_plan(order); long n = p.Nmapvec; healpixsht_destroy_plan(p); return n; }
long ref_nlm(long lmax) { return num_lms(lmax); }

/* RING-ordered full-sky map <-> the plan's padded ring-pair layout (healpix_shtrans.c:90-118) */
static void ring_to_mapvec(const float *ringmap, float *mapvec, HEALPixSHTPlan plan)
{
  long Nside = order2nside(plan.order), nring, ringpix;
  fftwf_complex *mc = (fftwf_complex*)mapvec;
  memset(mapvec, 0, sizeof(fftwf_complex) * plan.Nmapvec);
  for (nring = 1; nring <= 2 * Nside; ++nring) {
    ringpix = (nring < Nside) ? 4 * nring : 4 * Nside;
    memcpy((float*)(mc + plan.northStartIndMapvec[nring - 1]), ringmap + plan.northStartIndGlobalMap[nring - 1], sizeof(float) * ringpix);
    if (nring != 2 * Nside)
      memcpy((float*)(mc + plan.southStartIndMapvec[nring - 1]), ringmap + plan.southStartIndGlobalMap[nring - 1], sizeof(float) * ringpix);
  }
}
static void mapvec_to_ring(const float *mapvec, float *ringmap, HEALPixSHTPlan plan)
{
  long Nside = order2nside(plan.order), nring, ringpix;
  const fftwf_complex *mc = (const fftwf_complex*)mapvec;
  for (nring = 1; nring <= 2 * Nside; ++nring) {
    ringpix = (nring < Nside) ? 4 * nring : 4 * Nside;
    memcpy(ringmap + plan.northStartIndGlobalMap[nring - 1], (const float*)(mc + plan.northStartIndMapvec[nring - 1]), sizeof(float) * ringpix);
    if (nring != 2 * Nside)
      memcpy(ringmap + plan.southStartIndGlobalMap[nring - 1], (const float*)(mc + plan.southStartIndMapvec[nring - 1]), sizeof(float) * ringpix);
  }
}

/* map2alm_mpi (map2alm_transpose_mpi.c:54) on a RING-ordered float map; alm out m-major, Nlm = num_lms(lmax) */
void ref_map2alm(long order, long lmax, const double *ring_weights, const float *ringmap, double *alm_re, double *alm_im)
{
  HEALPixSHTPlan plan = make_plan(order, lmax, ring_weights);
  float *mapvec = (float*)malloc(sizeof(fftwf_complex) * plan.Nmapvec);
  ring_to_mapvec(ringmap, mapvec, plan);
  map2alm_mpi(alm_re, alm_im, mapvec, plan);
  free(mapvec);
  healpixsht_destroy_plan(plan);
}

/* the Poisson filter exactly as the caller applies it between the two transforms (shtpoissonsolve.c:526-550) */
void ref_poisson_filter(long lmax, double *alm_re, double *alm_im)
{
  long i = 0, l, m;
  for (m = 0; m <= lmax; ++m)
    for (l = m; l <= lmax; ++l) {
      if (l == 0 && m == 0) { alm_re[i] = 0.0; alm_im[i] = 0.0; }
      else {
        alm_re[i] *= (double)(-1.0 / ((double)l) / (((double)l) + 1.0));
        alm_im[i] *= (double)(-1.0 / ((double)l) / (((double)l) + 1.0));
      }
      ++i;
    }
}

/* alm2allmaps_mpi (alm2allmaps_transpose_mpi.c:53): six RING-ordered float maps out, maps[k*Npix + pix],
 * k = 0 phi, 1 grad_theta, 2 grad_phi, 3 grad_theta_theta, 4 grad_theta_phi, 5 grad_phi_phi (argument order of
 * the reference prototype, healpix_shtrans.h:70-72) */
void ref_alm2allmaps(long order, long lmax, double *alm_re, double *alm_im, float *maps)
{
  HEALPixSHTPlan plan = make_plan(order, lmax, NULL);
  long Npix = order2npix(order), k;
  float *mv[6];
  for (k = 0; k < 6; ++k) { mv[k] = (float*)malloc(sizeof(fftwf_complex) * plan.Nmapvec); memset(mv[k], 0, sizeof(fftwf_complex) * plan.Nmapvec); }
  alm2allmaps_mpi(alm_re, alm_im, mv[0], mv[1], mv[2], mv[3], mv[4], mv[5], plan);
  for (k = 0; k < 6; ++k) { mapvec_to_ring(mv[k], maps + k * Npix, plan); free(mv[k]); }
  healpixsht_destroy_plan(plan);
}

/* rayprop_sphere (rayprop.c:18) over a flat ray array, presented as one fake bundle cell */
void ref_rayprop(HEALPixRay *rays, long Nrays, double wp, double wpm1, double wpm2)
{
  HEALPixBundleCell cell;
  memset(&cell, 0, sizeof(cell));
  cell.Nrays = Nrays; cell.rays = rays;
  HEALPixBundleCell *save = bundleCells; long saveN = NbundleCells;
  bundleCells = &cell; NbundleCells = 1;
  rayprop_sphere(wp, wpm1, wpm2, 0);
  bundleCells = save; NbundleCells = saveN;
}

/* the same with the reference built -DBORNAPPRX (rayprop.c:40-62): the Makefile compiles rayprop.c a second time under
 * the symbol rayprop_sphere_born */
void rayprop_sphere_born(double wp, double wpm1, double wpm2, long bundleCellInd);
void ref_rayprop_born(HEALPixRay *rays, long Nrays, double wp, double wpm1, double wpm2)
{
  HEALPixBundleCell cell;
  memset(&cell, 0, sizeof(cell));
  cell.Nrays = Nrays; cell.rays = rays;
  HEALPixBundleCell *save = bundleCells; long saveN = NbundleCells;
  bundleCells = &cell; NbundleCells = 1;
  rayprop_sphere_born(wp, wpm1, wpm2, 0);
  bundleCells = save; NbundleCells = saveN;
}

/* shearinterp_comp + the caller's accumulation (shtpoissonsolve.c:666-702,1122-1204) on a full-sky domain:
 * every bundle cell is PRIMARY and owns its 4^(poissonOrder-bundleOrder) NEST-ordered map cells.
 * maps = six RING-ordered float maps in the ref_alm2allmaps order.  Returns the number of rays for which
 * shearinterp_comp reported a missing cell (must be 0). */
long ref_shearinterp(long poissonOrder, long bundleOrder, const float *maps, HEALPixRay *rays, long Nrays)
{
  long Npix = order2npix(poissonOrder), Nb = order2npix(bundleOrder), i, k, bad = 0;
  long shift = 2 * (poissonOrder - bundleOrder);
  HEALPixMapCell *cells[6];
  for (k = 0; k < 6; ++k) {
    cells[k] = (HEALPixMapCell*)malloc(sizeof(HEALPixMapCell) * Npix);
    for (i = 0; i < Npix; ++i) { cells[k][i].index = i; cells[k][i].val = maps[k * Npix + nest2ring(i, poissonOrder)]; }
  }
  HEALPixBundleCell *bc = (HEALPixBundleCell*)calloc(Nb, sizeof(HEALPixBundleCell));
  for (i = 0; i < Nb; ++i) { bc[i].nest = i; bc[i].active = 0; SETBITFLAG(bc[i].active, PRIMARY_BUNDLECELL); bc[i].firstMapCell = i << shift; }
  bundleCells = bc; NbundleCells = Nb;
  rayTraceData.poissonOrder = poissonOrder; rayTraceData.bundleOrder = bundleOrder;
  mapCells = cells[0]; mapCellsGradTheta = cells[1]; mapCellsGradPhi = cells[2];
  mapCellsGradThetaTheta = cells[3]; mapCellsGradThetaPhi = cells[4]; mapCellsGradPhiPhi = cells[5];
  NmapCells = Npix;
  for (i = 0; i < Nrays; ++i) {
    double rvec[3], alpha[2] = {0.0, 0.0}, U[4] = {0.0, 0.0, 0.0, 0.0}, lenspot = 0.0;
    rvec[0] = rays[i].n[0]; rvec[1] = rays[i].n[1]; rvec[2] = rays[i].n[2];
    if (shearinterp_comp(rvec, &lenspot, alpha, U)) { ++bad; continue; }
    rays[i].phi = lenspot;
    rays[i].alpha[0] += -1.0 * alpha[0];
    rays[i].alpha[1] += -1.0 * alpha[1];
    rays[i].U[0] += U[0]; rays[i].U[1] += U[1]; rays[i].U[2] += U[2]; rays[i].U[3] += U[3];
  }
  for (k = 0; k < 6; ++k) free(cells[k]);
  free(bc);
  bundleCells = NULL; NbundleCells = 0; NmapCells = 0;
  mapCells = mapCellsGradTheta = mapCellsGradPhi = mapCellsGradThetaTheta = mapCellsGradThetaPhi = mapCellsGradPhiPhi = NULL;
  return bad;
}

/* ring weights through the reference's own reader (healpix_shtrans.c:361) and the FITS stub */
long ref_read_ring_weights(const char *path, long order, double *out)
{
  HEALPixSHTPlan plan = healpixsht_plan(order);
  long n = 2 * order2nside(order);
  read_ring_weights((char*)path, &plan);
  if (!plan.ring_weights) { healpixsht_destroy_plan(plan); return -1; }
  memcpy(out, plan.ring_weights, sizeof(double) * n);
  healpixsht_destroy_plan(plan);
  return n;
}

/* plmgen (healpix_plmgen.c:73) for one (m, ring): vec[0..lmax], returns firstl */
long ref_plmgen(long lmax, double cth, double sth, long m, double *vec)
{
  plmgen_data *d = plmgen_init(lmax, 1e-30);
  long firstl = lmax + 1;
  plmgen(cth, sth, m, vec, &firstl, d);
  plmgen_destroy(d);
  return firstl;
}

/* init_rays (raytrace_utils.c:302-347) for NEST pixels first..first+n-1, using the reference's nest2vec */
void ref_init_rays(HEALPixRay *rays, long first, long n, long ray_order, double binL_2)
{
  long i;
  memset(rays, 0, sizeof(HEALPixRay) * n);
  for (i = 0; i < n; ++i) {
    rays[i].nest = first + i;
    nest2vec(rays[i].nest, rays[i].beta, ray_order);
    rays[i].n[0] = rays[i].beta[0] * binL_2; rays[i].n[1] = rays[i].beta[1] * binL_2; rays[i].n[2] = rays[i].beta[2] * binL_2;
    rays[i].A[0] = 1.0; rays[i].A[3] = 1.0; rays[i].Aprev[0] = 1.0; rays[i].Aprev[3] = 1.0;
  }
}

/* write_rays' pre-output transform (rayio.c:300-312): parallel transport A, Aprev from the ray's current position to
 * the pixel it is observed in, then switch every ray quantity from the (theta, phi) to the (ra, dec) basis */
void ref_ray_output(HEALPixRay *rays, long Nrays, long ray_order)
{
  long i;
  rayTraceData.rayOrder = ray_order;
  for (i = 0; i < Nrays; ++i) {
    paratrans_ray_curr2obs(&rays[i]);
    rot_ray_ang2radec(&rays[i]);
  }
}

/* NGP particle deposit, the loop of shtpoissonsolve.c:128-150 (NGPSHTDENS) on a full-sky single-rank domain, written
 * to a RING-ordered float map: mapCells[...].val += (float)(mass/MASS_SCALE) at ang2nest(vec2ang(pos), poissonOrder).
 * pos = 3 floats per particle (Part.pos, raytrace.h:246-253). */
void ref_deposit_ngp(const float *pos, const float *mass, long Nparts, long order, float *ringmap)
{
  long k;
  for (k = 0; k < Nparts; ++k) {
    double vec[3], theta, phi;
    vec[0] = (double)pos[3 * k]; vec[1] = (double)pos[3 * k + 1]; vec[2] = (double)pos[3 * k + 2];
    vec2ang(vec, &theta, &phi);
    long mapNest = ang2nest(theta, phi, order);
    ringmap[nest2ring(mapNest, order)] += (float)(mass[k] / MASS_SCALE);
  }
}

long ref_sizeof_ray(void) { return (long)sizeof(HEALPixRay); }
