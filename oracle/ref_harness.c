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

/* CLB_SHIM_BUILD: this harness is linked against shim/calclens_b200_shim.c instead of the reference's
 * map2alm_transpose_mpi.c / alm2allmaps_transpose_mpi.c / rayprop.c, and the shim also provides the coarse entry
 * do_healpix_sht_poisson_solve: the reference's own definition is compiled under another name (it still supplies the
 * static shearinterp_comp used by ref_shearinterp). */
#ifdef CLB_SHIM_BUILD
#define do_healpix_sht_poisson_solve ref_cpu_do_healpix_sht_poisson_solve
#include "shtpoissonsolve.c"   /* from -I/root/reference: brings in raytrace.h and static shearinterp_comp */
#undef do_healpix_sht_poisson_solve
void do_healpix_sht_poisson_solve(double densfact, double backdens);
int ref_is_shim(void) { return 1; }
#else
#include "shtpoissonsolve.c"   /* from -I/root/reference: brings in raytrace.h and static shearinterp_comp */
int ref_is_shim(void) { return 0; }
#endif

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

/* ---------------------------------------------------------------------------------------------------------------
 * The reference's own plane loop on the raw-map input path, for one or several ranks of the
 * shared-memory MPI stub: raytrace.c:88-96 (domain decomposition + ray allocation), then per plane raytrace.c:182-269
 * (plane parameters, load balance, ray reset, do_healpix_sht_poisson_solve, rayprop_sphere per owned bundle cell).
 * set_plane_params is static in raytrace.c, so the three SHTONLY lines of it that this path needs are restated here
 * (raytrace.c:452-455 poissonOrder, :486-491 minSL/maxSL/partBuffRad); distances and densfact/backdens come from the
 * caller (calclens_b200.poisson.plane_params mirrors raytrace.c:384-423).
 * ------------------------------------------------------------------------------------------------------------- */
void ref_driver_init(long bundleOrder, long rayOrder, long mapOrder, const char *mapPath, const char *mapName, double partMass,
                     double maxComvDistance, long NumLensPlanes, double OmegaM, const char *ringWeightPath)
{
  MPI_Comm_size(MPI_COMM_WORLD, &NTasks);
  MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
  memset(&rayTraceData, 0, sizeof(rayTraceData));
  rayTraceData.bundleOrder = bundleOrder; rayTraceData.rayOrder = rayOrder; rayTraceData.SHTOrder = mapOrder;
  rayTraceData.HEALPixLensPlaneMapOrder = mapOrder; rayTraceData.UseHEALPixLensPlaneMaps = 1;
  snprintf(rayTraceData.HEALPixLensPlaneMapPath, MAX_FILENAME, "%s", mapPath);
  snprintf(rayTraceData.HEALPixLensPlaneMapName, MAX_FILENAME, "%s", mapName);
  snprintf(rayTraceData.HEALPixRingWeightPath, MAX_FILENAME, "%s", ringWeightPath ? ringWeightPath : "");
  rayTraceData.partMass = partMass; rayTraceData.maxComvDistance = maxComvDistance; rayTraceData.NumLensPlanes = NumLensPlanes;
  rayTraceData.OmegaM = OmegaM;
  rayTraceData.minRa = 0.0; rayTraceData.maxRa = 360.0; rayTraceData.minDec = -90.0; rayTraceData.maxDec = 90.0;
  rayTraceData.maxRayMemImbalance = 0.25; rayTraceData.NumFilesIOInParallel = NTasks;
  rayTraceData.galImageSearchRayBufferRad = sqrt(4.0*M_PI/order2npix(bundleOrder)) + RAYBUFF_RADIUS_ARCMIN/60.0/180.0*M_PI;   /* config.c:226 */
  init_bundlecells();
  alloc_rays();
  init_rays();
}

void ref_driver_plane(long planeNum, double wpm1, double wp, double wpp1, double densfact, double backdens)
{
  long i, j;
  double bundleLength = sqrt(4.0*M_PI/order2npix(rayTraceData.bundleOrder));
  rayTraceData.CurrentPlaneNum = planeNum;
  rayTraceData.planeRadMinus1 = wpm1; rayTraceData.planeRad = wp; rayTraceData.planeRadPlus1 = wpp1;
  rayTraceData.densfact = densfact; rayTraceData.backdens = backdens;
  rayTraceData.poissonOrder = rayTraceData.HEALPixLensPlaneMapOrder;                                   /* raytrace.c:452-455 */
  rayTraceData.minSL = MIN_SMOOTH_TO_RAY_RATIO*sqrt(4.0*M_PI/order2npix(rayTraceData.poissonOrder));    /* raytrace.c:486-487 */
  rayTraceData.maxSL = MIN_SMOOTH_TO_RAY_RATIO*sqrt(4.0*M_PI/order2npix(rayTraceData.poissonOrder));
  rayTraceData.partBuffRad = sqrt(4.0*M_PI/order2npix(rayTraceData.poissonOrder))*10.0 + 2.0*bundleLength + rayTraceData.maxSL*2.0;  /* :490 */
  load_balance_tasks();                                                                                 /* raytrace.c:186 */
  for (i = 0; i < NbundleCells; ++i)                                                                    /* raytrace.c:213-230 */
    if (ISSETBITFLAG(bundleCells[i].active, PRIMARY_BUNDLECELL))
      for (j = 0; j < bundleCells[i].Nrays; ++j) {
        bundleCells[i].rays[j].phi = 0.0;
        bundleCells[i].rays[j].alpha[0] = 0.0; bundleCells[i].rays[j].alpha[1] = 0.0;
        bundleCells[i].rays[j].U[0] = 0.0; bundleCells[i].rays[j].U[1] = 0.0; bundleCells[i].rays[j].U[2] = 0.0; bundleCells[i].rays[j].U[3] = 0.0;
      }
  do_healpix_sht_poisson_solve(rayTraceData.densfact, rayTraceData.backdens);                           /* poissondrivers.c:142 */
  for (i = 0; i < NbundleCells; ++i)                                                                    /* raytrace.c:256-269 */
    if (ISSETBITFLAG(bundleCells[i].active, PRIMARY_BUNDLECELL))
      rayprop_sphere(rayTraceData.planeRadPlus1, rayTraceData.planeRad, rayTraceData.planeRadMinus1, i);
}

long ref_driver_nrays(void) { return NumAllRaysGlobal; }
#ifdef CLB_SHIM_BUILD
void calclens_b200_sync_rays(void);   /* device-resident mode of the shim: the host reads AllRaysGlobal only after this */
#endif
void ref_driver_get_rays(HEALPixRay *out)
{
#ifdef CLB_SHIM_BUILD
  calclens_b200_sync_rays();
#endif
  memcpy(out, AllRaysGlobal, sizeof(HEALPixRay) * NumAllRaysGlobal);
}
void ref_driver_finalize(void)
{
  free(AllRaysGlobal); AllRaysGlobal = NULL; NumAllRaysGlobal = 0;
  destroy_bundlecells();
  healpixsht_destroy_internaldata();
}
int ref_mpi_rank(void) { int r; MPI_Comm_rank(MPI_COMM_WORLD, &r); return r; }
int ref_mpi_size(void) { int n; MPI_Comm_size(MPI_COMM_WORLD, &n); return n; }

/* map2alm_mpi / alm2allmaps_mpi over the ranks of the shared-memory MPI stub with the reference's own plan
 * (healpixsht_plan: ring ranges and m ranges per rank), lmax optionally overridden on every rank.  Every rank passes the
 * full-sky RING map and receives its own alm slice / fills only its own rings of the six output maps. */
static HEALPixSHTPlan make_plan_mpi(long order, long lmax, const double *ring_weights)
{
  HEALPixSHTPlan plan = healpixsht_plan(order);
  int nt, me, t;
  MPI_Comm_size(MPI_COMM_WORLD, &nt); MPI_Comm_rank(MPI_COMM_WORLD, &me);
  if (lmax > 0 && lmax != plan.lmax) {
    assert(lmax <= order2lmax(order));
    /* keep the reference's m split where it still fits, clip to the new band limit */
    for (t = 0; t < nt; ++t) {
      if (plan.firstMTasks[t] > lmax) { plan.firstMTasks[t] = lmax + 1; plan.lastMTasks[t] = lmax; }
      else if (plan.lastMTasks[t] > lmax) plan.lastMTasks[t] = lmax;
    }
    plan.lmax = lmax;
    plan.Nlm = 0;
    for (long m = plan.firstMTasks[me]; m <= plan.lastMTasks[me]; ++m) plan.Nlm += lmax - m + 1;
  }
  if (ring_weights) {
    long n = 2 * order2nside(order);
    plan.ring_weights = (double*)malloc(sizeof(double) * n);
    memcpy(plan.ring_weights, ring_weights, sizeof(double) * n);
  }
  return plan;
}
static void ring_to_mapvec_local(const float *ringmap, float *mapvec, HEALPixSHTPlan plan)
{
  int me; MPI_Comm_rank(MPI_COMM_WORLD, &me);
  long Nside = order2nside(plan.order), nring, ringpix, first = plan.firstRingTasks[me], last = plan.lastRingTasks[me];
  fftwf_complex *mc = (fftwf_complex*)mapvec;
  memset(mapvec, 0, sizeof(fftwf_complex) * plan.Nmapvec);
  for (nring = first; nring <= last; ++nring) {
    ringpix = (nring < Nside) ? 4 * nring : 4 * Nside;
    memcpy((float*)(mc + plan.northStartIndMapvec[nring - first]), ringmap + plan.northStartIndGlobalMap[nring - first], sizeof(float) * ringpix);
    if (nring != 2 * Nside)
      memcpy((float*)(mc + plan.southStartIndMapvec[nring - first]), ringmap + plan.southStartIndGlobalMap[nring - first], sizeof(float) * ringpix);
  }
}
static void mapvec_to_ring_local(const float *mapvec, float *ringmap, HEALPixSHTPlan plan)
{
  int me; MPI_Comm_rank(MPI_COMM_WORLD, &me);
  long Nside = order2nside(plan.order), nring, ringpix, first = plan.firstRingTasks[me], last = plan.lastRingTasks[me];
  const fftwf_complex *mc = (const fftwf_complex*)mapvec;
  for (nring = first; nring <= last; ++nring) {
    ringpix = (nring < Nside) ? 4 * nring : 4 * Nside;
    memcpy(ringmap + plan.northStartIndGlobalMap[nring - first], (const float*)(mc + plan.northStartIndMapvec[nring - first]), sizeof(float) * ringpix);
    if (nring != 2 * Nside)
      memcpy(ringmap + plan.southStartIndGlobalMap[nring - first], (const float*)(mc + plan.southStartIndMapvec[nring - first]), sizeof(float) * ringpix);
  }
}
/* returns this rank's (firstM, lastM, Nlm, firstRing, lastRing) in info[5] */
void ref_mpi_plan_info(long order, long lmax, long *info)
{
  HEALPixSHTPlan plan = make_plan_mpi(order, lmax, NULL);
  int me; MPI_Comm_rank(MPI_COMM_WORLD, &me);
  info[0] = plan.firstMTasks[me]; info[1] = plan.lastMTasks[me]; info[2] = plan.Nlm;
  info[3] = plan.firstRingTasks[me]; info[4] = plan.lastRingTasks[me];
  healpixsht_destroy_plan(plan);
}
void ref_mpi_map2alm(long order, long lmax, const double *ring_weights, const float *ringmap, double *alm_re, double *alm_im)
{
  HEALPixSHTPlan plan = make_plan_mpi(order, lmax, ring_weights);
  float *mapvec = (float*)malloc(sizeof(fftwf_complex) * plan.Nmapvec);
  ring_to_mapvec_local(ringmap, mapvec, plan);
  map2alm_mpi(alm_re, alm_im, mapvec, plan);
  free(mapvec);
  healpixsht_destroy_plan(plan);
}
void ref_mpi_alm2allmaps(long order, long lmax, double *alm_re, double *alm_im, float *maps)
{
  HEALPixSHTPlan plan = make_plan_mpi(order, lmax, NULL);
  long Npix = order2npix(order), k;
  float *mv[6];
  for (k = 0; k < 6; ++k) { mv[k] = (float*)malloc(sizeof(fftwf_complex) * plan.Nmapvec); memset(mv[k], 0, sizeof(fftwf_complex) * plan.Nmapvec); }
  alm2allmaps_mpi(alm_re, alm_im, mv[0], mv[1], mv[2], mv[3], mv[4], mv[5], plan);
  for (k = 0; k < 6; ++k) { mapvec_to_ring_local(mv[k], maps + k * Npix, plan); free(mv[k]); }
  healpixsht_destroy_plan(plan);
}

/* map2alm_mpi / alm2allmaps_mpi restricted to the m range [m0, m1] of a single-rank plan: the reference functions
 * honour plan.firstMTasks/lastMTasks (map2alm_transpose_mpi.c:329-334,418-425), so this is the unmodified code doing
 * the ring FFT of every ring and the Legendre stage of the selected m only -- the sampled oracle used at Nside 4096.
 * alm arrays hold sum_{m=m0}^{m1} (lmax-m+1) entries, m-major. */
void ref_map2alm_mrange(long order, long lmax, long m0, long m1, const double *ring_weights, const float *ringmap, double *alm_re, double *alm_im)
{
  HEALPixSHTPlan plan = make_plan(order, lmax, ring_weights);
  long m;
  plan.firstMTasks[0] = m0; plan.lastMTasks[0] = m1; plan.Nlm = 0;
  for (m = m0; m <= m1; ++m) plan.Nlm += plan.lmax - m + 1;
  float *mapvec = (float*)malloc(sizeof(fftwf_complex) * plan.Nmapvec);
  ring_to_mapvec(ringmap, mapvec, plan);
  map2alm_mpi(alm_re, alm_im, mapvec, plan);
  free(mapvec);
  healpixsht_destroy_plan(plan);
}
void ref_alm2allmaps_mrange(long order, long lmax, long m0, long m1, double *alm_re, double *alm_im, float *maps)
{
  HEALPixSHTPlan plan = make_plan(order, lmax, NULL);
  long Npix = order2npix(order), k, m;
  plan.firstMTasks[0] = m0; plan.lastMTasks[0] = m1; plan.Nlm = 0;
  for (m = m0; m <= m1; ++m) plan.Nlm += plan.lmax - m + 1;
  float *mv[6];
  for (k = 0; k < 6; ++k) { mv[k] = (float*)malloc(sizeof(fftwf_complex) * plan.Nmapvec); memset(mv[k], 0, sizeof(fftwf_complex) * plan.Nmapvec); }
  alm2allmaps_mpi(alm_re, alm_im, mv[0], mv[1], mv[2], mv[3], mv[4], mv[5], plan);
  for (k = 0; k < 6; ++k) { mapvec_to_ring(mv[k], maps + k * Npix, plan); free(mv[k]); }
  healpixsht_destroy_plan(plan);
}

/* ---------------------------------------------------------------------------------------------------------------
 * SAMPLED oracle for sizes where the full reference transform takes CPU-hours (Nside 4096: 7e3 core-seconds).
 * Built from the reference's own primitives -- ring_analysis / ring_synthesis (healpix_shtrans.c:549,168, over the FFT
 * shim), plmgen (healpix_plmgen.c:73), get_ring_info2, get_lmin_ylm -- with only the glue loops restated (same
 * statements as oracle/port/calclens_port.c, which tests/test_oracle.py holds bit-identical to the full reference
 * functions; tests/test_oracle.py::test_sampled_oracle_equals_full_reference pins THESE functions the same way):
 *   pack + phase      map2alm_transpose_mpi.c:237-274        accumulate   :462-498, equator :500-534
 *   synthesis sums    alm2allmaps_transpose_mpi.c:320-447    unpack/fold  :826-883
 *   1/sin, cot terms  :1045-1051, :1097-1147
 * ------------------------------------------------------------------------------------------------------------- */
/* alm rows of the selected m (ascending), concatenated: sum over mlist of (lmax - m + 1) entries */
void ref_sample_map2alm(long order, long lmax, const double *ring_weights, const float *ringmap, long nm, const long *mlist,
                        double *alm_re, double *alm_im)
{
  long Nside = order2nside(order), Npix = order2npix(order), nrp = 2 * Nside, nslot = 4 * Nside;
  long ring, hemi, k, m, l, i, startpix, ringpix, shifted, firstl, lmin, mind;
  double cth, sth, quadweight = 4.0*M_PI/Npix, w, phase[2], tmp[2], v[2], sfact, fac1;
  double *gre = (double*)calloc((size_t)nm * nslot, sizeof(double)), *gim = (double*)calloc((size_t)nm * nslot, sizeof(double));
  fftwf_complex *buf = (fftwf_complex*)fftwf_malloc(sizeof(fftwf_complex) * (2 * Nside + 1));
  float *fb = (float*)buf;
  for (ring = 1; ring <= nrp; ++ring) {
    get_ring_info2(ring, &startpix, &ringpix, &cth, &sth, &shifted, order);
    w = ring_weights ? ring_weights[ring - 1] : 0.0;
    w += 1.0;
    w *= quadweight;
    for (hemi = 0; hemi < 2; ++hemi) {
      if (hemi && ring == nrp) break;
      const float *src = ringmap + (hemi ? Npix - startpix - ringpix : startpix);
      for (i = 0; i < ringpix; ++i) fb[i] = (float)(src[i] * w);
      ring_analysis(ringpix, fb);
      for (k = 0; k < nm; ++k) {
        m = mlist[k];
        mind = m % ringpix;
        if (mind > ringpix / 2) { mind = ringpix - mind; v[0] = buf[mind][0]; v[1] = -buf[mind][1]; }
        else { v[0] = buf[mind][0]; v[1] = buf[mind][1]; }
        if (shifted) {
          phase[0] = cos(m*M_PI/ringpix);
          phase[1] = -sin(m*M_PI/ringpix);
          tmp[0] = v[0]*phase[0] - v[1]*phase[1];
          tmp[1] = v[0]*phase[1] + v[1]*phase[0];
          v[0] = tmp[0]; v[1] = tmp[1];
        }
        gre[k * nslot + 2 * (ring - 1) + hemi] = v[0]; gim[k * nslot + 2 * (ring - 1) + hemi] = v[1];
      }
    }
  }
  fftwf_free(buf);
  plmgen_data *pd = plmgen_init(lmax, 1e-30);
  double *plm = (double*)malloc(sizeof(double) * (lmax + 1));
  long lmind = 0;
  for (k = 0; k < nm; ++k) {
    m = mlist[k];
    for (l = m; l <= lmax; ++l) { alm_re[lmind + l - m] = 0.0; alm_im[lmind + l - m] = 0.0; }
    for (ring = 1; ring <= nrp; ++ring) {
      get_ring_info2(ring, &startpix, &ringpix, &cth, &sth, &shifted, order);
      lmin = get_lmin_ylm(m, sth);
      if (lmin > lmax) continue;
      plmgen(cth, sth, m, plm, &firstl, pd);
      if (firstl > lmax) continue;
      double nr = gre[k * nslot + 2 * (ring - 1)], ni = gim[k * nslot + 2 * (ring - 1)];
      double sr = gre[k * nslot + 2 * (ring - 1) + 1], si = gim[k * nslot + 2 * (ring - 1) + 1];
      if (ring < nrp) {
        sfact = 1.0 - 2.0*((firstl+m)%2);
        for (l = firstl; l <= lmax; ++l) {
          alm_re[lmind + l-m] += nr*plm[l];
          alm_im[lmind + l-m] += ni*plm[l];
          fac1 = sfact*plm[l];
          alm_re[lmind + l-m] += sr*fac1;
          alm_im[lmind + l-m] += si*fac1;
          sfact = -sfact;
        }
      } else {
        for (l = firstl; l <= lmax; ++l) { alm_re[lmind + l-m] += nr*plm[l]; alm_im[lmind + l-m] += ni*plm[l]; }
      }
    }
    lmind += lmax - m + 1;
  }
  free(plm); plmgen_destroy(pd); free(gre); free(gim);
}

/* the six synthesis sums of one (m, ring) for both hemispheres: an[k], as[k] (re, im); alm row = l - m index.
 * Returns 0 when the pair is skipped by the lmin cut or by firstl (sums are then zero). */
static int sample_syn_sums(plmgen_data *pd, double *plm, long lmax, long m, double cth, double sth, const double *ar_, const double *ai_,
                           double an[6][2], double as[6][2])
{
  long l, firstl, k;
  for (k = 0; k < 6; ++k) { an[k][0] = an[k][1] = as[k][0] = as[k][1] = 0.0; }
  if (get_lmin_ylm(m, (float)sth) > lmax) return 0;                     /* alm2allmaps_transpose_mpi.c:308 (float cast) */
  plmgen(cth, sth, m, plm, &firstl, pd);
  if (firstl > lmax) return 0;
  double sfact = 1.0 - (((firstl + m) % 2) << 1);
  for (l = firstl; l <= lmax; ++l) {
    double rval = ar_[l - m] * plm[l], ival = ai_[l - m] * plm[l];
    an[0][0] += rval; an[0][1] += ival; as[0][0] += sfact * rval; as[0][1] += sfact * ival;
    rval *= m; ival *= m;
    an[2][0] -= ival; an[2][1] += rval; as[2][0] -= sfact * ival; as[2][1] += sfact * rval;
    rval *= m; ival *= m;
    an[5][0] -= rval; an[5][1] -= ival; as[5][0] -= sfact * rval; as[5][1] -= sfact * ival;
    sfact = -sfact;
  }
  sfact = 1.0 - 2.0 * ((firstl + m) % 2);
  double cs = cth / sth, cs2 = cth / sth / sth, c2s2 = cs * cs, gl1n, gl1s;
  l = firstl;
  if (l > 0) {
    double ar = ar_[l - m], ai = ai_[l - m];
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
    double ar = ar_[l - m], ai = ai_[l - m];
    double sv = sqrt((2.0 * l + 1.0) / (2.0 * l - 1.0) * ((double)(l * l - m * m)));
    double gn = ((double)l) * cs * plm[l] - sv * plm[l - 1] / sth, gs = -sfact * gn;
    an[1][0] += ar * gn; an[1][1] += ai * gn; as[1][0] += ar * gs; as[1][1] += ai * gs;
    double fac1 = ((double)l) * cs, fac2 = ((double)l) * plm[l] * (1.0 + c2s2), fac3 = sv / sth;
    double qn = fac1 * gn - fac2 - fac3 * gl1n, qs = -fac1 * gs - sfact * fac2 - fac3 * gl1s;
    double gf = sv * plm[l - 1] * cs2;
    qn += gf; qs += gf * sfact;
    an[3][0] += ar * qn; an[3][1] += ai * qn; as[3][0] += ar * qs; as[3][1] += ai * qs;
    gl1n = gn; gl1s = gs;
    gn *= m; gs *= m;
    an[4][0] -= ai * gn; an[4][1] += ar * gn; as[4][0] -= ai * gs; as[4][1] += ar * gs;
    sfact = -sfact;
  }
  return 1;
}

/* the six float maps on the selected ring pairs (rplist = north ring numbers 1..2Nside, both hemispheres), from the full
 * alm (m-major, all m): out[k][j] for k = 0..5, pixels of the north ring then of its mirror, ring pairs concatenated */
void ref_sample_alm2allmaps_rings(long order, long lmax, const double *alm_re, const double *alm_im, long nr, const long *rplist, float *out)
{
  long Nside = order2nside(order), nrp = 2 * Nside, NM = lmax + 1;
  long startpix, ringpix, shifted, m, k, i, ir, hemi;
  double cth, sth;
  plmgen_data *pd = plmgen_init(lmax, 1e-30);
  double *plm = (double*)malloc(sizeof(double) * (lmax + 1));
  double *qr = (double*)malloc(sizeof(double) * 2 * 6 * NM), *qi = (double*)malloc(sizeof(double) * 2 * 6 * NM);
  fftwf_complex *y = (fftwf_complex*)fftwf_malloc(sizeof(fftwf_complex) * (2 * Nside + 1));
  float *x = (float*)y;
  long total = 0, opos = 0;
  for (ir = 0; ir < nr; ++ir) { get_ring_info2(rplist[ir], &startpix, &ringpix, &cth, &sth, &shifted, order); total += ringpix * (rplist[ir] == nrp ? 1 : 2); }
  for (ir = 0; ir < nr; ++ir) {
    long ring = rplist[ir];
    get_ring_info2(ring, &startpix, &ringpix, &cth, &sth, &shifted, order);
    long n = ringpix, nc = n / 2 + 1, lmind = 0;
    for (m = 0; m <= lmax; ++m) {
      double an[6][2], as[6][2];
      sample_syn_sums(pd, plm, lmax, m, cth, sth, alm_re + lmind, alm_im + lmind, an, as);
      for (k = 0; k < 6; ++k) {
        qr[(0 * 6 + k) * NM + m] = an[k][0]; qi[(0 * 6 + k) * NM + m] = an[k][1];
        qr[(1 * 6 + k) * NM + m] = as[k][0]; qi[(1 * 6 + k) * NM + m] = as[k][1];
      }
      lmind += lmax - m + 1;
    }
    for (hemi = 0; hemi < 2; ++hemi) {
      if (hemi && ring == nrp) break;
      float *o[6];
      for (k = 0; k < 6; ++k) o[k] = out + (size_t)k * total + opos;
      for (k = 0; k < 6; ++k) {
        for (i = 0; i < nc; ++i) { y[i][0] = 0.0f; y[i][1] = 0.0f; }
        const double *br = qr + (hemi * 6 + k) * NM, *bi = qi + (hemi * 6 + k) * NM;
        for (m = 0; m <= lmax; ++m) {                                     /* alm2allmaps_transpose_mpi.c:826-883 */
          long mp = m % n;
          if (mp < nc) {
            long l = (m - mp) / n; double sk = (shifted && (l % 2)) ? -1.0 : 1.0;
            y[mp][0] += br[m] * sk; y[mp][1] += bi[m] * sk;
          }
          if (m > 0) {
            mp = n - 1 - ((m - 1) % n);
            if (mp < nc) {
              long l = (-m - mp) / n; double sk = (shifted && (l % 2)) ? -1.0 : 1.0;
              y[mp][0] += br[m] * sk; y[mp][1] -= bi[m] * sk;
            }
          }
        }
        ring_synthesis(n, shifted, x);                                    /* healpix_shtrans.c:168-205 (phase + c2r, in place) */
        if (k == 2 || k == 5 || k == 4) for (i = 0; i < n; ++i) x[i] /= sth;   /* :1045-1051 */
        if (k == 5) for (i = 0; i < n; ++i) x[i] /= sth;
        memcpy(o[k], x, sizeof(float) * n);
      }
      for (i = 0; i < n; ++i) {                                           /* :1097-1147 */
        if (!hemi) { o[4][i] = (float)(o[4][i] - cth / sth * o[2][i]); o[5][i] = (float)(o[5][i] + cth / sth * o[1][i]); }
        else { o[4][i] = (float)(o[4][i] + cth / sth * o[2][i]); o[5][i] = (float)(o[5][i] - cth / sth * o[1][i]); }
      }
      opos += n;
    }
  }
  fftwf_free(y); free(qr); free(qi); free(plm); plmgen_destroy(pd);
}
