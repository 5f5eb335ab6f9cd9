/* shim/calclens_b200_shim.c -- link-time drop-in for the CALCLENS SHTONLY hot path.
 *
 * Compile this file against the CALCLENS headers and link it INSTEAD OF map2alm_transpose_mpi.o,
 * alm2allmaps_transpose_mpi.o and rayprop.o (and, with -DCLB_SHIM_COARSE, instead of shtpoissonsolve.o), together with
 * libcalclens_b200.so.  It defines, with the reference's exact prototypes,
 *     map2alm_mpi                      healpix_shtrans.h:67      (caller shtpoissonsolve.c:522)
 *     alm2allmaps_mpi                  healpix_shtrans.h:70-72   (caller shtpoissonsolve.c:566)
 *     rayprop_sphere                   raytrace.h:431            (callers raytrace.c:262, propagate_to_cmb_from_restart.c:380)
 *     do_healpix_sht_poisson_solve     raytrace.h:358            (callers poissondrivers.c:61,142)   [CLB_SHIM_COARSE]
 * and forwards them to the C ABI of include/calclens_b200.h.  One MPI rank drives one GPU (rank % device count, or
 * CALCLENS_B200_DEVICE); with NTasks > 1 the two transposes (map2alm_transpose_mpi.c:339-381,
 * alm2allmaps_transpose_mpi.c:656-724) run over NVLink peer memory, bootstrapped through MPI_Allgather.
 * The plan is honoured as the reference passes it: lmax (healpix_shtrans.h:39), firstRingTasks/lastRingTasks,
 * firstMTasks/lastMTasks, per-rank mapvec slices (healpix_shtrans.c:54-160), ring weights.
 * Failure mode as the reference's: message on stderr + MPI_Abort(MPI_COMM_WORLD, 123).
 *
 * Coarse entry (CLB_SHIM_COARSE): the whole of do_healpix_sht_poisson_solve for the raw-map input path
 * (UseHEALPixLensPlaneMaps, shtpoissonsolve.c:342-436) and for NGP-deposited particles on one rank (:111-156), i.e.
 * scaling, both transforms, the Poisson filter and the per-ray interpolation (:666-702) on the GPU; the ring -> domain
 * map shuffles (map_shuffle.c) disappear because every rank's GPU receives the maps over NVLink.
 *   default                      rays travel host -> device -> host inside every call (no host code change at all)
 *   CALCLENS_B200_RESIDENT=1     rays stay on the device between planes: the first rayprop_sphere call of a plane
 *                                propagates ALL rays of the rank, the others are no-ops, and the host must call
 *                                calclens_b200_sync_rays() before it reads AllRaysGlobal (write_rays, write_restart).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <mpi.h>
#include "raytrace.h"          /* CALCLENS: HEALPixRay, HEALPixBundleCell, rayTraceData, healpix_shtrans.h */
#include "calclens_b200.h"

#ifndef MASS_SCALE
#define MASS_SCALE 1e10        /* shtpoissonsolve.c:36 */
#endif

static struct {
  clb_solver *s;
  unsigned long sig;
  int device_set;
  /* coarse entry */
  float *h_counts;             /* pinned full-sky count map (only this rank's rings are filled) */
  long h_counts_npix;
  int resident, rays_on_device, plane_propagated;
  double prop_key[3];
} G;

static void shim_die(const char *msg)
{
  fprintf(stderr, "calclens_b200 shim: %s\n", msg);
  fflush(stderr);
  MPI_Abort(MPI_COMM_WORLD, 123);
  abort();
}

static void shim_allgather(const void *send, void *recv, long bytes, void *ctx)
{
  (void)ctx;
  MPI_Allgather((void *)send, (int)bytes, MPI_BYTE, recv, (int)bytes, MPI_BYTE, MPI_COMM_WORLD);
}

static unsigned long mix(unsigned long h, unsigned long v) { h ^= v + 0x9e3779b97f4a7c15ul + (h << 6) + (h >> 2); return h; }

static unsigned long plan_signature(HEALPixSHTPlan plan, int ntasks)
{
  unsigned long h = 1469598103934665603ul;
  long i, n = 2 * order2nside(plan.order);
  h = mix(h, (unsigned long)plan.order); h = mix(h, (unsigned long)plan.lmax); h = mix(h, (unsigned long)ntasks);
  for (i = 0; i < ntasks; ++i) {
    h = mix(h, (unsigned long)plan.firstRingTasks[i]); h = mix(h, (unsigned long)plan.lastRingTasks[i]);
    h = mix(h, (unsigned long)plan.firstMTasks[i]); h = mix(h, (unsigned long)plan.lastMTasks[i]);
  }
  if (plan.ring_weights) for (i = 0; i < n; ++i) { unsigned long b; memcpy(&b, &plan.ring_weights[i], 8); h = mix(h, b); }
  return h;
}

/* the solver for this decomposition: created on first use, rebuilt when order / lmax / decomposition / weights change */
static clb_solver *shim_solver(HEALPixSHTPlan plan)
{
  int ntasks, me;
  long i, t, nrp = 2 * order2nside(plan.order);
  MPI_Comm_size(MPI_COMM_WORLD, &ntasks);
  MPI_Comm_rank(MPI_COMM_WORLD, &me);
  unsigned long sig = plan_signature(plan, ntasks);
  if (G.s && G.sig == sig) return G.s;
  if (G.s) { clb_solver_destroy(G.s); G.s = NULL; G.rays_on_device = 0; }
  if (!G.device_set) {
    const char *env = getenv("CALCLENS_B200_DEVICE");
    clb_set_device(env ? atoi(env) : me % clb_device_count());
    G.device_set = 1;
    G.resident = getenv("CALCLENS_B200_RESIDENT") && atoi(getenv("CALCLENS_B200_RESIDENT")) != 0;
  }
  int *rp_owner = (int *)malloc(sizeof(int) * nrp), *m_owner = (int *)malloc(sizeof(int) * (plan.lmax + 1));
  for (i = 0; i < nrp; ++i) rp_owner[i] = -1;
  for (i = 0; i <= plan.lmax; ++i) m_owner[i] = -1;
  for (t = 0; t < ntasks; ++t) {
    for (i = plan.firstRingTasks[t]; i <= plan.lastRingTasks[t]; ++i) if (i >= 1 && i <= nrp) rp_owner[i - 1] = (int)t;
    for (i = plan.firstMTasks[t]; i <= plan.lastMTasks[t]; ++i) if (i >= 0 && i <= plan.lmax) m_owner[i] = (int)t;
  }
  for (i = 0; i < nrp; ++i) if (rp_owner[i] < 0) shim_die("ring without an owner in the HEALPixSHTPlan");
  for (i = 0; i <= plan.lmax; ++i) if (m_owner[i] < 0) shim_die("m without an owner in the HEALPixSHTPlan");
  /* halo_deg = 0: every rank receives the full derivative maps (rays may be anywhere in the rank's Peano domain) */
  G.s = clb_solver_create(plan.order, plan.lmax, rayTraceData.rayOrder > 0 ? rayTraceData.rayOrder : plan.order, plan.ring_weights,
                          ntasks, me, rp_owner, m_owner, ntasks > 1 ? shim_allgather : NULL, NULL, 0.0);
  free(rp_owner); free(m_owner);
  if (!G.s) shim_die("no peer access between the GPUs of this job (the fused exchange needs one NVLink/NVSwitch node)");
  G.sig = sig;
  return G.s;
}

/* ---- healpix_shtrans.h:67 ---- */
void map2alm_mpi(double *alm_real, double *alm_imag, float *mapvec, HEALPixSHTPlan plan)
{
  clb_solver *s = shim_solver(plan);
  /* (the reference weights and transforms mapvec in place, i.e. destroys it; here it is left untouched) */
  clb_solver_map2alm_mapvec(s, mapvec, plan.northStartIndMapvec, plan.southStartIndMapvec, alm_real, alm_imag, NULL);
}

/* ---- healpix_shtrans.h:70-72 ---- */
void alm2allmaps_mpi(double *alm_real, double *alm_imag, float *mapvec, float *mapvec_gt, float *mapvec_gp,
                     float *mapvec_gtt, float *mapvec_gtp, float *mapvec_gpp, HEALPixSHTPlan plan)
{
  clb_solver *s = shim_solver(plan);
  float *mv[6] = {mapvec, mapvec_gt, mapvec_gp, mapvec_gtt, mapvec_gtp, mapvec_gpp};
  clb_solver_alm2allmaps_mapvec(s, alm_real, alm_imag, mv, plan.northStartIndMapvec, plan.southStartIndMapvec, NULL);
}

#ifdef BORNAPPRX
#define SHIM_PROP_MODE (4 | 8)     /* rayprop.c:40-62 */
#else
#define SHIM_PROP_MODE 4
#endif

/* all rays of this rank in one call (the loop of raytrace.c:256-269 collapsed); rays stay where they are (host) */
void calclens_b200_rayprop_all(double wp, double wpm1, double wpm2)
{
  if (!G.device_set) { int me; MPI_Comm_rank(MPI_COMM_WORLD, &me); clb_set_device(me % clb_device_count()); G.device_set = 1; }
  clb_ray_step(AllRaysGlobal, NumAllRaysGlobal, NULL, 0, wp, wpm1, wpm2, SHIM_PROP_MODE);
}

/* device-resident mode: the host changed AllRaysGlobal (restart, load balance): upload again at the next plane */
void calclens_b200_invalidate_rays(void) { G.rays_on_device = 0; }

/* device-resident mode: bring the rays back before the host reads them (write_rays, write_restart, gridsearch) */
void calclens_b200_sync_rays(void)
{
  if (G.s && G.rays_on_device) clb_solver_get_rays(G.s, AllRaysGlobal, NULL);
}

/* ---- raytrace.h:431 ---- */
void rayprop_sphere(double wp, double wpm1, double wpm2, long bundleCellInd)
{
  if (G.resident && G.s && G.rays_on_device) {
    /* the host calls once per owned bundle cell with the same radii: the first call of a plane moves every ray */
    if (!(G.plane_propagated && G.prop_key[0] == wp && G.prop_key[1] == wpm1 && G.prop_key[2] == wpm2)) {
      clb_solver_ray_update(G.s, wp, wpm1, wpm2, SHIM_PROP_MODE, 0, NULL);
      G.plane_propagated = 1; G.prop_key[0] = wp; G.prop_key[1] = wpm1; G.prop_key[2] = wpm2;
    }
    return;
  }
  if (!G.device_set) { int me; MPI_Comm_rank(MPI_COMM_WORLD, &me); clb_set_device(me % clb_device_count()); G.device_set = 1; }
  if (bundleCells[bundleCellInd].Nrays > 0)
    clb_ray_step(bundleCells[bundleCellInd].rays, bundleCells[bundleCellInd].Nrays, NULL, 0, wp, wpm1, wpm2, SHIM_PROP_MODE);
}

#ifdef CLB_SHIM_COARSE
/* ---- raytrace.h:358 ---- */
void do_healpix_sht_poisson_solve(double densfact, double backdens)
{
  long i, k, nring, ringpix;
  const long order = rayTraceData.poissonOrder, Nside = order2nside(order), Npix = order2npix(order);
  const double area = 4.0*M_PI/Npix;
  int ntasks, me;
  MPI_Comm_size(MPI_COMM_WORLD, &ntasks);
  MPI_Comm_rank(MPI_COMM_WORLD, &me);
  logProfileTag(PROFILETAG_SHT);
  HEALPixSHTPlan plan = healpixsht_plan(order);                        /* shtpoissonsolve.c:316 */
  if (strlen(rayTraceData.HEALPixRingWeightPath) > 0)
    read_ring_weights(rayTraceData.HEALPixRingWeightPath, &plan);      /* :317-323 */
  clb_solver *s = shim_solver(plan);
  if (!G.h_counts || G.h_counts_npix != Npix) {
    if (G.h_counts) { clb_host_unregister(G.h_counts); free(G.h_counts); }
    G.h_counts = (float *)malloc(sizeof(float) * Npix);
    if (!G.h_counts) shim_die("out of memory for the count map");
    clb_host_register(G.h_counts, (long)(sizeof(float) * Npix));       /* pinned in place: the GPU reads this rank's rings directly */
    G.h_counts_npix = Npix;
  }
  float premul;
  if (rayTraceData.UseHEALPixLensPlaneMaps) {
    /* raw RING-ordered float32 count map "<path>/<name>.<plane>" (shtpoissonsolve.c:342-436): this rank's rings only */
    char fname[MAX_FILENAME];
    long firstRing = plan.firstRingTasks[me], lastRing = plan.lastRingTasks[me];
    sprintf(fname, "%s/%s.%ld", rayTraceData.HEALPixLensPlaneMapPath, rayTraceData.HEALPixLensPlaneMapName, rayTraceData.CurrentPlaneNum);
    FILE *fp = fopen(fname, "r");
    if (!fp) shim_die("cannot open the HEALPix lens-plane map");
    for (nring = firstRing; nring <= lastRing; ++nring) {
      ringpix = (nring < Nside) ? 4 * nring : 4 * Nside;
      long start = plan.northStartIndGlobalMap[nring - firstRing];
      fseek(fp, start * sizeof(float), SEEK_SET);
      if (fread(G.h_counts + start, sizeof(float), (size_t)ringpix, fp) != (size_t)ringpix) shim_die("short read of the lens-plane map");
      if (nring != 2 * Nside) {
        start = plan.southStartIndGlobalMap[nring - firstRing];
        fseek(fp, start * sizeof(float), SEEK_SET);
        if (fread(G.h_counts + start, sizeof(float), (size_t)ringpix, fp) != (size_t)ringpix) shim_die("short read of the lens-plane map");
      }
    }
    fclose(fp);
    premul = (float)(rayTraceData.partMass/MASS_SCALE);                 /* :426 */
  } else {
    /* particles of the owned bundle cells, NGP-deposited (:111-156 with NGPSHTDENS); one rank only: the reference adds the
       deposits of all ranks in its peano2ring shuffle (map_shuffle.c:841) */
    if (ntasks != 1) shim_die("particle input with NTasks > 1 is not supported by the coarse entry (use the raw-map input path)");
    memset(G.h_counts, 0, sizeof(float) * Npix);
    for (i = 0; i < NbundleCells; ++i)
      if (ISSETBITFLAG(bundleCells[i].active, PRIMARY_BUNDLECELL) && bundleCells[i].Nparts > 0)
        for (k = 0; k < bundleCells[i].Nparts; ++k) {
          double vec[3], theta, phi;
          vec[0] = (double)lensPlaneParts[k + bundleCells[i].firstPart].pos[0];
          vec[1] = (double)lensPlaneParts[k + bundleCells[i].firstPart].pos[1];
          vec[2] = (double)lensPlaneParts[k + bundleCells[i].firstPart].pos[2];
          vec2ang(vec, &theta, &phi);
          G.h_counts[nest2ring(ang2nest(theta, phi, order), order)] += (float)(lensPlaneParts[k + bundleCells[i].firstPart].mass/MASS_SCALE);
        }
    premul = 1.0f;
  }
  const float densmul = (float)(densfact/area*MASS_SCALE);             /* :468 */
  const float fback = (float)backdens;                                 /* :478 */
  double t0 = MPI_Wtime();
  clb_solver_load_density(s, G.h_counts, premul, densmul, fback, NULL);
  clb_solver_solve(s, NULL, NULL);                                      /* map2alm, -1/(l(l+1)), alm2allmaps (:517-570) */
  /* interpolation at the rays (:666-702): rays of the owned bundle cells are contiguous in AllRaysGlobal (raytrace_utils.c:280-298) */
  if (!(G.resident && G.rays_on_device)) {
    clb_solver_set_rays(s, AllRaysGlobal, NumAllRaysGlobal, NULL);
    G.rays_on_device = G.resident;
  }
  clb_solver_ray_update(s, 0.0, 0.0, 0.0, G.resident ? (1 | 2) : 2, 0, NULL);   /* resident rays are reset here (raytrace.c:213-230) */
  G.plane_propagated = 0;
  if (!G.resident) clb_solver_get_rays(s, AllRaysGlobal, NULL);
  if (clb_solver_check(s, NULL)) shim_die("a ray left the part of the sky this rank received");
  if (me == 0) { fprintf(stderr, "GPU SHT Poisson solve + ray interpolation took %g seconds.\n", MPI_Wtime() - t0); fflush(stderr); }
  healpixsht_destroy_plan(plan);
  logProfileTag(PROFILETAG_SHT);
}
#endif
