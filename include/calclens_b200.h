/* include/calclens_b200.h -- C ABI of libcalclens_b200.so (sm_100a CUDA implementation of the CALCLENS SHTONLY
 * lens-plane hot path).  Plain C types only; every entry point cites the reference interface it replaces
 * (paths relative to the CALCLENS source tree).  There is no CPU fallback: every call needs a CUDA device and
 * aborts with a message on any CUDA error, which is the reference's own failure mode on this path
 * (MPI_Abort(MPI_COMM_WORLD,123), map2alm_transpose_mpi.c:129-138, shtpoissonsolve.c:683-689).
 *
 * Conventions
 *   ring pair rp = 0 .. 2*Nside-1 : north ring rp+1 and its mirror 4*Nside-1-rp (the last pair is the equator).
 *   maps      : float32, RING order, full-sky buffer of 12*Nside^2 pixels (a rank touches only its own rings).
 *   alm       : double, m-major: for each owned m (ascending), l = m..lmax contiguous
 *               (map2alm_transpose_mpi.c:418-425, healpix_shtrans.c:523-526 lm2index).
 *   six fields: 0 phi, 1 d_theta, 2 d_phi/sin, 3 d_theta d_theta, 4 d_theta d_phi, 5 d_phi d_phi -- the argument
 *               order of alm2allmaps_mpi (healpix_shtrans.h:70-72).
 *   rays      : the reference's 176-byte HEALPixRay records (raytrace.h:284-293), unchanged.
 *   "_dev" entry points take DEVICE pointers and a cudaStream_t (as void*); they return the number of kernels
 *   they launched.  The other entry points take HOST pointers and do their own transfers.
 */
#ifndef CALCLENS_B200_H
#define CALCLENS_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct clb_sht_plan clb_sht_plan;

int clb_abi_version(void);   /* 2 */
/* number of CUDA devices; aborts if the CUDA runtime reports none (no CPU fallback exists) */
int clb_device_count(void);
void clb_set_device(int device);
/* kernels launched by this library since load (bench.py reports it as gpu_launches) */
long clb_launch_count(void);
/* tuning knobs (development; the defaults are the measured optimum on B200):
 *   0 Legendre synthesis rings per thread, one-shell pass (1..4; 4)       1 analysis rings per thread, one-shell (1,2,4,6,8; 8)
 *   2 threads per CTA of the large ring FFTs (256..1024; 512; read at plan creation)
 *   3 warps per CTA of the Legendre kernels (1,2,4; 4)                    4 run every ring FFT from global scratch (0|1; 0)
 *   5 partial-sum rows per m of the Legendre analysis (0 = automatic)     6 field groups per ring in the ring synthesis (0|1|3)
 *   7 ring-FFT class launches on parallel streams (0|1; 1)                8 skip phases of the ring synthesis (timing aid, wrong results)
 *   9 synthesis rings per thread, two-shell pass (1..4; 3)               10 analysis rings per thread, two-shell pass (1,2,4,6,8; 8)
 *  11 shells per SHT pass a solver is provisioned for (1|2; 2; read by clb_solver_create)
 *  12 overlap the analysis warp sum with the next block: 0 never, 1 one-shell passes only (default), 2 always */
void clb_set_tuning(int what, int value);

/* ---- plan: replaces healpixsht_plan / healpixsht_destroy_plan (healpix_shtrans.c:54-160, :496-516) and
 * read_ring_weights' result (healpix_shtrans.c:361-423: ring_weights = the 2*Nside doubles of
 * weight_ring_nNNNNN.fits, or NULL).  lmax travels in the reference plan (healpix_shtrans.h:39) and is honoured
 * as given.  rp_owner[2*Nside] / m_owner[lmax+1] give the owning rank of every ring pair / every m (NULL = rank 0
 * owns all); they generalise the reference's contiguous firstRingTasks/lastRingTasks and firstMTasks/lastMTasks. */
clb_sht_plan *clb_sht_plan_create(long order, long lmax, const double *ring_weights, int nranks, int rank,
                                  const int *rp_owner, const int *m_owner);
void clb_sht_plan_destroy(clb_sht_plan *plan);
/* what: 0 Npix, 1 lmax, 2 number of local alm (Nlm of healpix_shtrans.h:41), 3 local ring pairs, 4 local m,
 * 5/6 g send/recv totals, 7/8 b send/recv totals (units of complex doubles), 9 nranks, 10 rank */
long clb_sht_plan_query(const clb_sht_plan *plan, int what);
/* per-peer element counts of the two transposes (complex doubles): which = 0 g_send, 1 g_recv, 2 b_send, 3 b_recv.
 * They replace the sendcnts/recvcnts of map2alm_transpose_mpi.c:329-347 and alm2allmaps_transpose_mpi.c:656-672. */
void clb_sht_plan_counts(const clb_sht_plan *plan, int which, long *counts);
void clb_sht_plan_local_m(const clb_sht_plan *plan, int *m_list);
void clb_sht_plan_local_ring_pairs(const clb_sht_plan *plan, int *rp_list);

/* ---- stages of map2alm_mpi (map2alm_transpose_mpi.c:54-641) ---- */
/* ring weights + r2c + alias/phase + pack: :151-315.  g_send: clb_sht_plan_query(5) complex doubles */
int clb_ring_analysis_dev(const clb_sht_plan *plan, const float *map, double *g_send, void *stream);
/* Legendre analysis :427-536, optionally fused with the Poisson filter of shtpoissonsolve.c:526-550 */
int clb_legendre_analysis_dev(clb_sht_plan *plan, const double *g_recv, double *alm_re, double *alm_im,
                              int apply_poisson_filter, void *stream);
/* ---- stages of alm2allmaps_mpi (alm2allmaps_transpose_mpi.c:53-1240) ---- */
/* Legendre synthesis of the six fields :272-595.  b_send: clb_sht_plan_query(7) complex doubles */
int clb_legendre_synthesis_dev(clb_sht_plan *plan, const double *alm_re, const double *alm_im, double *b_send, void *stream);
/* ---- two shells (lens planes) per Legendre pass -- SURVEY.md section 8f-4; the reference has no counterpart (it solves
 * plane by plane, shtpoissonsolve.c:517-570, although plane p+1's solve does not depend on the rays).  The lambda_lm
 * recurrence is generated once and applied to both shells: 3 instead of 4 FP64 instructions per (m, ring pair, l, shell) in
 * the analysis, 7 instead of 8 in the synthesis.  nshell = 1 or 2; shell s reads g at g_recv + s * query(6) complex doubles
 * and delivers alm at alm_re/alm_im + s * query(2); the synthesis reads alm the same way and writes b at
 * b_send + s * query(7).  Every shell's result is bit-identical to what the one-shell call gives for it. ---- */
int clb_legendre_analysis_shells_dev(clb_sht_plan *plan, const double *g_recv, double *alm_re, double *alm_im,
                                     int apply_poisson_filter, int nshell, void *stream);
int clb_legendre_synthesis_shells_dev(clb_sht_plan *plan, const double *alm_re, const double *alm_im, double *b_send,
                                      int nshell, void *stream);
/* unpack/alias fold + phase + c2r + 1/sin scalings + cot terms :818-1147 */
int clb_ring_synthesis_dev(const clb_sht_plan *plan, const double *b_recv, float *const maps[6], void *stream);

/* ---- fused exchange over peer memory (one process per GPU, one NVLink/NVSwitch node).  It replaces the MPI
 * hypercube transposes (map2alm_transpose_mpi.c:339-381, alm2allmaps_transpose_mpi.c:656-724) and the ring -> domain
 * map shuffle (map_shuffle.c:22-631) by stores from the producing kernels straight into the consumer rank's buffers:
 *   clb_peer_alloc/export/import : device buffers other ranks may map (CUDA IPC; the 64-byte handle travels through
 *                                  whatever the host uses for bootstrap -- MPI_Allgather in CALCLENS, the
 *                                  torch.distributed store here); import returns NULL when peer access is unavailable
 *   clb_sht_plan_set_peers       : g_send_ptrs[q] = rank q's g SEND buffer (clb_sht_plan_query(5) complex doubles on
 *                                  rank q), b_recv_ptrs[q] = rank q's b RECEIVE buffer (clb_sht_plan_query(8)), as
 *                                  mapped into this process.  Afterwards clb_legendre_analysis_dev ignores g_recv and
 *                                  reads g straight out of the ring owners' send buffers (1 KB coalesced NVLink reads),
 *                                  and clb_legendre_synthesis_dev ignores b_send and stores b into the ring owners'
 *                                  receive buffers.  The host orders producer and consumer stages with a stream
 *                                  barrier across ranks.
 *   clb_maps_broadcast_dev       : store this rank's rings of the six maps into the peers' maps
 *                                  (peer_maps[q*6+k] = map k of rank q, up to 8 ranks);
 *                                  need (device, may be NULL = every pixel to every rank) and coarse_order come from
 *                                  clb_domain_masks: pixels go only to the ranks whose ray domain + halo covers them
 *   clb_domain_masks (host)      : mask[12*4^coarse_order], bit q set when rank q needs the cell -- the counterpart of
 *                                  the reference's halo ("buffer") bundle cells, raytrace_utils.c:116-161 ---- */
void *clb_peer_alloc(long bytes);
void clb_peer_free(void *p);
void clb_peer_export(void *p, void *handle64);
void *clb_peer_import(const void *handle64);
void clb_peer_release(void *p);
void clb_sht_plan_set_peers(clb_sht_plan *plan, void *const *g_send_ptrs, void *const *b_recv_ptrs);
/* the same with peer buffers that hold nshell (1 or 2) shells back to back: rank q's second shell starts at
 * + its query(5) (g) / query(8) (b) complex doubles */
void clb_sht_plan_set_peers_shells(clb_sht_plan *plan, void *const *g_send_ptrs, void *const *b_recv_ptrs, int nshell);
int clb_maps_broadcast_dev(const clb_sht_plan *plan, float *const local_maps[6], float *const *peer_maps,
                           const unsigned char *need, long coarse_order, void *stream);
void clb_domain_masks(long ray_order, int nranks, long coarse_order, double margin_rad, unsigned char *mask);
/* clb_ray_step_dev with two optional extras: (need, coarse_order, rank, err) verifies, per ray, that its interpolation
 * stencil lies inside the cells this rank received (*err |= 1 otherwise; the reference aborts on a missing map cell,
 * shtpoissonsolve.c:683-689); sum6 (device, 6 doubles) receives the sums of clb_ray_summary_dev without a second pass
 * over the rays.  Any of need / err / sum6 may be NULL. */
int clb_ray_step_ex_dev(void *rays, long nrays, const float *const maps[6], long map_order, double wp, double wpm1,
                        double wpm2, int mode, const unsigned char *need, long coarse_order, int rank, int *err,
                        double *sum6, void *stream);

/* ---- density scaling of shtpoissonsolve.c:426,454-502 (full-sky: no vacuum cells):
 * map = (map * premul) * densmul - backdens, all in float like the reference ---- */
int clb_scale_density_dev(float *map, long npix, float premul, float densmul, float backdens, void *stream);
/* the same scaling while loading this rank's rings of a full-sky count map: src may be device memory or PINNED
 * host memory (the raw map of shtpoissonsolve.c:342-436), dst is a device map; only the owned rings are read, so on
 * N ranks each one moves 1/N of the map across PCIe */
int clb_load_density_dev(const clb_sht_plan *plan, const float *src, float *dst, float premul, float densmul,
                         float backdens, void *stream);

/* ---- ray step.  mode bits: 1 zero phi/alpha/U (raytrace.c:213-230); 2 interpolate + accumulate
 * (shtpoissonsolve.c:666-702 with shearinterp_comp :1122-1204); 4 propagate = rayprop_sphere(wp, wpm1, wpm2, .)
 * (rayprop.c:18-189; argument names as in raytrace.h:431); 8 (with 4) the -DBORNAPPRX form of it (rayprop.c:40-62).
 * maps are needed only with bit 2. ---- */
int clb_ray_step_dev(void *rays, long nrays, const float *const maps[6], long map_order, double wp, double wpm1,
                     double wpm2, int mode, void *stream);

/* ray initialisation of init_rays (raytrace_utils.c:302-347): ray i observes from NEST pixel first_nest+i at
 * ray_order, n = beta*binL/2, A = Aprev = identity */
int clb_ray_init_dev(void *rays, long nrays, long first_nest, long ray_order, double binL_2, void *stream);
/* six sums over the rays (convergence, shear 1/2, |alpha|^2, phi, rotation): the per-plane scalar a host reads back */
int clb_ray_summary_dev(const void *rays, long nrays, double *out6, void *stream);

/* ---- the callers either side of the path ("next" rows) ----
 * write_rays' pre-output transform (rayio.c:300-312: paratrans_ray_curr2obs + rot_ray_ang2radec,
 * rot_paratrans.c:274-302,375-411) from the device-resident rays into out_rays (device, same layout), which the host
 * copies back and hands to the unchanged FITS/binary writer; the resident rays are left untouched (the reference
 * undoes the transform after writing, rayio.c:340-352). */
int clb_ray_output_dev(const void *rays, void *out_rays, long nrays, long ray_order, void *stream);
/* NGP particle deposit of shtpoissonsolve.c:128-150: ringmap[pix(pos)] += (float)(mass/1e10), pos = 3 floats per particle
 * (Part.pos, raytrace.h:246-253), ringmap RING-ordered and zeroed by the caller; feeds clb_load_density_dev */
int clb_deposit_ngp_dev(const float *pos, const float *mass, long nparts, long order, float *ringmap, void *stream);

/* ---- host-pointer entry points (transfers inside; device buffers are pooled between calls, clb_pool_release frees
 * them) ---- */
void clb_pool_release(void);
/* map2alm_mpi on a RING-ordered map (healpix_shtrans.h:67) */
void clb_map2alm(clb_sht_plan *plan, const float *ringmap, double *alm_re, double *alm_im, int apply_poisson_filter);
/* alm2allmaps_mpi; maps = 6 consecutive RING-ordered maps (healpix_shtrans.h:70-72) */
void clb_alm2allmaps(clb_sht_plan *plan, const double *alm_re, const double *alm_im, float *maps);
/* The same two transforms on the reference's padded ring-pair buffers ("mapvec", healpix_shtrans.c:90-118): ring
 * pair i starts at north_start[i] / south_start[i] (units of 8 bytes, -1 for the equator's missing mirror).  These
 * are what a link-time replacement of map2alm_mpi / alm2allmaps_mpi forwards to. */
void clb_map2alm_mapvec(clb_sht_plan *plan, float *mapvec, const long *north_start, const long *south_start,
                        double *alm_re, double *alm_im);
void clb_alm2allmaps_mapvec(clb_sht_plan *plan, const double *alm_re, const double *alm_im, float *const mapvec[6],
                            const long *north_start, const long *south_start);
/* rayprop_sphere over a host array of HEALPixRay (rayprop.c:18); mode as clb_ray_step_dev; maps (host, 6
 * consecutive RING maps) may be NULL unless bit 2 is set */
void clb_ray_step(void *rays, long nrays, const float *maps, long map_order, double wp, double wpm1, double wpm2, int mode);
/* One lens plane end to end, the work of do_healpix_sht_poisson_solve (shtpoissonsolve.c:38-708) from the scaled
 * map on plus the plane's rayprop_sphere calls (raytrace.c:256-269): counts map (host, RING) -> scale -> map2alm ->
 * filter -> alm2allmaps -> zero/interpolate/propagate the rays (host array, updated in place). */
void clb_lens_plane(clb_sht_plan *plan, const float *ringmap, float premul, float densmul, float backdens, void *rays,
                    long nrays, double wp, double wpm1, double wpm2);

/* ---- HEALPix indexing on the device, exposed for the bit-exactness tests (device pointers) ----
 * what: 0 ring2nest, 1 nest2ring, 2 ang2nest(theta,phi), 3 nest2peano      (healpix_utils.c:420,413,548,427) */
void clb_healpix_index_dev(int what, long order, long n, const long *in, const double *theta, const double *phi,
                           long *out, void *stream);
/* vec2ang + get_interpol (healpix_utils.c:120,971): vec[3n] -> pix[4n], wgt[4n] */
void clb_healpix_interpol_dev(long order, long n, const double *vec, long *pix, double *wgt, void *stream);
/* the same stencil exactly as the ray kernel forms it (tabulated ring colatitudes, reciprocal weights): the index path
 * that runs in clb_ray_step*_dev, exposed for the bit-exactness tests */
void clb_ray_stencil_dev(long order, long n, const double *vec, long *pix, double *wgt, void *stream);

/* ---- persistent lens-plane solver: one object per rank / GPU that owns everything living across planes (plan,
 * exchange buffers, six derivative maps, double-buffered density, device-resident rays) and runs the whole per-plane
 * path: do_healpix_sht_poisson_solve from the raw count map on (shtpoissonsolve.c:342-708) plus the plane's
 * rayprop_sphere calls (raytrace.c:256-269).  It replaces the body of the reference's plane loop between the ray reset
 * (raytrace.c:213-230) and the end of the propagation loop (:256-269).
 *   nranks > 1: one process per GPU on one NVLink/NVSwitch node.  The exchange buffers are mapped into every process
 *   (CUDA IPC); the handles travel through `allgather` (every rank contributes `bytes` bytes, receives nranks*bytes in
 *   rank order -- MPI_Allgather(send, bytes, MPI_BYTE, recv, bytes, MPI_BYTE, MPI_COMM_WORLD) in CALCLENS).  The two
 *   transposes (map2alm_transpose_mpi.c:339-381, alm2allmaps_transpose_mpi.c:656-724) and the ring -> domain shuffle
 *   (map_shuffle.c:22-631) are then stores/loads of the producing/consuming kernels over NVLink, ordered by a
 *   device-side barrier; no host synchronisation on the per-plane path.  Returns NULL if peer mapping is unavailable.
 *   rp_owner / m_owner as in clb_sht_plan_create (NULL: round-robin, balanced); rays are split into contiguous NEST
 *   ranges (compact sky domains, cf. loadbalance.c:151-181); halo_deg > 0 limits the map broadcast to each rank's
 *   domain + halo (the reference's MAPBUFF bundle cells, raytrace_utils.c:116-161), 0 sends full maps.
 *   Entry points that take a stream enqueue on it and return without synchronising unless stated. ---- */
typedef struct clb_solver clb_solver;
typedef void (*clb_allgather_fn)(const void *send, void *recv, long bytes, void *ctx);
clb_solver *clb_solver_create(long sht_order, long lmax, long ray_order, const double *ring_weights, int nranks, int rank,
                              const int *rp_owner, const int *m_owner, clb_allgather_fn allgather, void *ctx,
                              double halo_deg);
void clb_solver_destroy(clb_solver *s);   /* collective on multi-rank solvers */
/* what: 0 rays on this rank, 1 first NEST index, 2 fused exchange active, 3 kernels launched, 4 halo mask in use,
 * 5 mean fraction of the sky a rank receives x 1e6, 6 host barriers in use (ranks time-share a GPU), 7 Npix,
 * 8 shells per pass the solver is provisioned for */
long clb_solver_query(const clb_solver *s, int what);
/* device pointers owned by the solver: 0 six maps [6][Npix], 1 rays, 2/3 alm re/im, 4 the clb_sht_plan, 5 halo mask,
 * 6/7 the two density buffers, 8 the six ray sums, 9 the second map set (the partner plane of a two-shell pass) */
void *clb_solver_ptr(clb_solver *s, int what);
/* alloc_rays + init_rays (raytrace_utils.c:265-347) for this rank's NEST range; returns the number of rays */
long clb_solver_init_rays(clb_solver *s, double binL_2, void *stream);
/* rays from / to a host array of HEALPixRay (restart, write_rays); get synchronises */
void clb_solver_set_rays(clb_solver *s, const void *host_rays, long nrays, void *stream);
void clb_solver_get_rays(clb_solver *s, void *host_rays, void *stream);
/* write_rays' pre-output transform (rayio.c:300-312) into a host array; the resident rays are untouched; synchronises */
void clb_solver_ray_output(clb_solver *s, void *host_out, void *stream);
/* One lens plane.  counts_map: full-sky RING float32 count map in device memory, pinned/registered host memory (read by
 * the GPU directly, only this rank's rings cross PCIe) or pageable host memory (staged); scalings as
 * clb_scale_density_dev; (wpp1, wp, wpm1) = rayprop_sphere's arguments at raytrace.c:262.  sum6 != NULL: the six ray
 * sums of this rank are copied back and the call synchronises; returns 0, or bit 0 = a ray left this rank's domain +
 * halo (the reference aborts: shtpoissonsolve.c:683-689), bit 1 = a peer never reached a barrier. */
int clb_solver_step(clb_solver *s, const float *counts_map, float premul, float densmul, float backdens, double wpp1,
                    double wp, double wpm1, double *sum6, void *stream);
/* register the NEXT plane's map: the following clb_solver_step starts loading it on a side stream behind its own
 * kernels, and the step after that, given the same pointer and scalings, finds its density already on the device */
void clb_solver_set_next(clb_solver *s, const float *next_counts_map, float premul, float densmul, float backdens);
/* (up to two planes may be registered between steps; they are loaded in that order into the two density buffers)
 * Two planes per SHT pass: register the plane AFTER the one the next clb_solver_step will be given.  That step then runs
 * both planes through one pass of each Legendre kernel (clb_legendre_*_shells_dev), updates the rays with its own plane
 * and keeps the partner's six maps; the following clb_solver_step, given the partner's pointer and scalings, finds them
 * and only updates the rays.  Rays, maps and sums are bit-identical to two ordinary steps.  A plane is identified by its
 * map pointer and the three scalings; the partner's map is read during the step that follows clb_solver_set_pair (and a
 * map registered with clb_solver_set_next while that step's kernels run), so its contents must be final by then. */
void clb_solver_set_pair(clb_solver *s, const float *partner_counts_map, float premul, float densmul, float backdens);
/* synchronise and return the error bits of clb_solver_step */
int clb_solver_check(clb_solver *s, void *stream);
/* the pieces of a step, for callers that interleave their own work: density load into the current buffer, the SHT
 * Poisson solve (density_dev NULL = the current buffer), alm2allmaps_mpi from device alm of this rank's m, ray update
 * (mode bits as clb_ray_step_dev: 1 reset, 2 interpolate + accumulate, 4 propagate, 8 Born form) */
void clb_solver_load_density(clb_solver *s, const float *counts_map, float premul, float densmul, float backdens, void *stream);
void clb_solver_solve(clb_solver *s, const float *density_dev, void *stream);
void clb_solver_alm2allmaps(clb_solver *s, const double *alm_re, const double *alm_im, void *stream);
void clb_solver_ray_update(clb_solver *s, double wpp1, double wp, double wpm1, int mode, int with_summary, void *stream);
/* map2alm_mpi / alm2allmaps_mpi with the reference's per-rank arguments (healpix_shtrans.h:67,70-72): mapvec = this
 * rank's ring pairs in the padded layout of healpix_shtrans.c:90-118 (north_start/south_start = the plan's
 * northStartIndMapvec/southStartIndMapvec, indexed by ring - firstRingTasks[rank]), alm = the m range this rank owns.
 * The solver must have been created with rp_owner/m_owner describing firstRingTasks..lastRingTasks and
 * firstMTasks..lastMTasks.  Host pointers; collective over the ranks; synchronises.  (No Poisson filter: the caller
 * applies it, shtpoissonsolve.c:526-550.) */
void clb_solver_map2alm_mapvec(clb_solver *s, const float *mapvec, const long *north_start, const long *south_start,
                               double *alm_re, double *alm_im, void *stream);
void clb_solver_alm2allmaps_mapvec(clb_solver *s, const double *alm_re, const double *alm_im, float *const mapvec[6],
                                   const long *north_start, const long *south_start, void *stream);
/* per-stage CUDA-event timing of the last step: scale, fft_analysis, exchange g, legendre_analysis, legendre_synthesis,
 * exchange b, fft_synthesis, map broadcast, rays (milliseconds; synchronises) */
void clb_solver_set_timing(clb_solver *s, int on);
void clb_solver_stage_ms(clb_solver *s, double *ms9);
/* pin / unpin a host array in place so transfers from it run at full PCIe rate (AllRaysGlobal, raw maps) */
int clb_host_register(void *p, long bytes);
void clb_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif
