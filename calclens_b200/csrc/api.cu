// calclens_b200/csrc/api.cu -- extern "C" boundary of libcalclens_b200.so (see include/calclens_b200.h).
#include "launch.cuh"
#include "../../include/calclens_b200.h"
#include <string.h>

namespace clb {
extern int g_syn_rings_per_thread, g_ana_rings_per_thread, g_fft_threads_big, g_leg_warps_per_cta, g_fft_force_scratch, g_ana_rows, g_syn2_rings_per_thread, g_ana2_rings_per_thread, g_fft_field_groups, g_fft_debug, g_solver_shells, g_ana_pipeline, g_fft_streams;

// safe[c] = AND of mask[d] over every cell d whose centre lies within neighbour_rad of c's centre (c included): a ray whose
// stencil starts in a "safe" cell cannot touch an undelivered pixel, so the ray kernel skips the per-pixel mask check there
void safe_masks(long coarse_order, double neighbour_rad, const unsigned char *mask, unsigned char *safe)
{
  const long nc = 12L << (2 * coarse_order);
  std::vector<double> cen(3 * nc);
  for (long c = 0; c < nc; ++c) nest2vec(c, coarse_order, &cen[3 * c]);
  const double cosr = cos(neighbour_rad);
  for (long c = 0; c < nc; ++c) {
    unsigned m = mask[c];
    const double *a = &cen[3 * c];
    for (long d = 0; d < nc && m; ++d) {
      if ((mask[d] & m) == m) continue;
      const double *b = &cen[3 * d];
      if (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] >= cosr) m &= mask[d];
    }
    safe[c] = (unsigned char)m;
  }
}

static long g_launches = 0;
void count_launches(int n) { g_launches += n; }

__global__ void scale_density_kernel(float *__restrict__ map, long npix, float premul, float densmul, float backdens)
{
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    float v = map[i];
    v = __fmul_rn(v, premul);       // mapvec[i] *= (float)(partMass/MASS_SCALE)             shtpoissonsolve.c:426
    v = __fmul_rn(v, densmul);      // mapvec[i] *= (float)(densfact/area*MASS_SCALE)        shtpoissonsolve.c:468
    v = __fsub_rn(v, backdens);     // mapvec[i] -= (float) backdens                          shtpoissonsolve.c:478
    map[i] = v;
  }
}
// Host helper: which ranks need the map pixels of every coarse NEST cell.  Rank q traces the rays whose NEST index at
// ray_order lies in [Nray q / nranks, Nray (q+1) / nranks) (cf. loadbalance.c:151-181); its rays stay within the halo
// of that domain (the reference's MAPBUFF cells, raytrace_utils.c:116-161), so it needs every cell whose centre is
// within margin_rad of the centre of a cell of its domain (margin_rad must include two coarse cell radii).
void domain_masks(long ray_order, int nranks, long coarse_order, double margin_rad, unsigned char *mask)
{
  if (nranks > 8 || coarse_order < 0 || coarse_order > 8) { fprintf(stderr, "calclens_b200: clb_domain_masks arguments\n"); abort(); }
  const long nc = 12L << (2 * coarse_order);
  const long nray = 12L << (2 * ray_order);
  std::vector<double> cen(3 * nc);
  std::vector<unsigned char> own(nc, 0);
  for (long c = 0; c < nc; ++c) {
    nest2vec(c, coarse_order, &cen[3 * c]);
    long lo, hi;   // ray NEST range covered by (or covering) the cell
    if (ray_order >= coarse_order) { lo = c << (2 * (ray_order - coarse_order)); hi = (c + 1) << (2 * (ray_order - coarse_order)); }
    else { lo = c >> (2 * (coarse_order - ray_order)); hi = lo + 1; }
    for (int q = 0; q < nranks; ++q) {
      const long qlo = (nray * q) / nranks, qhi = (nray * (q + 1)) / nranks;
      if (lo < qhi && qlo < hi) own[c] |= (unsigned char)(1u << q);
    }
  }
  const double cosm = cos(margin_rad);
  for (long c = 0; c < nc; ++c) {
    unsigned m = own[c];
    const double *a = &cen[3 * c];
    for (long d = 0; d < nc && m != ((1u << nranks) - 1u); ++d) {
      if ((own[d] | m) == m) continue;
      const double *b = &cen[3 * d];
      if (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] >= cosm) m |= own[d];
    }
    mask[c] = (unsigned char)m;
  }
}

}  // namespace clb

using namespace clb;


static ShtPlan *P(const clb_sht_plan *plan)
{
  if (!plan || !plan->p) { fprintf(stderr, "calclens_b200: NULL plan\n"); abort(); }
  return plan->p;
}

extern "C" {

int clb_abi_version(void) { return 2; }

int clb_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    fprintf(stderr, "calclens_b200: no CUDA device available (%s); this library has no CPU fallback\n",
            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    abort();
  }
  return n;
}
void clb_set_device(int device) { CLB_CUDA_CHECK(cudaSetDevice(device)); }
int clb_host_register(void *p, long bytes)
{
  cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }   // not fatal: transfers fall back to pageable copies
  return 0;
}
void clb_host_unregister(void *p) { cudaHostUnregister(p); cudaGetLastError(); }
void clb_pool_release(void);
long clb_launch_count(void) { return g_launches; }
void clb_set_tuning(int what, int value)
{
  if (what == 0 && value >= 1 && value <= 4) g_syn_rings_per_thread = value;
  if (what == 4) g_fft_force_scratch = value ? 1 : 0;
  if (what == 5 && value >= 0) g_ana_rows = value;   // partial-sum rows per m of the Legendre analysis (0 = automatic)
  if (what == 7) g_fft_streams = value ? 1 : 0;
  if (what == 12 && value >= 0 && value <= 2) g_ana_pipeline = value;
  if (what == 11 && (value == 1 || value == 2)) g_solver_shells = value;   // read by clb_solver_create
  if (what == 9 && value >= 1 && value <= 4) g_syn2_rings_per_thread = value;
  if (what == 10 && (value == 1 || value == 2 || value == 4 || value == 6 || value == 8)) g_ana2_rings_per_thread = value;
  if (what == 6 && (value == 0 || value == 1 || value == 3)) g_fft_field_groups = value;
  if (what == 8) g_fft_debug = value;   // development aid: skip phases of the ring synthesis (wrong results)
  if (what == 3 && (value == 1 || value == 2 || value == 4)) g_leg_warps_per_cta = value;
  if (what == 2 && (value == 256 || value == 512 || value == 768 || value == 1024)) g_fft_threads_big = value;   // read at plan creation
  if (what == 1 && (value == 1 || value == 2 || value == 4 || value == 6 || value == 8)) g_ana_rings_per_thread = value;
}

clb_sht_plan *clb_sht_plan_create(long order, long lmax, const double *ring_weights, int nranks, int rank,
                                  const int *rp_owner, const int *m_owner)
{
  clb_device_count();
  if (order < 0 || order > 13 || lmax < 0 || nranks < 1 || rank < 0 || rank >= nranks) {
    fprintf(stderr, "calclens_b200: bad plan arguments order=%ld lmax=%ld nranks=%d rank=%d\n", order, lmax, nranks, rank);
    abort();
  }
  clb_sht_plan *h = new clb_sht_plan();
  h->p = sht_plan_create(order, lmax, ring_weights, nranks, rank, rp_owner, m_owner);
  return h;
}
void clb_sht_plan_destroy(clb_sht_plan *plan)
{
  if (!plan) return;
  sht_plan_destroy(plan->p);
  delete plan;
}
long clb_sht_plan_query(const clb_sht_plan *plan, int what)
{
  const ShtPlan *p = P(plan);
  switch (what) {
    case 0: return p->npix;
    case 1: return p->lmax;
    case 2: return p->alm_total;
    case 3: return p->nrp_loc;
    case 4: return p->nm_loc;
    case 5: return p->g_send_total;
    case 6: return p->g_recv_total;
    case 7: return p->b_send_total;
    case 8: return p->b_recv_total;
    case 9: return p->nranks;
    case 10: return p->rank;
    default: return -1;
  }
}
void clb_sht_plan_counts(const clb_sht_plan *plan, int which, long *counts)
{
  const ShtPlan *p = P(plan);
  const std::vector<long> *v = which == 0 ? &p->g_send_count : which == 1 ? &p->g_recv_count
                             : which == 2 ? &p->b_send_count : &p->b_recv_count;
  for (int q = 0; q < p->nranks; ++q) counts[q] = (*v)[q];
}
void clb_sht_plan_local_m(const clb_sht_plan *plan, int *m_list)
{
  const ShtPlan *p = P(plan);
  for (int i = 0; i < p->nm_loc; ++i) m_list[i] = p->m_loc[i];
}
void clb_sht_plan_local_ring_pairs(const clb_sht_plan *plan, int *rp_list)
{
  const ShtPlan *p = P(plan);
  for (int i = 0; i < p->nrp_loc; ++i) rp_list[i] = p->rp_loc[i];
}

// ---- fused exchange over peer memory (one process per GPU on one NVLink/NVSwitch node) ----
void *clb_peer_alloc(long bytes)
{
  void *p = nullptr;
  CLB_CUDA_CHECK(cudaMalloc(&p, bytes > 0 ? (size_t)bytes : 16));
  return p;
}
void clb_peer_free(void *p) { if (p) CLB_CUDA_CHECK(cudaFree(p)); }
void clb_peer_export(void *p, void *handle64)
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CLB_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
}
void *clb_peer_import(const void *handle64)
{
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    fprintf(stderr, "calclens_b200: cudaIpcOpenMemHandle failed (%s): no peer access between the GPUs of this job\n",
            cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;   // the caller falls back to the NCCL all-to-all exchange
  }
  return p;
}
void clb_peer_release(void *p) { if (p) CLB_CUDA_CHECK(cudaIpcCloseMemHandle(p)); }
void clb_sht_plan_set_peers(clb_sht_plan *plan, void *const *g_send_ptrs, void *const *b_recv_ptrs)
{
  sht_plan_set_peers(P(plan), g_send_ptrs, b_recv_ptrs, 1);
}
void clb_sht_plan_set_peers_shells(clb_sht_plan *plan, void *const *g_send_ptrs, void *const *b_recv_ptrs, int nshell)
{
  sht_plan_set_peers(P(plan), g_send_ptrs, b_recv_ptrs, nshell);
}
int clb_maps_broadcast_dev(const clb_sht_plan *plan, float *const local_maps[6], float *const *peer_maps,
                           const unsigned char *need, long coarse_order, void *stream)
{
  int n = launch_maps_broadcast(P(plan), local_maps, peer_maps, need, coarse_order, (cudaStream_t)stream);
  g_launches += n; return n;
}
void clb_domain_masks(long ray_order, int nranks, long coarse_order, double margin_rad, unsigned char *mask)
{
  domain_masks(ray_order, nranks, coarse_order, margin_rad, mask);
}

int clb_ring_analysis_dev(const clb_sht_plan *plan, const float *map, double *g_send, void *stream)
{
  int n = launch_ring_analysis(P(plan), map, reinterpret_cast<double2 *>(g_send), (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_legendre_analysis_dev(clb_sht_plan *plan, const double *g_recv, double *alm_re, double *alm_im,
                              int apply_poisson_filter, void *stream)
{
  int n = launch_legendre_analysis(P(plan), reinterpret_cast<const double2 *>(g_recv), alm_re, alm_im,
                                   apply_poisson_filter, (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_legendre_synthesis_dev(clb_sht_plan *plan, const double *alm_re, const double *alm_im, double *b_send, void *stream)
{
  int n = launch_legendre_synthesis(P(plan), alm_re, alm_im, reinterpret_cast<double2 *>(b_send), (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_legendre_analysis_shells_dev(clb_sht_plan *plan, const double *g_recv, double *alm_re, double *alm_im,
                                     int apply_poisson_filter, int nshell, void *stream)
{
  int n = launch_legendre_analysis(P(plan), reinterpret_cast<const double2 *>(g_recv), alm_re, alm_im,
                                   apply_poisson_filter, (cudaStream_t)stream, nshell);
  g_launches += n; return n;
}
int clb_legendre_synthesis_shells_dev(clb_sht_plan *plan, const double *alm_re, const double *alm_im, double *b_send,
                                      int nshell, void *stream)
{
  int n = launch_legendre_synthesis(P(plan), alm_re, alm_im, reinterpret_cast<double2 *>(b_send), (cudaStream_t)stream, nshell);
  g_launches += n; return n;
}
int clb_ring_synthesis_dev(const clb_sht_plan *plan, const double *b_recv, float *const maps[6], void *stream)
{
  int n = launch_ring_synthesis(P(plan), reinterpret_cast<const double2 *>(b_recv), maps, (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_scale_density_dev(float *map, long npix, float premul, float densmul, float backdens, void *stream)
{
  if (npix <= 0) return 0;
  scale_density_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(map, npix, premul, densmul, backdens);
  CLB_CUDA_CHECK(cudaGetLastError());
  g_launches += 1; return 1;
}
int clb_load_density_dev(const clb_sht_plan *plan, const float *src, float *dst, float premul, float densmul,
                         float backdens, void *stream)
{
  int n = launch_load_density(P(plan), src, dst, premul, densmul, backdens, (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_ray_step_dev(void *rays, long nrays, const float *const maps[6], long map_order, double wp, double wpm1,
                     double wpm2, int mode, void *stream)
{
  if ((mode & 2) && !maps) { fprintf(stderr, "calclens_b200: clb_ray_step_dev mode 2 needs maps\n"); abort(); }
  int n = launch_ray_step(reinterpret_cast<Ray *>(rays), nrays, maps, map_order, wp, wpm1, wpm2, mode, (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_ray_step_ex_dev(void *rays, long nrays, const float *const maps[6], long map_order, double wp, double wpm1,
                        double wpm2, int mode, const unsigned char *need, long coarse_order, int rank, int *err,
                        double *sum6, void *stream)
{
  if ((mode & 2) && !maps) { fprintf(stderr, "calclens_b200: clb_ray_step_ex_dev mode 2 needs maps\n"); abort(); }
  int n = launch_ray_step(reinterpret_cast<Ray *>(rays), nrays, maps, map_order, wp, wpm1, wpm2, mode, (cudaStream_t)stream,
                          need, coarse_order, rank, err, sum6);
  g_launches += n; return n;
}
int clb_ray_init_dev(void *rays, long nrays, long first_nest, long ray_order, double binL_2, void *stream)
{
  int n = launch_ray_init(reinterpret_cast<Ray *>(rays), nrays, first_nest, ray_order, binL_2, (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_ray_summary_dev(const void *rays, long nrays, double *out6, void *stream)
{
  int n = launch_ray_summary(reinterpret_cast<const Ray *>(rays), nrays, out6, (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_ray_output_dev(const void *rays, void *out_rays, long nrays, long ray_order, void *stream)
{
  int n = launch_ray_output(reinterpret_cast<const Ray *>(rays), reinterpret_cast<Ray *>(out_rays), nrays, ray_order,
                            (cudaStream_t)stream);
  g_launches += n; return n;
}
int clb_deposit_ngp_dev(const float *pos, const float *mass, long nparts, long order, float *ringmap, void *stream)
{
  int n = launch_deposit_ngp(pos, mass, nparts, order, ringmap, (cudaStream_t)stream);
  g_launches += n; return n;
}
void clb_healpix_index_dev(int what, long order, long n, const long *in, const double *theta, const double *phi,
                           long *out, void *stream)
{
  launch_healpix_index(what, order, n, in, theta, phi, out, (cudaStream_t)stream);
  g_launches += 1;
}
void clb_healpix_interpol_dev(long order, long n, const double *vec, long *pix, double *wgt, void *stream)
{
  launch_healpix_interpol(order, n, vec, pix, wgt, 0, (cudaStream_t)stream);
  g_launches += 1;
}
void clb_ray_stencil_dev(long order, long n, const double *vec, long *pix, double *wgt, void *stream)
{
  launch_healpix_interpol(order, n, vec, pix, wgt, 1, (cudaStream_t)stream);
  g_launches += 1;
}

// ---------------------------------------------------------------------------------------------------------------
// host-pointer entry points (single rank)
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

static void require_single(const ShtPlan *p, const char *who)
{
  if (p->nranks != 1) { fprintf(stderr, "calclens_b200: %s needs a single-rank plan (use the _dev stages)\n", who); abort(); }
}

// Device buffers of the host-pointer entry points: grow-only, kept per (device, slot) between calls so that a host that
// calls map2alm_mpi / alm2allmaps_mpi / rayprop_sphere once per plane (or per bundle cell) does not pay cudaMalloc/cudaFree
// every time.  clb_pool_release() frees them.
struct DevPool {
  static constexpr int kSlots = 8;
  void *p[64][kSlots] = {};
  size_t cap[64][kSlots] = {};
  void *get(int slot, size_t bytes)
  {
    int dev = 0;
    CLB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || slot < 0 || slot >= kSlots) { fprintf(stderr, "calclens_b200: DevPool(%d, %d)\n", dev, slot); abort(); }
    if (bytes > cap[dev][slot]) {
      if (p[dev][slot]) CLB_CUDA_CHECK(cudaFree(p[dev][slot]));
      CLB_CUDA_CHECK(cudaMalloc(&p[dev][slot], bytes));
      cap[dev][slot] = bytes;
    }
    return p[dev][slot];
  }
  void release()
  {
    for (int d = 0; d < 64; ++d)
      for (int k = 0; k < kSlots; ++k)
        if (p[d][k]) { cudaSetDevice(d); cudaFree(p[d][k]); p[d][k] = nullptr; cap[d][k] = 0; }
  }
};
static DevPool g_pool;
struct DevBuf {   // a view of one pool slot
  void *p = nullptr;
  DevBuf(int slot, size_t bytes) { p = g_pool.get(slot, bytes ? bytes : 16); }
  template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

extern "C" {

void clb_map2alm(clb_sht_plan *plan, const float *ringmap, double *alm_re, double *alm_im, int apply_poisson_filter)
{
  ShtPlan *p = P(plan);
  require_single(p, "clb_map2alm");
  DevBuf map(0, sizeof(float) * p->npix), g(1, sizeof(double2) * p->g_send_total), are(2, sizeof(double) * p->alm_total),
      aim(3, sizeof(double) * p->alm_total);
  CLB_CUDA_CHECK(cudaMemcpy(map.p, ringmap, sizeof(float) * p->npix, cudaMemcpyHostToDevice));
  clb_ring_analysis_dev(plan, map.as<float>(), g.as<double>(), nullptr);
  clb_legendre_analysis_dev(plan, g.as<double>(), are.as<double>(), aim.as<double>(), apply_poisson_filter, nullptr);
  CLB_CUDA_CHECK(cudaMemcpy(alm_re, are.p, sizeof(double) * p->alm_total, cudaMemcpyDeviceToHost));
  CLB_CUDA_CHECK(cudaMemcpy(alm_im, aim.p, sizeof(double) * p->alm_total, cudaMemcpyDeviceToHost));
}

void clb_alm2allmaps(clb_sht_plan *plan, const double *alm_re, const double *alm_im, float *maps)
{
  ShtPlan *p = P(plan);
  require_single(p, "clb_alm2allmaps");
  DevBuf dm(0, sizeof(float) * 6 * p->npix), b(4, sizeof(double2) * p->b_send_total), are(2, sizeof(double) * p->alm_total),
      aim(3, sizeof(double) * p->alm_total);
  CLB_CUDA_CHECK(cudaMemcpy(are.p, alm_re, sizeof(double) * p->alm_total, cudaMemcpyHostToDevice));
  CLB_CUDA_CHECK(cudaMemcpy(aim.p, alm_im, sizeof(double) * p->alm_total, cudaMemcpyHostToDevice));
  float *mp[6];
  for (int k = 0; k < 6; ++k) mp[k] = dm.as<float>() + (size_t)k * p->npix;
  clb_legendre_synthesis_dev(plan, are.as<double>(), aim.as<double>(), b.as<double>(), nullptr);
  clb_ring_synthesis_dev(plan, b.as<double>(), mp, nullptr);
  CLB_CUDA_CHECK(cudaMemcpy(maps, dm.p, sizeof(float) * 6 * p->npix, cudaMemcpyDeviceToHost));
}

}  // extern "C"

static void mapvec_to_ring(const ShtPlan *p, const float *mapvec, const long *ns, const long *ss, float *ring)
{
  for (int rp = 0; rp < p->nrp; ++rp) {
    const size_t n = (size_t)p->h_nphi[rp];
    memcpy(ring + p->h_startN[rp], mapvec + 2 * ns[rp], sizeof(float) * n);
    if (p->h_startS[rp] >= 0) memcpy(ring + p->h_startS[rp], mapvec + 2 * ss[rp], sizeof(float) * n);
  }
}
static void ring_to_mapvec(const ShtPlan *p, const float *ring, const long *ns, const long *ss, float *mapvec)
{
  for (int rp = 0; rp < p->nrp; ++rp) {
    const size_t n = (size_t)p->h_nphi[rp];
    memcpy(mapvec + 2 * ns[rp], ring + p->h_startN[rp], sizeof(float) * n);
    if (p->h_startS[rp] >= 0) memcpy(mapvec + 2 * ss[rp], ring + p->h_startS[rp], sizeof(float) * n);
  }
}

extern "C" {

void clb_map2alm_mapvec(clb_sht_plan *plan, float *mapvec, const long *north_start, const long *south_start,
                        double *alm_re, double *alm_im)
{
  ShtPlan *p = P(plan);
  require_single(p, "clb_map2alm_mapvec");
  std::vector<float> ring(p->npix);
  mapvec_to_ring(p, mapvec, north_start, south_start, ring.data());
  clb_map2alm(plan, ring.data(), alm_re, alm_im, 0);
}

void clb_alm2allmaps_mapvec(clb_sht_plan *plan, const double *alm_re, const double *alm_im, float *const mapvec[6],
                            const long *north_start, const long *south_start)
{
  ShtPlan *p = P(plan);
  require_single(p, "clb_alm2allmaps_mapvec");
  std::vector<float> maps((size_t)6 * p->npix);
  clb_alm2allmaps(plan, alm_re, alm_im, maps.data());
  for (int k = 0; k < 6; ++k) ring_to_mapvec(p, maps.data() + (size_t)k * p->npix, north_start, south_start, mapvec[k]);
}

void clb_ray_step(void *rays, long nrays, const float *maps, long map_order, double wp, double wpm1, double wpm2, int mode)
{
  clb_device_count();
  const long npix = 12L << (2 * map_order);
  DevBuf dr(5, sizeof(Ray) * nrays), dm(0, (mode & 2) ? sizeof(float) * 6 * npix : 16);
  CLB_CUDA_CHECK(cudaMemcpy(dr.p, rays, sizeof(Ray) * nrays, cudaMemcpyHostToDevice));
  const float *mp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (mode & 2) {
    if (!maps) { fprintf(stderr, "calclens_b200: clb_ray_step mode 2 needs maps\n"); abort(); }
    CLB_CUDA_CHECK(cudaMemcpy(dm.p, maps, sizeof(float) * 6 * npix, cudaMemcpyHostToDevice));
    for (int k = 0; k < 6; ++k) mp[k] = dm.as<float>() + (size_t)k * npix;
  }
  clb_ray_step_dev(dr.p, nrays, mp, map_order, wp, wpm1, wpm2, mode, nullptr);
  CLB_CUDA_CHECK(cudaMemcpy(rays, dr.p, sizeof(Ray) * nrays, cudaMemcpyDeviceToHost));
}

void clb_lens_plane(clb_sht_plan *plan, const float *ringmap, float premul, float densmul, float backdens, void *rays,
                    long nrays, double wp, double wpm1, double wpm2)
{
  ShtPlan *p = P(plan);
  require_single(p, "clb_lens_plane");
  DevBuf dm(0, sizeof(float) * 6 * p->npix), g(1, sizeof(double2) * p->g_send_total), b(4, sizeof(double2) * p->b_send_total),
      are(2, sizeof(double) * p->alm_total), aim(3, sizeof(double) * p->alm_total), dr(5, sizeof(Ray) * nrays);
  float *mp[6];
  for (int k = 0; k < 6; ++k) mp[k] = dm.as<float>() + (size_t)k * p->npix;
  CLB_CUDA_CHECK(cudaMemcpyAsync(mp[0], ringmap, sizeof(float) * p->npix, cudaMemcpyHostToDevice, nullptr));
  CLB_CUDA_CHECK(cudaMemcpyAsync(dr.p, rays, sizeof(Ray) * nrays, cudaMemcpyHostToDevice, nullptr));
  clb_scale_density_dev(mp[0], p->npix, premul, densmul, backdens, nullptr);
  clb_ring_analysis_dev(plan, mp[0], g.as<double>(), nullptr);
  clb_legendre_analysis_dev(plan, g.as<double>(), are.as<double>(), aim.as<double>(), 1, nullptr);
  clb_legendre_synthesis_dev(plan, are.as<double>(), aim.as<double>(), b.as<double>(), nullptr);
  clb_ring_synthesis_dev(plan, b.as<double>(), mp, nullptr);
  clb_ray_step_dev(dr.p, nrays, mp, p->order, wp, wpm1, wpm2, 1 | 2 | 4, nullptr);
  CLB_CUDA_CHECK(cudaMemcpy(rays, dr.p, sizeof(Ray) * nrays, cudaMemcpyDeviceToHost));
}

void clb_pool_release(void) { g_pool.release(); }

}  // extern "C"
