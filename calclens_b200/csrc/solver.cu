// calclens_b200/csrc/solver.cu -- persistent lens-plane solver below the C ABI (clb_solver_* in include/calclens_b200.h).
// One instance per rank / GPU.  It owns everything that lives across planes -- the SHT plan, the exchange buffers, the
// six derivative maps, the double-buffered density, the device-resident rays -- and strings the stages together:
//   density load (shtpoissonsolve.c:342-502) -> map2alm_mpi -> Poisson filter (:526-550) -> alm2allmaps_mpi ->
//   interpolation at the rays (:666-702) -> rayprop_sphere for every ray (raytrace.c:256-269).
// Multi-rank jobs (one process per GPU on one NVLink/NVSwitch node) use the fused exchange: the buffers of all ranks are
// mapped into every process (CUDA IPC; the handles travel through a caller-supplied allgather, MPI_Allgather in
// CALCLENS) and the producing kernels store straight into the consumer's memory.  The producer/consumer stages are
// ordered by a device-side barrier over peer memory (peer_barrier_kernel: one flag word per rank pair), so no host
// synchronisation and no communication library sits on the per-plane path.
#include "launch.cuh"
#include "../../include/calclens_b200.h"
#include <math.h>
#include <string.h>
#include <unistd.h>
#include <vector>

namespace clb {

constexpr int kMaxPeers = 8;          // ring_broadcast_kernel addresses up to 8 ranks (one NVSwitch node)
constexpr int kCoarseOrder = 5;       // halo masks live on a NEST grid of 12 * 4^5 cells
constexpr int kStages = 10;           // stage boundaries recorded per step (timing)
int g_solver_shells = 2;              // clb_set_tuning(11, 1|2): shells per SHT pass a solver is provisioned for (read at creation)

struct PeerFlags { unsigned *p[kMaxPeers]; };

// kernels launched: per solver and in the library-wide counter (clb_launch_count)
struct LaunchAdder {
  long *local;
  void operator+=(int n) { *local += n; count_launches(n); }
};
#define LAUNCHED(s) LaunchAdder{&(s)->launches} +=

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Stream-ordered barrier across the ranks of one node: thread q tells rank q that this rank has reached `epoch`
// (release store into q's flag array over NVLink) and waits until rank q has said the same here.  Everything enqueued
// before the barrier on any rank has completed (kernel boundary) before anything enqueued after it starts.
// A rank that never arrives (dead process) would hang the node: after 30 s the wait gives up and raises bit 2 of *err.
__global__ void peer_barrier_kernel(PeerFlags peers, unsigned *mine, int nranks, int rank, unsigned epoch, int *err)
{
  const int q = threadIdx.x;
  if (q < nranks) {
    __threadfence_system();
    st_release_sys(peers.p[q] + rank, epoch);
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(mine + q) - epoch) < 0) {
      if (global_ns() - t0 > 30000000000ull) { atomicOr(err, 2); break; }
    }
  }
  __syncthreads();
}

struct PeerCard {            // what every rank publishes through the allgather
  int pid, device, ok, pad;
  cudaIpcMemHandle_t handle[4];
  void *raw[4];              // same-process ranks (emulation, tests) use the pointers directly
};

struct Solver {
  int nranks = 1, rank = 0;
  long order = 0, lmax = 0, ray_order = 0, npix = 0;
  clb_sht_plan *plan_h = nullptr;
  ShtPlan *plan = nullptr;
  clb_allgather_fn allgather = nullptr;
  void *ctx = nullptr;
  int host_barriers = 0;
  bool fused = false;
  // buffers
  double2 *g_send = nullptr, *g_recv = nullptr, *b_send = nullptr, *b_recv = nullptr;
  double *alm_re = nullptr, *alm_im = nullptr;
  float *maps = nullptr;
  float *dens[2] = {nullptr, nullptr};
  Ray *rays = nullptr;
  long nrays = 0, first_nest = 0, rays_cap = 0;
  double *d_sum6 = nullptr;
  int *d_err = nullptr;
  unsigned char *d_need = nullptr;
  unsigned char *d_gmask = nullptr;    // d_need evaluated per group of four ring pixels of this rank's rings (map broadcast)
  double need_fraction = 1.0, halo_deg = 0.0;
  double *h_sum6 = nullptr;   // pinned: 6 sums + err word
  // peers
  void *own[4] = {nullptr, nullptr, nullptr, nullptr};   // g send, b receive, maps, flags
  std::vector<std::vector<void *>> peer;                  // [rank][4]
  std::vector<char> peer_is_ipc;
  unsigned epoch = 0;
  // streams / staging
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t dens_ready[2] = {nullptr, nullptr}, dens_free[2] = {nullptr, nullptr}, ev_tmp = nullptr;
  bool dens_free_valid[2] = {false, false};
  int cur = 0;                       // density buffer of the plane being solved
  struct PlaneKey {                  // a plane as the caller names it: source pointer + the three scalings
    const void *src = nullptr; float scal[3] = {0, 0, 0};
    bool is(const void *p, float a, float b, float c) const { return src && src == p && scal[0] == a && scal[1] == b && scal[2] == c; }
    void set(const void *p, float a, float b, float c) { src = p; scal[0] = a; scal[1] = b; scal[2] = c; }
  };
  PlaneKey staged[2];                // density buffer k holds this prefetched plane (src == nullptr: nothing staged)
  float *h_stage[2] = {nullptr, nullptr};   // pinned staging for pageable host maps
  float *d_raw[2] = {nullptr, nullptr};     // single rank: a host map crosses PCIe whole, on a copy engine (no SM slots), then is scaled
  PlaneKey next[2];                  // planes registered by clb_solver_set_next: prefetched behind the current step's kernels
  int n_next = 0;
  // two planes per SHT pass: the partner registered by clb_solver_set_pair is solved together with
  // the next step's plane; its six maps wait in the second map set until the step that names it
  int shells = 2;                    // shells the exchange buffers and map sets are provisioned for (clb_set_tuning(11, .))
  PlaneKey pair, cached;             // partner to solve with the next step / plane whose maps are in map set 1
  // timing
  int timing = 0;
  cudaEvent_t ev[kStages + 1] = {};
  long launches = 0;
};

static void die(const char *msg)
{
  fprintf(stderr, "calclens_b200: %s\n", msg);
  abort();   // the reference's failure mode on this path is MPI_Abort(MPI_COMM_WORLD, 123)
}

static void host_barrier(Solver *s)
{
  if (s->nranks == 1 || !s->allgather) return;
  std::vector<char> in(8, 0), out(8 * (size_t)s->nranks, 0);
  s->allgather(in.data(), out.data(), 8, s->ctx);
}

static void stream_barrier(Solver *s, cudaStream_t st)
{
  if (s->nranks == 1) return;
  if (s->host_barriers) {   // ranks that time-share one GPU cannot spin on each other inside kernels
    CLB_CUDA_CHECK(cudaStreamSynchronize(st));
    host_barrier(s);
    return;
  }
  PeerFlags pf;
  for (int q = 0; q < kMaxPeers; ++q) pf.p[q] = q < s->nranks ? reinterpret_cast<unsigned *>(s->peer[q][3]) : nullptr;
  ++s->epoch;
  peer_barrier_kernel<<<1, 32, 0, st>>>(pf, reinterpret_cast<unsigned *>(s->own[3]), s->nranks, s->rank, s->epoch, s->d_err);
  CLB_CUDA_CHECK(cudaGetLastError());
  LAUNCHED(s) 1;
}

static void mark(Solver *s, int k, cudaStream_t st)
{
  if (s->timing) CLB_CUDA_CHECK(cudaEventRecord(s->ev[k], st));
}

// Ring pairs are dealt to ranks round-robin in groups of four adjacent pairs (128-byte runs in the exchange buffers), m
// round-robin: every rank sees all latitudes and all m magnitudes, which balances the FFT and the Legendre stage without
// the reference's cost polynomials (healpix_shtrans.c:219-250, :597-626).
static void default_owners(long order, long lmax, int nranks, std::vector<int> &rp_owner, std::vector<int> &m_owner)
{
  const int nrp = (int)(2L << order);
  const int group = (nrp >= 32 * nranks) ? 4 : 1;
  rp_owner.resize(nrp); m_owner.resize(lmax + 1);
  for (int rp = 0; rp < nrp; ++rp) rp_owner[rp] = (rp / group) % nranks;
  for (long m = 0; m <= lmax; ++m) m_owner[m] = (int)(m % nranks);
}

static bool setup_peers(Solver *s)
{
  ShtPlan *p = s->plan;
  const size_t sh = (size_t)s->shells;
  const size_t sizes[4] = {16 * sh * (size_t)std::max<long>(p->g_send_total, 1), 16 * sh * (size_t)std::max<long>(p->b_recv_total, 1),
                           4 * 6 * sh * (size_t)s->npix, 256};
  PeerCard mine;
  memset(&mine, 0, sizeof(mine));
  mine.pid = (int)getpid();
  CLB_CUDA_CHECK(cudaGetDevice(&mine.device));
  for (int k = 0; k < 4; ++k) {
    CLB_CUDA_CHECK(cudaMalloc(&s->own[k], sizes[k]));
    CLB_CUDA_CHECK(cudaMemset(s->own[k], 0, k >= 2 ? sizes[k] : 16));
    CLB_CUDA_CHECK(cudaIpcGetMemHandle(&mine.handle[k], s->own[k]));
    mine.raw[k] = s->own[k];
  }
  CLB_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<PeerCard> cards(s->nranks);
  s->allgather(&mine, cards.data(), (long)sizeof(PeerCard), s->ctx);
  s->peer.assign(s->nranks, std::vector<void *>(4, nullptr));
  s->peer_is_ipc.assign(s->nranks, 0);
  int ok = 1, shared_device = 0;
  for (int q = 0; q < s->nranks; ++q) {
    if (q == s->rank) { for (int k = 0; k < 4; ++k) s->peer[q][k] = s->own[k]; continue; }
    if (cards[q].device == mine.device) shared_device = 1;   // time-shared GPU (or one stream shared by emulated ranks)
    if (cards[q].pid == mine.pid) {      // ranks emulated inside one process: plain pointers
      if (cards[q].device != mine.device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(cards[q].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
        cudaGetLastError();
      }
      for (int k = 0; k < 4; ++k) s->peer[q][k] = cards[q].raw[k];
      continue;
    }
    s->peer_is_ipc[q] = 1;
    for (int k = 0; k < 4; ++k) {
      void *ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, cards[q].handle[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        fprintf(stderr, "calclens_b200: cudaIpcOpenMemHandle failed (%s): no peer access between the GPUs of this job\n",
                cudaGetErrorString(e));
        cudaGetLastError();
        ok = 0; ptr = nullptr;
      }
      s->peer[q][k] = ptr;
    }
  }
  // every rank must agree on the outcome (and on whether several ranks time-share one GPU)
  int flag[2] = {ok, shared_device};
  std::vector<int> flags(2 * (size_t)s->nranks);
  s->allgather(flag, flags.data(), (long)sizeof(flag), s->ctx);
  for (int q = 0; q < s->nranks; ++q) { ok &= flags[2 * q]; shared_device |= flags[2 * q + 1]; }
  if (!ok) return false;
  if (shared_device) s->host_barriers = 1;
  std::vector<void *> gp(s->nranks), bp(s->nranks);
  for (int q = 0; q < s->nranks; ++q) { gp[q] = s->peer[q][0]; bp[q] = s->peer[q][1]; }
  sht_plan_set_peers(p, gp.data(), bp.data(), s->shells);
  s->g_send = reinterpret_cast<double2 *>(s->own[0]);
  s->b_recv = reinterpret_cast<double2 *>(s->own[1]);
  s->maps = reinterpret_cast<float *>(s->own[2]);
  return true;
}

static void release_peers(Solver *s)
{
  if (!s->own[0]) return;
  CLB_CUDA_CHECK(cudaDeviceSynchronize());
  host_barrier(s);     // every rank has stopped using the mappings
  for (int q = 0; q < (int)s->peer.size(); ++q)
    if (q != s->rank && s->peer_is_ipc[q])
      for (void *ptr : s->peer[q]) if (ptr) cudaIpcCloseMemHandle(ptr);
  host_barrier(s);     // nobody still maps the memory that is about to be freed
  for (int k = 0; k < 4; ++k) { cudaFree(s->own[k]); s->own[k] = nullptr; }
}

static const float *device_view(Solver *s, const float *src, int slot, cudaStream_t st)
{
  // device memory and pinned/registered host memory are read by the load kernel directly (only this rank's rings cross
  // PCIe); pageable host memory is first copied into a pinned staging buffer
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, src);
  if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return src;
  if (attr.type == cudaMemoryTypeHost) return attr.devicePointer ? reinterpret_cast<const float *>(attr.devicePointer) : src;
  if (!s->h_stage[slot]) CLB_CUDA_CHECK(cudaHostAlloc(&s->h_stage[slot], sizeof(float) * s->npix, cudaHostAllocDefault));
  // the staging buffer may still be read by the previous load on `st`
  CLB_CUDA_CHECK(cudaStreamSynchronize(st));
  memcpy(s->h_stage[slot], src, sizeof(float) * s->npix);
  return s->h_stage[slot];
}

static void load_density(Solver *s, const float *src, int k, float premul, float densmul, float backdens, cudaStream_t st)
{
  const float *v = device_view(s, src, k, st);
  if (s->nranks == 1) {
    // One rank needs every ring: a plain asynchronous copy moves the map at full PCIe rate on a copy engine while the SMs
    // compute (the kernel that reads pinned host memory directly holds SM slots for the ~15 ms the transfer takes; with
    // several ranks it is still the better choice because every rank then moves only its own rings).
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, v) != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
    if (attr.type == cudaMemoryTypeHost) {
      if (!s->d_raw[k]) CLB_CUDA_CHECK(cudaMalloc(&s->d_raw[k], sizeof(float) * s->npix));
      const void *hsrc = attr.hostPointer ? attr.hostPointer : (const void *)v;   // (the pinned staging copy for pageable maps)
      CLB_CUDA_CHECK(cudaMemcpyAsync(s->d_raw[k], hsrc, sizeof(float) * s->npix, cudaMemcpyHostToDevice, st));
      v = s->d_raw[k];
    }
  }
  LAUNCHED(s) launch_load_density(s->plan, v, s->dens[k], premul, densmul, backdens, st);
}

// density buffers (this rank's rings valid) -> six derivative maps per shell.  nshell = 2 runs the two planes through ONE
// pass of each Legendre kernel (the lambda_lm recurrence is generated once for both); the ring FFTs run per shell.  Shell s
// uses the second half of the exchange buffers and map set s.
static void solve(Solver *s, const float *const dens[2], int nshell, const int dens_buf[2], cudaStream_t st)
{
  ShtPlan *p = s->plan;
  if (nshell > s->shells) die("solver was provisioned for one shell per pass (clb_set_tuning(11, 2) before clb_solver_create)");
  float *mp[2][6];
  for (int q = 0; q < nshell; ++q)
    for (int k = 0; k < 6; ++k) mp[q][k] = s->maps + ((size_t)q * 6 + k) * s->npix;
  if (s->fused) stream_barrier(s, st);          // every rank is done with the previous plane's g, b and maps
  fft_fork(p, st);                              // the class launches of both shells share the parallel streams
  for (int q = 0; q < nshell; ++q) LAUNCHED(s) launch_ring_analysis(p, dens[q], s->g_send + (size_t)q * p->g_send_total, st);
  fft_join(p, st);
  for (int q = 0; q < nshell; ++q)
    if (dens_buf && dens_buf[q] >= 0) {         // the density buffer is free again: a prefetch may overwrite it
      CLB_CUDA_CHECK(cudaEventRecord(s->dens_free[dens_buf[q]], st));
      s->dens_free_valid[dens_buf[q]] = true;
    }
  mark(s, 2, st);
  if (s->fused) stream_barrier(s, st);
  mark(s, 3, st);
  LAUNCHED(s) launch_legendre_analysis(p, s->fused ? nullptr : s->g_recv, s->alm_re, s->alm_im, 1, st, nshell); mark(s, 4, st);
  LAUNCHED(s) launch_legendre_synthesis(p, s->alm_re, s->alm_im, s->fused ? nullptr : s->b_send, st, nshell); mark(s, 5, st);
  if (s->fused) stream_barrier(s, st);
  mark(s, 6, st);
  fft_fork(p, st);
  for (int q = 0; q < nshell; ++q) LAUNCHED(s) launch_ring_synthesis(p, s->b_recv + (size_t)q * p->b_recv_total, mp[q], st);
  fft_join(p, st);
  mark(s, 7, st);
  if (s->fused) {
    for (int q = 0; q < nshell; ++q) {
      float *pm[kMaxPeers * 6];
      for (int r = 0; r < s->nranks; ++r)
        for (int k = 0; k < 6; ++k) pm[r * 6 + k] = reinterpret_cast<float *>(s->peer[r][2]) + ((size_t)q * 6 + k) * s->npix;
      LAUNCHED(s) launch_maps_broadcast(p, mp[q], pm, s->d_need, kCoarseOrder, st, s->d_gmask);
    }
    stream_barrier(s, st);
  }
  mark(s, 8, st);
}
static void solve(Solver *s, const float *dens, cudaStream_t st)
{
  const float *d[2] = {dens, nullptr};
  solve(s, d, 1, nullptr, st);
}

static void ray_update(Solver *s, double wpp1, double wp, double wpm1, int mode, bool with_summary, cudaStream_t st, int set = 0)
{
  const float *mp[6];
  for (int k = 0; k < 6; ++k) mp[k] = s->maps + ((size_t)set * 6 + k) * s->npix;
  LAUNCHED(s) launch_ray_step(s->rays, s->nrays, mp, s->order, wpp1, wp, wpm1, mode, st, s->d_need, kCoarseOrder, s->rank,
                                 s->d_need ? s->d_err : nullptr, with_summary ? s->d_sum6 : nullptr,
                                 s->d_need ? s->d_need + (12L << (2 * kCoarseOrder)) : nullptr);
  mark(s, 9, st);
}


// The reference's padded ring-pair buffers ("mapvec", healpix_shtrans.c:90-118: local ring pair i occupies ringpix+2
// floats at 2*north_start[i], its mirror at 2*south_start[i], -1 for the equator's missing mirror) <-> this rank's rings
// of a RING-ordered device map.  One CTA per (local ring pair, hemisphere).
__global__ void mapvec_copy_kernel(float *__restrict__ mapvec, float *__restrict__ ring, const long *__restrict__ ns,
                                   const long *__restrict__ ss, const int *__restrict__ rp_loc, const int *__restrict__ nphi,
                                   const long *__restrict__ startN, const long *__restrict__ startS, int to_ring)
{
  const int i = blockIdx.x >> 1, hemi = blockIdx.x & 1;
  const int rp = rp_loc[i];
  const long off = hemi ? ss[i] : ns[i], start = hemi ? startS[rp] : startN[rp];
  if (off < 0 || start < 0) return;
  const int n = nphi[rp];
  float *a = mapvec + 2 * off, *b = ring + start;
  if (to_ring) for (int k = threadIdx.x; k < n; k += blockDim.x) b[k] = a[k];
  else for (int k = threadIdx.x; k < n; k += blockDim.x) a[k] = b[k];
}

struct MapvecScratch {   // grow-only device scratch of the mapvec entry points
  float *mv = nullptr; size_t mv_cap = 0;
  long *idx = nullptr; size_t idx_cap = 0;
};
static MapvecScratch g_mvs[64];
static MapvecScratch &mapvec_scratch(size_t mv_bytes, size_t idx_bytes)
{
  int dev = 0;
  CLB_CUDA_CHECK(cudaGetDevice(&dev));
  MapvecScratch &m = g_mvs[dev & 63];
  if (mv_bytes > m.mv_cap) { if (m.mv) cudaFree(m.mv); CLB_CUDA_CHECK(cudaMalloc(&m.mv, mv_bytes)); m.mv_cap = mv_bytes; }
  if (idx_bytes > m.idx_cap) { if (m.idx) cudaFree(m.idx); CLB_CUDA_CHECK(cudaMalloc(&m.idx, idx_bytes)); m.idx_cap = idx_bytes; }
  return m;
}
// number of 8-byte units of this rank's mapvec (Nmapvec of healpix_shtrans.h:33)
static long mapvec_units(const ShtPlan *p)
{
  long n = 0;
  for (int rp : p->rp_loc) n += (long)(p->h_nphi[rp] / 2 + 1) * (p->h_startS[rp] >= 0 ? 2 : 1);
  return n;
}

}  // namespace clb

using namespace clb;

struct clb_solver { Solver s; };

extern "C" {

clb_solver *clb_solver_create(long sht_order, long lmax, long ray_order, const double *ring_weights, int nranks, int rank,
                              const int *rp_owner, const int *m_owner, clb_allgather_fn allgather, void *ctx, double halo_deg)
{
  clb_device_count();
  if (nranks > 1 && !allgather) die("clb_solver_create: a multi-rank solver needs an allgather callback");
  if (nranks > kMaxPeers) die("clb_solver_create: the fused exchange supports up to 8 ranks (one NVSwitch node)");
  clb_solver *h = new clb_solver();
  Solver *s = &h->s;
  s->nranks = nranks; s->rank = rank; s->order = sht_order; s->lmax = lmax; s->ray_order = ray_order < 0 ? sht_order : ray_order;
  s->npix = 12L << (2 * sht_order);
  s->allgather = allgather; s->ctx = ctx; s->halo_deg = halo_deg;
  s->shells = g_solver_shells;
  std::vector<int> ro, mo;
  if (nranks > 1 && (!rp_owner || !m_owner)) { default_owners(sht_order, lmax, nranks, ro, mo); rp_owner = ro.data(); m_owner = mo.data(); }
  s->plan_h = clb_sht_plan_create(sht_order, lmax, ring_weights, nranks, rank, rp_owner, m_owner);
  s->plan = s->plan_h->p;
  ShtPlan *p = s->plan;
  auto dmalloc = [](size_t bytes) { void *q = nullptr; CLB_CUDA_CHECK(cudaMalloc(&q, bytes ? bytes : 16)); return q; };
  s->alm_re = (double *)dmalloc(sizeof(double) * s->shells * std::max<long>(p->alm_total, 1));
  s->alm_im = (double *)dmalloc(sizeof(double) * s->shells * std::max<long>(p->alm_total, 1));
  for (int k = 0; k < 2; ++k) {
    s->dens[k] = (float *)dmalloc(sizeof(float) * s->npix);
    CLB_CUDA_CHECK(cudaMemset(s->dens[k], 0, sizeof(float) * s->npix));
    CLB_CUDA_CHECK(cudaEventCreateWithFlags(&s->dens_ready[k], cudaEventDisableTiming));
    CLB_CUDA_CHECK(cudaEventCreateWithFlags(&s->dens_free[k], cudaEventDisableTiming));
  }
  CLB_CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_tmp, cudaEventDisableTiming));
  s->d_sum6 = (double *)dmalloc(sizeof(double) * 8);
  s->d_err = (int *)dmalloc(sizeof(int) * 4);
  CLB_CUDA_CHECK(cudaMemset(s->d_sum6, 0, sizeof(double) * 8));
  CLB_CUDA_CHECK(cudaMemset(s->d_err, 0, sizeof(int) * 4));
  CLB_CUDA_CHECK(cudaHostAlloc(&s->h_sum6, sizeof(double) * 8, cudaHostAllocDefault));
  int lo = 0, hi = 0;
  CLB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CLB_CUDA_CHECK(cudaStreamCreateWithPriority(&s->copy_stream, cudaStreamNonBlocking, hi));   // gets SM slots first
  for (int k = 0; k <= kStages; ++k) CLB_CUDA_CHECK(cudaEventCreate(&s->ev[k]));
  if (nranks > 1) {
    if (!setup_peers(s)) {
      fprintf(stderr, "calclens_b200: clb_solver_create: peer mapping unavailable, no fused exchange possible\n");
      release_peers(s);
      clb_solver_destroy(h);
      return nullptr;   // the caller may fall back to an exchange of its own over the _dev stage API
    }
    s->fused = true;
    // halo-limited map broadcast: a pixel goes to the ranks whose ray domain, grown by halo_deg, can reach it.  Cells of a
    // coarse NEST grid stand in for the reference's halo bundle cells (raytrace_utils.c:116-161); the margin adds two
    // coarse cell radii (a HEALPix pixel's radius is < 1.2 x its mean spacing).
    if (halo_deg > 0 && sht_order >= kCoarseOrder) {
      const long nc = 12L << (2 * kCoarseOrder);
      const double spacing = sqrt(4.0 * CLB_PI / (double)nc);
      std::vector<unsigned char> mask(nc);
      domain_masks(s->ray_order, nranks, kCoarseOrder, halo_deg * CLB_PI / 180.0 + 2.0 * 1.2 * spacing, mask.data());
      // second half of the buffer: cells whose whole neighbourhood (2.5 cell spacings: every cell a stencil can reach) is delivered
      std::vector<unsigned char> safe(nc);
      safe_masks(kCoarseOrder, 2.5 * spacing, mask.data(), safe.data());
      s->d_need = (unsigned char *)dmalloc(2 * nc);
      CLB_CUDA_CHECK(cudaMemcpy(s->d_need, mask.data(), nc, cudaMemcpyHostToDevice));
      CLB_CUDA_CHECK(cudaMemcpy(s->d_need + nc, safe.data(), nc, cudaMemcpyHostToDevice));
      s->d_gmask = (unsigned char *)dmalloc((size_t)(s->npix / 4));
      LAUNCHED(s) launch_group_masks(p, s->d_need, kCoarseOrder, s->d_gmask, nullptr);
      long bits = 0;
      for (long c = 0; c < nc; ++c) bits += __builtin_popcount(mask[c]);
      s->need_fraction = (double)bits / ((double)nc * nranks);
    }
    host_barrier(s);
  } else {
    s->g_send = (double2 *)dmalloc(sizeof(double2) * s->shells * std::max<long>(p->g_send_total, 1));
    s->b_send = (double2 *)dmalloc(sizeof(double2) * s->shells * std::max<long>(p->b_send_total, 1));
    s->g_recv = s->g_send; s->b_recv = s->b_send;
    s->maps = (float *)dmalloc(sizeof(float) * 6 * s->shells * s->npix);
    CLB_CUDA_CHECK(cudaMemset(s->maps, 0, sizeof(float) * 6 * s->shells * s->npix));
  }
  CLB_CUDA_CHECK(cudaDeviceSynchronize());
  return h;
}

void clb_solver_destroy(clb_solver *h)
{
  if (!h) return;
  Solver *s = &h->s;
  cudaDeviceSynchronize();
  if (s->fused) release_peers(s);
  else if (s->nranks == 1) { cudaFree(s->g_send); cudaFree(s->b_send); cudaFree(s->maps); }
  cudaFree(s->alm_re); cudaFree(s->alm_im); cudaFree(s->dens[0]); cudaFree(s->dens[1]); cudaFree(s->rays);
  cudaFree(s->d_sum6); cudaFree(s->d_err); cudaFree(s->d_need); cudaFree(s->d_gmask);
  cudaFreeHost(s->h_sum6); cudaFreeHost(s->h_stage[0]); cudaFreeHost(s->h_stage[1]); cudaFree(s->d_raw[0]); cudaFree(s->d_raw[1]);
  for (int k = 0; k < 2; ++k) { if (s->dens_ready[k]) cudaEventDestroy(s->dens_ready[k]); if (s->dens_free[k]) cudaEventDestroy(s->dens_free[k]); }
  if (s->ev_tmp) cudaEventDestroy(s->ev_tmp);
  for (int k = 0; k <= kStages; ++k) if (s->ev[k]) cudaEventDestroy(s->ev[k]);
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  clb_sht_plan_destroy(s->plan_h);
  delete h;
}

long clb_solver_query(const clb_solver *h, int what)
{
  const Solver *s = &h->s;
  switch (what) {
    case 0: return s->nrays;
    case 1: return s->first_nest;
    case 2: return s->fused ? 1 : 0;
    case 3: return s->launches;
    case 4: return s->d_need ? 1 : 0;
    case 5: return (long)(s->need_fraction * 1e6);
    case 6: return s->host_barriers;
    case 7: return s->npix;
    case 8: return s->shells;
    default: return -1;
  }
}

void *clb_solver_ptr(clb_solver *h, int what)
{
  Solver *s = &h->s;
  switch (what) {
    case 0: return s->maps;
    case 1: return s->rays;
    case 2: return s->alm_re;
    case 3: return s->alm_im;
    case 4: return s->plan_h;
    case 5: return s->d_need;
    case 6: return s->dens[0];
    case 7: return s->dens[1];
    case 8: return s->d_sum6;
    case 9: return s->shells >= 2 ? s->maps + 6 * (size_t)s->npix : nullptr;
    default: return nullptr;
  }
}

void clb_solver_set_timing(clb_solver *h, int on) { h->s.timing = on ? 1 : 0; }

void clb_solver_stage_ms(clb_solver *h, double *ms9)
{
  Solver *s = &h->s;
  CLB_CUDA_CHECK(cudaEventSynchronize(s->ev[kStages - 1]));
  for (int k = 0; k < kStages - 1; ++k) {
    float t = 0.f;
    CLB_CUDA_CHECK(cudaEventElapsedTime(&t, s->ev[k], s->ev[k + 1]));
    ms9[k] = t;
  }
}

long clb_solver_init_rays(clb_solver *h, double binL_2, void *stream)
{
  Solver *s = &h->s;
  const long tot = 12L << (2 * s->ray_order);
  const long lo = (tot * s->rank) / s->nranks, hi = (tot * (s->rank + 1)) / s->nranks;
  s->first_nest = lo; s->nrays = hi - lo;
  if (s->nrays > s->rays_cap) {
    if (s->rays) CLB_CUDA_CHECK(cudaFree(s->rays));
    CLB_CUDA_CHECK(cudaMalloc(&s->rays, sizeof(Ray) * (size_t)s->nrays));
    s->rays_cap = s->nrays;
  }
  LAUNCHED(s) launch_ray_init(s->rays, s->nrays, lo, s->ray_order, binL_2, (cudaStream_t)stream);
  return s->nrays;
}

void clb_solver_set_rays(clb_solver *h, const void *host_rays, long nrays, void *stream)
{
  Solver *s = &h->s;
  if (nrays > s->rays_cap) {
    if (s->rays) CLB_CUDA_CHECK(cudaFree(s->rays));
    CLB_CUDA_CHECK(cudaMalloc(&s->rays, sizeof(Ray) * (size_t)std::max<long>(nrays, 1)));
    s->rays_cap = nrays;
  }
  s->nrays = nrays;
  if (nrays > 0) {
    s->first_nest = reinterpret_cast<const Ray *>(host_rays)[0].nest;
    CLB_CUDA_CHECK(cudaMemcpyAsync(s->rays, host_rays, sizeof(Ray) * (size_t)nrays, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
}

void clb_solver_get_rays(clb_solver *h, void *host_rays, void *stream)
{
  Solver *s = &h->s;
  if (s->nrays > 0)
    CLB_CUDA_CHECK(cudaMemcpyAsync(host_rays, s->rays, sizeof(Ray) * (size_t)s->nrays, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CLB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
}

void clb_solver_load_density(clb_solver *h, const float *counts_map, float premul, float densmul, float backdens, void *stream)
{
  Solver *s = &h->s;
  if (s->staged[s->cur].src) {   // a prefetch is (or was) filling this buffer: let it finish, then drop it
    CLB_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, s->dens_ready[s->cur], 0));
    s->staged[s->cur].src = nullptr;
  }
  load_density(s, counts_map, s->cur, premul, densmul, backdens, (cudaStream_t)stream);
}

void clb_solver_solve(clb_solver *h, const float *density_dev, void *stream)
{
  Solver *s = &h->s;
  cudaStream_t st = (cudaStream_t)stream;
  mark(s, 0, st); mark(s, 1, st);
  solve(s, density_dev ? density_dev : s->dens[s->cur], st);
}

void clb_solver_alm2allmaps(clb_solver *h, const double *alm_re, const double *alm_im, void *stream)
{
  Solver *s = &h->s;
  cudaStream_t st = (cudaStream_t)stream;
  ShtPlan *p = s->plan;
  float *mp[6];
  for (int k = 0; k < 6; ++k) mp[k] = s->maps + (size_t)k * s->npix;
  if (s->fused) stream_barrier(s, st);
  LAUNCHED(s) launch_legendre_synthesis(p, alm_re, alm_im, s->fused ? nullptr : s->b_send, st);
  if (s->fused) stream_barrier(s, st);
  LAUNCHED(s) launch_ring_synthesis(p, s->b_recv, mp, st);
  if (s->fused) {
    float *pm[kMaxPeers * 6];
    for (int q = 0; q < s->nranks; ++q)
      for (int k = 0; k < 6; ++k) pm[q * 6 + k] = reinterpret_cast<float *>(s->peer[q][2]) + (size_t)k * s->npix;
    LAUNCHED(s) launch_maps_broadcast(p, mp, pm, s->d_need, kCoarseOrder, st, s->d_gmask);
    stream_barrier(s, st);
  }
}

void clb_solver_ray_update(clb_solver *h, double wpp1, double wp, double wpm1, int mode, int with_summary, void *stream)
{
  ray_update(&h->s, wpp1, wp, wpm1, mode, with_summary != 0, (cudaStream_t)stream);
}

void clb_solver_set_next(clb_solver *h, const float *next_counts_map, float premul, float densmul, float backdens)
{
  Solver *s = &h->s;
  if (s->n_next >= 2) { s->next[0] = s->next[1]; s->n_next = 1; }   // at most two planes ahead: the oldest request goes
  s->next[s->n_next++].set(next_counts_map, premul, densmul, backdens);
}

void clb_solver_set_pair(clb_solver *h, const float *partner_counts_map, float premul, float densmul, float backdens)
{
  Solver *s = &h->s;
  if (s->shells < 2) die("clb_solver_set_pair: the solver was created for one shell per pass (clb_set_tuning(11, 2))");
  s->pair.set(partner_counts_map, premul, densmul, backdens);
}

// density buffer that holds plane (src, scalings): the staged one if the plane was prefetched, else a buffer loaded now
// (never `avoid`, the buffer of the other plane of this pass)
static int acquire_density(Solver *s, const float *src, float premul, float densmul, float backdens, int avoid, cudaStream_t st)
{
  for (int k = 0; k < 2; ++k)
    if (k != avoid && s->staged[k].is(src, premul, densmul, backdens)) {
      CLB_CUDA_CHECK(cudaStreamWaitEvent(st, s->dens_ready[k], 0));     // prefetched: nothing crosses PCIe now
      s->staged[k].src = nullptr;
      return k;
    }
  int k = -1;
  for (int c = 0; c < 2; ++c) if (c != avoid && !s->staged[c].src) { k = c; break; }   // leave a pending prefetch alone
  if (k < 0) {       // both hold other prefetched planes: the stale one is overwritten once its load has finished
    k = (avoid == 0) ? 1 : 0;
    CLB_CUDA_CHECK(cudaStreamWaitEvent(st, s->dens_ready[k], 0));
    s->staged[k].src = nullptr;
  }
  load_density(s, src, k, premul, densmul, backdens, st);
  return k;
}

static void prefetch_queued(Solver *s)
{
  for (int i = 0; i < s->n_next; ++i) {
    const Solver::PlaneKey &n = s->next[i];
    bool have = false;
    for (int k = 0; k < 2; ++k) have |= s->staged[k].is(n.src, n.scal[0], n.scal[1], n.scal[2]);
    if (have) continue;
    int k = -1;
    for (int c = 0; c < 2; ++c) if (!s->staged[c].src) { k = c; break; }
    if (k < 0) break;     // both buffers already hold planes to come
    if (s->dens_free_valid[k]) CLB_CUDA_CHECK(cudaStreamWaitEvent(s->copy_stream, s->dens_free[k], 0));
    load_density(s, reinterpret_cast<const float *>(n.src), k, n.scal[0], n.scal[1], n.scal[2], s->copy_stream);
    CLB_CUDA_CHECK(cudaEventRecord(s->dens_ready[k], s->copy_stream));
    s->staged[k] = n;
  }
  s->n_next = 0;
}

int clb_solver_check(clb_solver *h, void *stream)
{
  Solver *s = &h->s;
  int *he = reinterpret_cast<int *>(s->h_sum6 + 6);
  CLB_CUDA_CHECK(cudaMemcpyAsync(he, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CLB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return *he;
}

int clb_solver_step(clb_solver *h, const float *counts_map, float premul, float densmul, float backdens, double wpp1,
                    double wp, double wpm1, double *sum6, void *stream)
{
  Solver *s = &h->s;
  cudaStream_t st = (cudaStream_t)stream;
  mark(s, 0, st);
  if (s->cached.is(counts_map, premul, densmul, backdens)) {
    // this plane was solved together with the previous one: its six maps wait in map set 1, only the rays are left
    s->cached.src = nullptr;
    for (int k = 1; k <= 8; ++k) mark(s, k, st);
    ray_update(s, wpp1, wp, wpm1, 1 | 2 | 4, sum6 != nullptr, st, 1);
  } else {
    s->cached.src = nullptr;
    int buf[2] = {-1, -1};
    buf[0] = acquire_density(s, counts_map, premul, densmul, backdens, -1, st);
    int nshell = 1;
    if (s->pair.src) {
      buf[1] = acquire_density(s, reinterpret_cast<const float *>(s->pair.src), s->pair.scal[0], s->pair.scal[1], s->pair.scal[2], buf[0], st);
      nshell = 2;
    }
    s->cur = buf[0];
    mark(s, 1, st);
    const float *dens[2] = {s->dens[buf[0]], nshell == 2 ? s->dens[buf[1]] : nullptr};
    solve(s, dens, nshell, buf, st);
    if (nshell == 2) { s->cached = s->pair; s->pair.src = nullptr; }
    ray_update(s, wpp1, wp, wpm1, 1 | 2 | 4, sum6 != nullptr, st, 0);
  }
  prefetch_queued(s);   // the next planes' maps start streaming in behind this plane's kernels
  if (!sum6) return 0;
  CLB_CUDA_CHECK(cudaMemcpyAsync(s->h_sum6, s->d_sum6, sizeof(double) * 6, cudaMemcpyDeviceToHost, st));
  const int err = clb_solver_check(h, stream);
  memcpy(sum6, s->h_sum6, sizeof(double) * 6);
  return err;
}

void clb_solver_ray_output(clb_solver *h, void *host_out, void *stream)
{
  Solver *s = &h->s;
  if (s->nrays <= 0) return;
  Ray *tmp = nullptr;
  CLB_CUDA_CHECK(cudaMalloc(&tmp, sizeof(Ray) * (size_t)s->nrays));
  LAUNCHED(s) launch_ray_output(s->rays, tmp, s->nrays, s->ray_order, (cudaStream_t)stream);
  CLB_CUDA_CHECK(cudaMemcpyAsync(host_out, tmp, sizeof(Ray) * (size_t)s->nrays, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CLB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  CLB_CUDA_CHECK(cudaFree(tmp));
}

// map2alm_mpi / alm2allmaps_mpi on the reference's per-rank buffers (healpix_shtrans.h:67,70-72): mapvec holds this rank's
// ring pairs only (local index i = ring - firstRingTasks[rank]), alm the m range this rank owns; the transposes run
// through the fused exchange.  Host pointers; synchronises.
void clb_solver_map2alm_mapvec(clb_solver *h, const float *mapvec, const long *north_start, const long *south_start,
                               double *alm_re, double *alm_im, void *stream)
{
  Solver *s = &h->s;
  ShtPlan *p = s->plan;
  cudaStream_t st = (cudaStream_t)stream;
  const int nl = p->nrp_loc;
  const long units = mapvec_units(p);
  MapvecScratch &m = mapvec_scratch(8 * (size_t)std::max<long>(units, 1), sizeof(long) * 2 * (size_t)std::max(nl, 1));
  if (nl > 0) {
    CLB_CUDA_CHECK(cudaMemcpyAsync(m.mv, mapvec, 8 * (size_t)units, cudaMemcpyHostToDevice, st));
    CLB_CUDA_CHECK(cudaMemcpyAsync(m.idx, north_start, sizeof(long) * nl, cudaMemcpyHostToDevice, st));
    CLB_CUDA_CHECK(cudaMemcpyAsync(m.idx + nl, south_start, sizeof(long) * nl, cudaMemcpyHostToDevice, st));
    mapvec_copy_kernel<<<2 * nl, 256, 0, st>>>(m.mv, s->dens[s->cur], m.idx, m.idx + nl, p->d_rp_loc, p->d_nphi, p->d_startN,
                                               p->d_startS, 1);
    CLB_CUDA_CHECK(cudaGetLastError());
    LAUNCHED(s) 1;
  }
  if (s->fused) stream_barrier(s, st);
  LAUNCHED(s) launch_ring_analysis(p, s->dens[s->cur], s->g_send, st);
  if (s->fused) stream_barrier(s, st);
  LAUNCHED(s) launch_legendre_analysis(p, s->fused ? nullptr : s->g_recv, s->alm_re, s->alm_im, 0, st);
  if (p->alm_total > 0) {
    CLB_CUDA_CHECK(cudaMemcpyAsync(alm_re, s->alm_re, sizeof(double) * p->alm_total, cudaMemcpyDeviceToHost, st));
    CLB_CUDA_CHECK(cudaMemcpyAsync(alm_im, s->alm_im, sizeof(double) * p->alm_total, cudaMemcpyDeviceToHost, st));
  }
  CLB_CUDA_CHECK(cudaStreamSynchronize(st));
}

void clb_solver_alm2allmaps_mapvec(clb_solver *h, const double *alm_re, const double *alm_im, float *const mapvec[6],
                                   const long *north_start, const long *south_start, void *stream)
{
  Solver *s = &h->s;
  ShtPlan *p = s->plan;
  cudaStream_t st = (cudaStream_t)stream;
  const int nl = p->nrp_loc;
  const long units = mapvec_units(p);
  MapvecScratch &m = mapvec_scratch(8 * (size_t)std::max<long>(units, 1), sizeof(long) * 2 * (size_t)std::max(nl, 1));
  if (p->alm_total > 0) {
    CLB_CUDA_CHECK(cudaMemcpyAsync(s->alm_re, alm_re, sizeof(double) * p->alm_total, cudaMemcpyHostToDevice, st));
    CLB_CUDA_CHECK(cudaMemcpyAsync(s->alm_im, alm_im, sizeof(double) * p->alm_total, cudaMemcpyHostToDevice, st));
  }
  float *mp[6];
  for (int k = 0; k < 6; ++k) mp[k] = s->maps + (size_t)k * s->npix;
  if (s->fused) stream_barrier(s, st);
  LAUNCHED(s) launch_legendre_synthesis(p, s->alm_re, s->alm_im, s->fused ? nullptr : s->b_send, st);
  if (s->fused) stream_barrier(s, st);
  LAUNCHED(s) launch_ring_synthesis(p, s->b_recv, mp, st);
  if (nl > 0) {
    CLB_CUDA_CHECK(cudaMemcpyAsync(m.idx, north_start, sizeof(long) * nl, cudaMemcpyHostToDevice, st));
    CLB_CUDA_CHECK(cudaMemcpyAsync(m.idx + nl, south_start, sizeof(long) * nl, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 6; ++k) {
      // the two pad floats behind every ring are whatever the caller had there (the reference leaves FFT workspace in them)
      CLB_CUDA_CHECK(cudaMemcpyAsync(m.mv, mapvec[k], 8 * (size_t)units, cudaMemcpyHostToDevice, st));
      mapvec_copy_kernel<<<2 * nl, 256, 0, st>>>(m.mv, mp[k], m.idx, m.idx + nl, p->d_rp_loc, p->d_nphi, p->d_startN, p->d_startS, 0);
      CLB_CUDA_CHECK(cudaGetLastError());
      LAUNCHED(s) 1;
      CLB_CUDA_CHECK(cudaMemcpyAsync(mapvec[k], m.mv, 8 * (size_t)units, cudaMemcpyDeviceToHost, st));
    }
  }
  CLB_CUDA_CHECK(cudaStreamSynchronize(st));
}

}  // extern "C"
