// calclens_b200/csrc/rays.cu   (compiled with -fmad=false, see raymath.cuh)
// Per-ray lens-plane update on the GPU: one thread per ray.
//   step 1  interpolate phi, grad phi, grad grad phi at the ray position from the six derivative maps,
//           parallel-transporting each pixel's vector/tensor to the ray       [shtpoissonsolve.c:666-702, :1122-1204]
//   step 2  deflect, advance to the next shell, A-matrix recursion, transport  [rayprop.c:18-189]
// Rays stay in the reference's 176-byte HEALPixRay layout (raytrace.h:284-293); a CTA stages its rays through
// shared memory so that HBM sees fully coalesced 16-byte accesses on the AoS array.
#include "sht_internal.cuh"
#include "raymath.cuh"
#include <algorithm>
#include <mutex>

namespace clb {

constexpr int kRayThreads = 128;

static_assert(sizeof(Ray) % 16 == 0, "ray struct must be a multiple of 16 bytes");

struct RayMaps { const float *p[6]; };

// per-ring table of the fast interpolation path (raymath.cuh), one per (device, map order), built on first use
__global__ void ring_table_kernel(RingTab *__restrict__ tab, long order)
{
  const long ring = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nside = 1L << order;
  if (ring < 1 || ring > 4 * nside - 1) return;
  const RingInfo ri = ring_info(ring, order);
  RingTab t;
  t.theta = atan2(ri.sintheta, ri.costheta);                  // get_interpol's theta1/theta2
  const double th = acos(ri.costheta);                         // nest2ang returns acos(z), ang2vec takes cos of it again
  t.cz = cos(th);
  t.sz = sqrt((1.0 + t.cz) * (1.0 - t.cz));
  t.inv_sz = 1.0 / t.sz;
  const double dphi = CLB_PI_2 / (double)(ri.ringpix / 4);
  t.cd = cos(dphi); t.sd = sin(dphi);
  t.dphi = 2.0 * CLB_PI / ri.ringpix;
  t.inv_dphi = 1.0 / t.dphi;
  t.inv_ringpix = 1.0 / (double)ri.ringpix;
  t.inv_dtheta = 0.0;
  if (ring < 4 * nside - 1) {
    const RingInfo rn = ring_info(ring + 1, order);
    t.inv_dtheta = 1.0 / (atan2(rn.sintheta, rn.costheta) - t.theta);
  }
  tab[ring] = t;
}

// (built once per (device, order) and kept for the life of the process; the build is synchronised, so any stream may
// use the table afterwards)
static const RingTab *ring_table(long order, cudaStream_t st)
{
  static RingTab *cache[16][32] = {};
  static std::mutex mu;
  int dev = 0;
  CLB_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || order < 0 || order >= 32) { fprintf(stderr, "calclens_b200: ring_table(%d, %ld)\n", dev, order); abort(); }
  std::lock_guard<std::mutex> lock(mu);
  if (!cache[dev][order]) {
    const long n = 4L << order;
    RingTab *t = nullptr;
    CLB_CUDA_CHECK(cudaMalloc(&t, sizeof(RingTab) * n));
    ring_table_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(t, order);
    CLB_CUDA_CHECK(cudaGetLastError());
    CLB_CUDA_CHECK(cudaStreamSynchronize(st));
    cache[dev][order] = t;
  }
  return cache[dev][order];
}

// ---- TMA (bulk async copy) helpers: one thread moves a whole tile of ray records between HBM and shared memory ----
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// mode bit 0: zero phi/alpha/U first (the driver's pre-solve reset, raytrace.c:213-230)
// mode bit 1: interpolate + accumulate;  mode bit 2: propagate;  mode bit 3 (with bit 2): Born-approximation propagate
// Persistent CTAs walk the ray array in tiles of kRayThreads records (22.5 KB, contiguous).  Tiles move by TMA bulk
// copies issued by one thread: the next tile streams into the second shared-memory buffer (completion on an mbarrier)
// while the current one is being computed in place, and a finished tile leaves as one bulk store.
__global__ void __launch_bounds__(kRayThreads, 4)
ray_step_kernel(Ray *__restrict__ rays, long nrays, RayMaps maps, const RingTab *__restrict__ tab, long order, double wp,
                double wpm1, double wpm2, int mode, const unsigned char *__restrict__ need,
                const unsigned char *__restrict__ safe, int coarse_shift, unsigned rank_bit, int *__restrict__ err,
                double *__restrict__ sum6, PlaneCoef pc)
{
  __shared__ __align__(128) unsigned char s_raw[2][kRayThreads * sizeof(Ray)];
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ double s_sum[6][kRayThreads / 32];   // per-warp running sums of the plane summary (sum6 != nullptr)
  if (sum6 && (threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) s_sum[k][threadIdx.x >> 5] = 0.0;
  }
  const long ntiles = (nrays + kRayThreads - 1) / kRayThreads;
  long tile = blockIdx.x;
  if (tile >= ntiles) return;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto tile_bytes = [&](long t) { return (unsigned)(min((long)kRayThreads, nrays - t * kRayThreads) * (long)sizeof(Ray)); };
  if (threadIdx.x == 0) {
    mbar_expect_tx(&s_bar[0], tile_bytes(tile));
    tma_load_1d(s_raw[0], rays + tile * kRayThreads, tile_bytes(tile), &s_bar[0]);
  }
  int buf = 0;
  unsigned phase[2] = {0u, 0u};
  for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const long next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) {
      // the other buffer was handed to a bulk store one iteration ago: its shared-memory reads must be over
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_expect_tx(&s_bar[buf ^ 1], tile_bytes(next));
      tma_load_1d(s_raw[buf ^ 1], rays + next * kRayThreads, tile_bytes(next), &s_bar[buf ^ 1]);
    }
    mbar_wait(&s_bar[buf], phase[buf]);
    phase[buf] ^= 1u;
    const long first = tile * kRayThreads;
    const int nblk = (int)min((long)kRayThreads, nrays - first);
    Ray *s_rays = reinterpret_cast<Ray *>(s_raw[buf]);
    double ps[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (threadIdx.x < nblk) {
      Ray &ray = s_rays[threadIdx.x];   // operate on the staged record in place: fields are read where they are used
      if (mode & 1) {
        ray.phi = 0.0; ray.alpha[0] = 0.0; ray.alpha[1] = 0.0;
        ray.U[0] = 0.0; ray.U[1] = 0.0; ray.U[2] = 0.0; ray.U[3] = 0.0;
      }
      if (mode & 2) {
        long px[4];
        ray_interp_accumulate_fast(ray, order, tab, maps.p[0], maps.p[1], maps.p[2], maps.p[3], maps.p[4], maps.p[5], px);
        // sharded runs: every pixel of the stencil must lie inside the part of the sky this rank received (the reference
        // aborts on a missing map cell, shtpoissonsolve.c:683-689; here the flag is raised and the host aborts)
        // safe[c] (optional) says that cell c AND every cell around it were delivered, which covers any stencil that
        // starts in c; only rays in the rim of the received region pay for the per-pixel check
        if (need) {
          const long c0 = ring2nest(px[0], order) >> coarse_shift;
          if (!safe || !(safe[c0] & rank_bit)) {
            unsigned ok = rank_bit & need[c0];
#pragma unroll
            for (int k = 1; k < 4; ++k) ok &= need[ring2nest(px[k], order) >> coarse_shift];
            if (!ok) atomicOr(err, 1);
          }
        }
      }
      if (mode & 4) {
        if (mode & 8) ray_propagate_born(ray, wp, wpm1, wpm2);
        else ray_propagate_fast(ray, wp, wpm1, pc);
      }
      if (sum6) {   // same six sums as ray_summary_kernel, without a second pass over the ray array
        ps[0] = 1.0 - 0.5 * (ray.A[0] + ray.A[3]);
        ps[1] = 0.5 * (ray.A[3] - ray.A[0]);
        ps[2] = -0.5 * (ray.A[1] + ray.A[2]);
        ps[3] = ray.alpha[0] * ray.alpha[0] + ray.alpha[1] * ray.alpha[1];
        ps[4] = ray.phi;
        ps[5] = 0.5 * (ray.A[2] - ray.A[1]);
      }
    }
    if (sum6) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double v = ps[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_sum[k][threadIdx.x >> 5] += v;
      }
    }
    // the records were written through the generic proxy: make them visible to the async (TMA) proxy, then one
    // thread hands the tile to a bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) tma_store_1d(rays + first, s_raw[buf], (unsigned)(nblk * (int)sizeof(Ray)));
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before exit
  if (sum6 && (threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(&sum6[k], s_sum[k][threadIdx.x >> 5]);
  }
}

int launch_ray_step(Ray *d_rays, long nrays, const float *const d_maps[6], long order, double wp, double wpm1,
                    double wpm2, int mode, cudaStream_t st, const unsigned char *d_need, long coarse_order, int rank, int *d_err,
                    double *d_sum6, const unsigned char *d_safe)
{
  if (nrays <= 0) return 0;
  RayMaps m;
  for (int k = 0; k < 6; ++k) m.p[k] = d_maps ? d_maps[k] : nullptr;
  const long ntiles = (nrays + kRayThreads - 1) / kRayThreads;
  const long nblocks = std::min<long>(ntiles, (long)sm_count() * 4);
  const RingTab *tab = (mode & 2) ? ring_table(order, st) : nullptr;
  if (d_sum6) CLB_CUDA_CHECK(cudaMemsetAsync(d_sum6, 0, 6 * sizeof(double), st));
  if (d_need && (coarse_order > order || !d_err)) d_need = nullptr;
  if (!d_need) d_safe = nullptr;
  // plane constants of the A recursion with the reference's expressions (rayprop.c:134-139), host double arithmetic
  PlaneCoef pc;
  pc.ccur = wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2);
  pc.cprev = 1.0 - wpm1 * (wp - wpm2) / wp / (wpm1 - wpm2);
  pc.cu = (wp - wpm1) / wp;
  ray_step_kernel<<<(unsigned)nblocks, kRayThreads, 0, st>>>(d_rays, nrays, m, tab, order, wp, wpm1, wpm2, mode, d_need, d_safe,
                                                             (int)(2 * (order - coarse_order)), 1u << rank, d_err, d_sum6, pc);
  if (d_sum6) CLB_CUDA_CHECK(cudaGetLastError());
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// Ray initialisation: ray i observes from the centre of NEST pixel first_nest+i at rayOrder, sits at radius binL/2,
// A = Aprev = identity                                                  [raytrace_utils.c:302-347 init_rays]
__global__ void ray_init_kernel(Ray *__restrict__ rays, long nrays, long first_nest, long ray_order, double binL_2)
{
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrays) return;
  Ray ray;
  ray.nest = first_nest + i;
  nest2vec(ray.nest, ray_order, ray.beta);
  ray.n[0] = ray.beta[0] * binL_2; ray.n[1] = ray.beta[1] * binL_2; ray.n[2] = ray.beta[2] * binL_2;
  ray.A[0] = 1.0; ray.A[1] = 0.0; ray.A[2] = 0.0; ray.A[3] = 1.0;
  ray.Aprev[0] = 1.0; ray.Aprev[1] = 0.0; ray.Aprev[2] = 0.0; ray.Aprev[3] = 1.0;
  ray.phi = 0.0; ray.alpha[0] = 0.0; ray.alpha[1] = 0.0;
  ray.U[0] = 0.0; ray.U[1] = 0.0; ray.U[2] = 0.0; ray.U[3] = 0.0;
  rays[i] = ray;
}
int launch_ray_init(Ray *d_rays, long nrays, long first_nest, long ray_order, double binL_2, cudaStream_t st)
{
  if (nrays <= 0) return 0;
  ray_init_kernel<<<(unsigned)((nrays + 255) / 256), 256, 0, st>>>(d_rays, nrays, first_nest, ray_order, binL_2);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// Per-step scalar summary read back by the host each plane: sums over rays of
// kappa = 1 - (A00+A11)/2, gamma1 = (A11-A00)/2, gamma2 = -(A01+A10)/2, |alpha|^2, phi, omega = (A10-A01)/2
__global__ void ray_summary_kernel(const Ray *__restrict__ rays, long nrays, double *__restrict__ out)
{
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nrays; i += (long)gridDim.x * blockDim.x) {
    const Ray &r = rays[i];
    s[0] += 1.0 - 0.5 * (r.A[0] + r.A[3]);
    s[1] += 0.5 * (r.A[3] - r.A[0]);
    s[2] += -0.5 * (r.A[1] + r.A[2]);
    s[3] += r.alpha[0] * r.alpha[0] + r.alpha[1] * r.alpha[1];
    s[4] += r.phi;
    s[5] += 0.5 * (r.A[2] - r.A[1]);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[k], s[k]);
  }
}
int launch_ray_summary(const Ray *d_rays, long nrays, double *d_out6, cudaStream_t st)
{
  CLB_CUDA_CHECK(cudaMemsetAsync(d_out6, 0, 6 * sizeof(double), st));
  if (nrays <= 0) return 0;
  ray_summary_kernel<<<sm_count() * 4, 256, 0, st>>>(d_rays, nrays, d_out6);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// write_rays' pre-output transform (rayio.c:300-312) into a separate output array (the device-resident rays keep
// their current-position basis): A, Aprev parallel transported from the ray's position to the centre of the pixel it
// is observed in (paratrans_ray_curr2obs, rot_paratrans.c:274-302), then alpha, A, Aprev, U rotated from the
// (theta, phi) to the (ra, dec) basis (rot_ray_ang2radec, rot_paratrans.c:375-411).
__global__ void ray_output_kernel(const Ray *__restrict__ rays, Ray *__restrict__ out, long nrays, long ray_order)
{
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrays) return;
  Ray r = rays[i];
  double obs[3], c, s, T[2][2], RT[2][2];
  nest2vec(r.nest, ray_order, obs);
  paratrans_angle(r.n, obs, c, s);   // the reference evaluates the same angle once per tensor
  T[0][0] = r.Aprev[0]; T[0][1] = r.Aprev[1]; T[1][0] = r.Aprev[2]; T[1][1] = r.Aprev[3];
  transport_tensor(T, c, s, RT);
  r.Aprev[0] = RT[0][0]; r.Aprev[1] = RT[0][1]; r.Aprev[2] = RT[1][0]; r.Aprev[3] = RT[1][1];
  T[0][0] = r.A[0]; T[0][1] = r.A[1]; T[1][0] = r.A[2]; T[1][1] = r.A[3];
  transport_tensor(T, c, s, RT);
  r.A[0] = RT[0][0]; r.A[1] = RT[0][1]; r.A[2] = RT[1][0]; r.A[3] = RT[1][1];
  const double a0 = r.alpha[0], a1 = r.alpha[1];
  r.alpha[0] = a1; r.alpha[1] = -1.0 * a0;
  double *M[3] = {r.A, r.Aprev, r.U};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double m00 = M[k][0], m01 = M[k][1], m10 = M[k][2], m11 = M[k][3];
    M[k][0] = m11; M[k][2] = -1.0 * m01; M[k][1] = -1.0 * m10; M[k][3] = m00;
  }
  out[i] = r;
}
int launch_ray_output(const Ray *d_rays, Ray *d_out, long nrays, long ray_order, cudaStream_t st)
{
  if (nrays <= 0) return 0;
  ray_output_kernel<<<(unsigned)((nrays + 127) / 128), 128, 0, st>>>(d_rays, d_out, nrays, ray_order);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// NGP particle deposit (shtpoissonsolve.c:128-150, NGPSHTDENS): val += (float)(mass/MASS_SCALE) in the pixel
// ang2nest(vec2ang(pos), poissonOrder), written to a RING-ordered map (the reference deposits into NEST-ordered local
// cells and shuffles them to rings, map_shuffle.c:633).  The pixel index is bit-exact; float atomics make the sum
// independent of the particle order exactly when all particles of a pixel carry the same mass (the usual case: every
// partial sum k*m is then formed by the same sequence of additions), otherwise it agrees to float round-off.
__global__ void deposit_ngp_kernel(const float *__restrict__ pos, const float *__restrict__ mass, long nparts, long order,
                                   float *__restrict__ ringmap)
{
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nparts; k += (long)gridDim.x * blockDim.x) {
    double vec[3] = {(double)pos[3 * k], (double)pos[3 * k + 1], (double)pos[3 * k + 2]}, theta, phi;
    vec2ang(vec, theta, phi);
    const long pix = nest2ring(ang2nest(theta, phi, order), order);
    atomicAdd(&ringmap[pix], (float)(mass[k] / 1e10));   // MASS_SCALE
  }
}
int launch_deposit_ngp(const float *d_pos, const float *d_mass, long nparts, long order, float *d_ringmap, cudaStream_t st)
{
  if (nparts <= 0) return 0;
  const long blocks = std::min<long>((nparts + 255) / 256, (long)sm_count() * 16);
  deposit_ngp_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_pos, d_mass, nparts, order, d_ringmap);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// ---- small utility kernels used by tests: device versions of the indexing functions ----
__global__ void healpix_index_kernel(int what, long order, long n, const long *__restrict__ in, const double *__restrict__ th,
                                     const double *__restrict__ ph, long *__restrict__ out)
{
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (what) {
    case 0: out[i] = ring2nest(in[i], order); break;
    case 1: out[i] = nest2ring(in[i], order); break;
    case 2: out[i] = ang2nest(th[i], ph[i], order); break;
    case 3: out[i] = nest2peano(in[i], order); break;
    default: out[i] = -1;
  }
}
__global__ void healpix_interpol_kernel(long order, long n, const double *__restrict__ vec, long *__restrict__ pix, double *__restrict__ wgt)
{
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[3] = {vec[3 * i], vec[3 * i + 1], vec[3 * i + 2]}, theta, phi, w[4];
  long p[4];
  vec2ang(v, theta, phi);
  get_interpol(theta, phi, p, w, order);
  for (int k = 0; k < 4; ++k) { pix[4 * i + k] = p[k]; wgt[4 * i + k] = w[k]; }
}

void launch_healpix_index(int what, long order, long n, const long *in, const double *th, const double *ph, long *out, cudaStream_t st)
{
  if (n <= 0) return;
  healpix_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(what, order, n, in, th, ph, out);
  CLB_CUDA_CHECK(cudaGetLastError());
}
// the stencil exactly as ray_step_kernel forms it: vec2ang + get_interpol_tab (tabulated ring colatitudes, reciprocal
// weights) -- exposed so the bit-exactness tests exercise the hot kernel's own index path
__global__ void ray_stencil_kernel(long order, long n, const double *__restrict__ vec, const RingTab *__restrict__ tab,
                                   long *__restrict__ pix, double *__restrict__ wgt)
{
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[3] = {vec[3 * i], vec[3 * i + 1], vec[3 * i + 2]}, theta, phi, w[4];
  long p[4], ra, rb;
  vec2ang(v, theta, phi);
  get_interpol_tab(theta, phi, p, w, order, tab, ra, rb);
  for (int k = 0; k < 4; ++k) { pix[4 * i + k] = p[k]; wgt[4 * i + k] = w[k]; }
}

void launch_healpix_interpol(long order, long n, const double *vec, long *pix, double *wgt, int use_table, cudaStream_t st)
{
  if (n <= 0) return;
  if (use_table) ray_stencil_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order, n, vec, ring_table(order, st), pix, wgt);
  else healpix_interpol_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order, n, vec, pix, wgt);
  CLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace clb
