// calclens_b200/csrc/rays.cu   (compiled with -fmad=false, see raymath.cuh)
// Per-ray lens-plane update on the GPU: one thread per ray.
//   step 1  interpolate phi, grad phi, grad grad phi at the ray position from the six derivative maps,
//           parallel-transporting each pixel's vector/tensor to the ray       [shtpoissonsolve.c:666-702, :1122-1204]
//   step 2  deflect, advance to the next shell, A-matrix recursion, transport  [rayprop.c:18-189]
// Rays stay in the reference's 176-byte HEALPixRay layout (raytrace.h:284-293); a CTA stages its rays through
// shared memory so that HBM sees fully coalesced 16-byte accesses on the AoS array.
#include "sht_internal.cuh"
#include "raymath.cuh"

namespace clb {

constexpr int kRayThreads = 128;
constexpr int kRayWords = sizeof(Ray) / 16;   // 11 x 16 bytes per ray
static_assert(sizeof(Ray) % 16 == 0, "ray struct must be a multiple of 16 bytes");

struct RayMaps { const float *p[6]; };

// mode bit 0: zero phi/alpha/U first (the driver's pre-solve reset, raytrace.c:213-230)
// mode bit 1: interpolate + accumulate;  mode bit 2: propagate
__global__ void __launch_bounds__(kRayThreads)
ray_step_kernel(Ray *__restrict__ rays, long nrays, RayMaps maps, long order, double wp, double wpm1, double wpm2, int mode)
{
  __shared__ __align__(16) unsigned char s_raw[kRayThreads * sizeof(Ray)];
  Ray *s_rays = reinterpret_cast<Ray *>(s_raw);
  const long first = (long)blockIdx.x * kRayThreads;
  const int nblk = (int)min((long)kRayThreads, nrays - first);
  const int4 *src = reinterpret_cast<const int4 *>(rays + first);
  int4 *dst = reinterpret_cast<int4 *>(s_raw);
  for (int i = threadIdx.x; i < nblk * kRayWords; i += kRayThreads) dst[i] = src[i];
  __syncthreads();
  if (threadIdx.x < nblk) {
    Ray ray = s_rays[threadIdx.x];
    if (mode & 1) {
      ray.phi = 0.0; ray.alpha[0] = 0.0; ray.alpha[1] = 0.0;
      ray.U[0] = 0.0; ray.U[1] = 0.0; ray.U[2] = 0.0; ray.U[3] = 0.0;
    }
    if (mode & 2) ray_interp_accumulate(ray, order, maps.p[0], maps.p[1], maps.p[2], maps.p[3], maps.p[4], maps.p[5]);
    if (mode & 4) ray_propagate(ray, wp, wpm1, wpm2);
    s_rays[threadIdx.x] = ray;
  }
  __syncthreads();
  int4 *gdst = reinterpret_cast<int4 *>(rays + first);
  for (int i = threadIdx.x; i < nblk * kRayWords; i += kRayThreads) gdst[i] = dst[i];
}

int launch_ray_step(Ray *d_rays, long nrays, const float *const d_maps[6], long order, double wp, double wpm1,
                    double wpm2, int mode, cudaStream_t st)
{
  if (nrays <= 0) return 0;
  RayMaps m;
  for (int k = 0; k < 6; ++k) m.p[k] = d_maps ? d_maps[k] : nullptr;
  const long nblocks = (nrays + kRayThreads - 1) / kRayThreads;
  ray_step_kernel<<<(unsigned)nblocks, kRayThreads, 0, st>>>(d_rays, nrays, m, order, wp, wpm1, wpm2, mode);
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
void launch_healpix_interpol(long order, long n, const double *vec, long *pix, double *wgt, cudaStream_t st)
{
  if (n <= 0) return;
  healpix_interpol_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order, n, vec, pix, wgt);
  CLB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace clb
