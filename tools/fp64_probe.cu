// FP64 pipe probe: DFMA throughput versus resident warps per SM and independent chains per thread, with and without
// broadcast LDS.128 operand traffic (the Legendre synthesis pattern: 8*R DFMA per two LDS.128).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_k(double *out, int iters, double a, double b)
{
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// synthesis-like: R rings, per degree 2 LDS.128 (broadcast) + R*(6 acc FMA + DMUL + DFMA)
template <int R>
__global__ void __launch_bounds__(128, (R >= 4) ? 3 : 4) syn_k(double *out, int iters, double a)
{
  __shared__ __align__(16) double tiles[2][16 * 8];
  tiles[0][threadIdx.x] = 1e-3 * (threadIdx.x % 7) + a;
  tiles[1][threadIdx.x] = 2e-3 * (threadIdx.x % 5) + a;
  __syncthreads();
  double acc[R][12], mp[R], mc[R], x[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    mp[j] = 0.1 * j; mc[j] = 0.2 + threadIdx.x * 1e-4; x[j] = 0.3 + j * 0.01;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[j][k] = 0;
  }
  for (int it = 0; it < iters; ++it) {
    const double *tile = tiles[it & 1];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double4 ra = *reinterpret_cast<const double4 *>(&tile[i * 8]);
      const double4 rb = *reinterpret_cast<const double4 *>(&tile[i * 8 + 4]);
      const int par = (i & 1) * 6;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double mu = mc[j];
        acc[j][par + 0] = fma(mu, ra.y, acc[j][par + 0]);
        acc[j][par + 1] = fma(mu, ra.z, acc[j][par + 1]);
        acc[j][par + 2] = fma(mu, ra.w, acc[j][par + 2]);
        acc[j][par + 3] = fma(mu, rb.x, acc[j][par + 3]);
        acc[j][par + 4] = fma(mu, rb.y, acc[j][par + 4]);
        acc[j][par + 5] = fma(mu, rb.z, acc[j][par + 5]);
        const double mn = fma(x[j] * ra.x, mu, -mp[j]);
        mp[j] = mu; mc[j] = mn;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int k = 0; k < 12; ++k) s += acc[j][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  return ms;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *out; cudaMalloc(&out, sizeof(double) * sms * 2048);
  const int iters = 1 << 14;
  printf("plain DFMA chains: TFLOP/s by warps/SM (rows) and ILP (cols 2,4,8,16)\n");
  for (int warps = 4; warps <= 32; warps += (warps < 16 ? 4 : 16)) {
    printf("warps/SM %2d:", warps);
    float ms;
    ms = timeit([&] { dfma_k<2><<<sms, warps * 32>>>(out, iters * 8, 1.0000001, 1e-9); });
    printf(" %6.2f", 2.0 * 2 * iters * 8 * sms * warps * 32 / ms * 1e-9);
    ms = timeit([&] { dfma_k<4><<<sms, warps * 32>>>(out, iters * 4, 1.0000001, 1e-9); });
    printf(" %6.2f", 2.0 * 4 * iters * 4 * sms * warps * 32 / ms * 1e-9);
    ms = timeit([&] { dfma_k<8><<<sms, warps * 32>>>(out, iters * 2, 1.0000001, 1e-9); });
    printf(" %6.2f", 2.0 * 8 * iters * 2 * sms * warps * 32 / ms * 1e-9);
    ms = timeit([&] { dfma_k<16><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); });
    printf(" %6.2f\n", 2.0 * 16 * iters * sms * warps * 32 / ms * 1e-9);
  }
  printf("synthesis-like loop (16 flop per ring-degree executed): TFLOP/s by warps/SM (rows) and R (cols 1,2,3,4)\n");
  const int it2 = 1 << 10;
  for (int warps = 4; warps <= 16; warps += 4) {
    printf("warps/SM %2d:", warps);
    float ms;
    ms = timeit([&] { syn_k<1><<<sms * warps / 4, 128>>>(out, it2 * 4, 1e-9); });
    printf(" %6.2f", 16.0 * 1 * 16 * it2 * 4 * sms * warps * 32 / ms * 1e-9);
    ms = timeit([&] { syn_k<2><<<sms * warps / 4, 128>>>(out, it2 * 2, 1e-9); });
    printf(" %6.2f", 16.0 * 2 * 16 * it2 * 2 * sms * warps * 32 / ms * 1e-9);
    if (warps <= 16) {
      ms = timeit([&] { syn_k<3><<<sms * warps / 4, 128>>>(out, it2, 1e-9); });
      printf(" %6.2f", 16.0 * 3 * 16 * it2 * sms * warps * 32 / ms * 1e-9);
    }
    if (warps <= 12) {
      ms = timeit([&] { syn_k<4><<<sms * warps / 4, 128>>>(out, it2, 1e-9); });
      printf(" %6.2f", 16.0 * 4 * 16 * it2 * sms * warps * 32 / ms * 1e-9);
    }
    printf("\n");
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
