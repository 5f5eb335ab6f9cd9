// Micro-benchmark: FP64 FMA throughput, 64-bit shuffle throughput, HBM copy bandwidth on the
// current GPU.  Used once to establish the FP64 roofline denominator (MEASURED_PEAKS.json only has
// HBM and bf16).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
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

__global__ void shfl_kernel(double *out, int iters)
{
  double v = threadIdx.x, w = threadIdx.x * 2.0;
  for (int it = 0; it < iters; ++it) {
    v = __shfl_xor_sync(0xffffffffu, v, 1);
    w = __shfl_xor_sync(0xffffffffu, w, 2);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + w;
}

__global__ void copy_kernel(const double4 *__restrict__ in, double4 *__restrict__ out, size_t n)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs %d clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 16 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 1 << 16;
  for (int threads = 128; threads <= 1024; threads *= 2) {
    int blocks = p.multiProcessorCount * (2048 / threads > 4 ? 4 : 2048 / threads);
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    double flops = 2.0 * 8 * iters * (double)blocks * threads;
    printf("dfma ILP8 blocks %d threads %d: %.3f ms  %.2f TFLOP/s\n", blocks, threads, ms, flops / ms * 1e-9);
  }
  {
    int threads = 256, blocks = p.multiProcessorCount * 4;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      dfma_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    double flops = 2.0 * 4 * iters * (double)blocks * threads;
    printf("dfma ILP4 blocks %d threads %d: %.3f ms  %.2f TFLOP/s\n", blocks, threads, ms, flops / ms * 1e-9);
    // long sustained run (about 2 s) to see the power-capped rate
    CK(cudaEventRecord(e0));
    for (int k = 0; k < 40; ++k) dfma_kernel<8><<<blocks, threads>>>(out, iters * 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    flops = 40.0 * 2.0 * 8 * iters * 4 * (double)blocks * threads;
    printf("dfma sustained: %.1f ms  %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      shfl_kernel<<<blocks, threads>>>(out, iters);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    double sh = 2.0 * iters * (double)blocks * threads / 32;  // 64-bit warp shuffles
    printf("shfl64: %.3f ms  %.3f G warp-shfl64/s  (%.2f per clk per SM at %d kHz)\n", ms, sh / ms * 1e-6,
           sh / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate);
  }
  {
    size_t n = (size_t)1 << 27;  // 128 Mi double4 = 4 GiB
    double4 *a, *b; CK(cudaMalloc(&a, n * 32)); CK(cudaMalloc(&b, n * 32));
    CK(cudaMemset(a, 1, n * 32));
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      copy_kernel<<<p.multiProcessorCount * 16, 512>>>(a, b, n);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      printf("copy 4 GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * n * 32 / ms * 1e-6);
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
