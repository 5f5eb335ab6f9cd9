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
__global__ void __launch_bounds__(128, (R >= 5) ? 2 : (R >= 4) ? 3 : 4) syn_k(double *out, int iters, double a)
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


// analysis-like: R rings in registers, per degree one LDS.64 (broadcast) + R*(DMUL + 3 DFMA); every KB degrees an
// optional warp transpose-reduce of the 2*KB partial sums (MODE 1: shuffles, MODE 2: through shared memory)
template <int R, int KB, int MODE, int NB = 3>
__global__ void __launch_bounds__(128, NB) ana_k(double *out, int iters, double a0)
{
  constexpr int V = 2 * KB;
  __shared__ double sA[2][KB];
  __shared__ double sT[4][V][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < KB) { sA[0][threadIdx.x] = 1.0 + 1e-3 * threadIdx.x + a0; sA[1][threadIdx.x] = 1.0 - 1e-3 * threadIdx.x + a0; }
  __syncthreads();
  double mp[R], mc[R], x[R], gpx[R], gpy[R], gmx[R], gmy[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    mp[j] = 0.1 * j; mc[j] = 0.2 + threadIdx.x * 1e-4; x[j] = 0.3 + j * 0.01;
    gpx[j] = 1 + j; gpy[j] = 2 + j; gmx[j] = 3 + j; gmy[j] = 4 + j + a0;
  }
  double tot = 0;
  for (int it = 0; it < iters; ++it) {
    const double *sa = sA[it & 1];
    double v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.0;
#pragma unroll
    for (int i = 0; i < KB; ++i) {
      const double a = sa[i];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double mu = mc[j];
        if (i & 1) { v[i] = fma(mu, gmx[j], v[i]); v[KB + i] = fma(mu, gmy[j], v[KB + i]); }
        else       { v[i] = fma(mu, gpx[j], v[i]); v[KB + i] = fma(mu, gpy[j], v[KB + i]); }
        const double mn = fma(x[j] * a, mu, -mp[j]);
        mp[j] = mu; mc[j] = mn;
      }
    }
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) tot += v[i];
    } else if (MODE == 1) {
#pragma unroll
      for (int s = V / 2; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
          const double send = upper ? v[k] : v[k + s];
          const double keep = upper ? v[k + s] : v[k];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      if (V == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
      tot += v[0];
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) sT[w][i][lane] = v[i];
      __syncwarp();
      double t = 0;
      const int row = lane % V, half = (V == 16) ? (lane >> 4) : 0;
      constexpr int NC = (V == 16) ? 16 : 32;
#pragma unroll
      for (int k = 0; k < NC; ++k) t += sT[w][row][half * 16 + ((k + lane) & (NC - 1))];
      if (V == 16) t += __shfl_xor_sync(0xffffffffu, t, 16);
      tot += t;
      __syncwarp();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot + mc[0] + mc[R - 1];
}

// analysis-like with the warp reduce of block b software-pipelined into the compute of block b+1:
// stage s of the transpose-reduce runs between degrees s and s+1 of the next block, so shuffle latency hides
// behind FP64 work.  Blocks are processed in pairs so the roles of the two accumulator sets alternate without moves.
template <int R, int NB>
__global__ void __launch_bounds__(128, NB) ana_pipe_k(double *out, int iters, double a0)
{
  constexpr int KB = 8, V = 16;
  __shared__ double sA[2][KB];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < KB) { sA[0][threadIdx.x] = 1.0 + 1e-3 * threadIdx.x + a0; sA[1][threadIdx.x] = 1.0 - 1e-3 * threadIdx.x + a0; }
  __syncthreads();
  double mp[R], mc[R], x[R], gpx[R], gpy[R], gmx[R], gmy[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    mp[j] = 0.1 * j; mc[j] = 0.2 + threadIdx.x * 1e-4; x[j] = 0.3 + j * 0.01;
    gpx[j] = 1 + j; gpy[j] = 2 + j; gmx[j] = 3 + j; gmy[j] = 4 + j + a0;
  }
  double tot = 0;
  double va[V], vb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { va[i] = 0.0; vb[i] = 0.0; }
  auto compute_i = [&](double (&v)[V], const double *sa, int i) {
    const double a = sa[i];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const double mu = mc[j];
      if (i & 1) { v[i] = fma(mu, gmx[j], v[i]); v[KB + i] = fma(mu, gmy[j], v[KB + i]); }
      else       { v[i] = fma(mu, gpx[j], v[i]); v[KB + i] = fma(mu, gpy[j], v[KB + i]); }
      const double mn = fma(x[j] * a, mu, -mp[j]);
      mp[j] = mu; mc[j] = mn;
    }
  };
  auto reduce_stage = [&](double (&v)[V], int st) {   // st = 0..3: s = 8,4,2,1 ; st = 4: final xor 16 ; then consume
    if (st < 4) {
      const int s = 8 >> st;
      const bool upper = (lane & s) != 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < s) {
          const double send = upper ? v[k] : v[k + s];
          const double keep = upper ? v[k + s] : v[k];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
    } else if (st == 4) {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
    } else if (st == 5) {
      tot += v[0];
    }
  };
  for (int it = 0; it < iters; it += 2) {
    const double *sa = sA[0];
    // block A: accumulate into va while reducing vb (previous block)
#pragma unroll
    for (int i = 0; i < KB; ++i) {
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < V; ++k) va[k] = 0.0;   // folded into the first FMAs by the compiler
      }
      compute_i(va, sa, i);
      if (i < 6) reduce_stage(vb, i);
    }
    sa = sA[1];
#pragma unroll
    for (int i = 0; i < KB; ++i) {
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < V; ++k) vb[k] = 0.0;
      }
      compute_i(vb, sa, i);
      if (i < 6) reduce_stage(va, i);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot + mc[0] + mc[R - 1];
}

// FP64 tensor-core probe: mma.sync m8n8k4 f64 (DMMA) alone, and interleaved with independent DFMA chains in the same
// warps, to see whether the two share the SM's FP64 datapath (BASELINE north_star: "DMMA only if it pays").
template <int NMMA, int NFMA>
__global__ void dmma_k(double *out, int iters, double a0)
{
  double c0[NMMA > 0 ? NMMA : 1][2];
  double v[NFMA > 0 ? NFMA : 1];
  const double a = 1.0 + 1e-9 * threadIdx.x + a0, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < (NMMA > 0 ? NMMA : 1); ++i) { c0[i][0] = i; c0[i][1] = -i; }
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NMMA; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i][0]), "+d"(c0[i][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int i = 0; i < NFMA; ++i) v[i] = fma(v[i], 1.0000001, 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < (NMMA > 0 ? NMMA : 1); ++i) s += c0[i][0] + c0[i][1];
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) s += v[i];
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
  {
    float ms;
    ms = timeit([&] { syn_k<5><<<sms * 2, 128>>>(out, it2, 1e-9); });
    printf("synthesis-like R=5, 8 warps/SM: %6.2f\n", 16.0 * 5 * 16 * it2 * sms * 2 * 128 / ms * 1e-9);
    ms = timeit([&] { syn_k<6><<<sms * 2, 128>>>(out, it2, 1e-9); });
    printf("synthesis-like R=6, 8 warps/SM: %6.2f\n", 16.0 * 6 * 16 * it2 * sms * 2 * 128 / ms * 1e-9);
    ms = timeit([&] { syn_k<4><<<sms * 2, 128>>>(out, it2, 1e-9); });
    printf("synthesis-like R=4, 8 warps/SM: %6.2f\n", 16.0 * 4 * 16 * it2 * sms * 2 * 128 / ms * 1e-9);
  }
  printf("analysis-like loop at 12 warps/SM (8 flop per ring-degree): TFLOP/s, modes none / shuffle reduce / smem reduce\n");
  {
    float ms; const int it3 = 1 << 12; const int ctas = sms * 3;
#define ANA(R, KB) \
    printf("R=%d KB=%2d:", R, KB); \
    ms = timeit([&] { ana_k<R, KB, 0><<<ctas, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * R * KB * it3 * ctas * 128 / ms * 1e-9); \
    ms = timeit([&] { ana_k<R, KB, 1><<<ctas, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * R * KB * it3 * ctas * 128 / ms * 1e-9); \
    ms = timeit([&] { ana_k<R, KB, 2><<<ctas, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * R * KB * it3 * ctas * 128 / ms * 1e-9);
    ANA(8, 8) ANA(6, 16)
    {
      const int ctas2 = sms * 2;
      printf("2 CTAs/SM R=12 KB=8:");
      ms = timeit([&] { ana_k<12, 8, 0, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 12 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<12, 8, 1, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 12 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<12, 8, 2, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 12 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      printf("2 CTAs/SM R=10 KB=16:");
      ms = timeit([&] { ana_k<10, 16, 0, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 10 * 16 * it3 * ctas2 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<10, 16, 1, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 10 * 16 * it3 * ctas2 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<10, 16, 2, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 10 * 16 * it3 * ctas2 * 128 / ms * 1e-9);
      printf("2 CTAs/SM R=8 KB=8:");
      ms = timeit([&] { ana_k<8, 8, 0, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 8 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<8, 8, 1, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 8 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<8, 8, 2, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 8 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      printf("pipelined reduce R=8 KB=8 (3 CTAs/SM):");
      ms = timeit([&] { ana_pipe_k<8, 3><<<ctas, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 8 * 8 * it3 * ctas * 128 / ms * 1e-9);
      printf("pipelined reduce R=6 KB=8 (3 CTAs/SM):");
      ms = timeit([&] { ana_pipe_k<6, 3><<<ctas, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 6 * 8 * it3 * ctas * 128 / ms * 1e-9);
      printf("pipelined reduce R=8 KB=8 (4 CTAs/SM):");
      ms = timeit([&] { ana_pipe_k<8, 4><<<sms * 4, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 8 * 8 * it3 * sms * 4 * 128 / ms * 1e-9);
      printf("pipelined reduce R=12 KB=8 (2 CTAs/SM):");
      ms = timeit([&] { ana_pipe_k<12, 2><<<ctas2, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 12 * 8 * it3 * ctas2 * 128 / ms * 1e-9);
      const int ctas4 = sms * 4;
      printf("4 CTAs/SM R=8 KB=8 (128 regs):");
      ms = timeit([&] { ana_k<8, 8, 0, 4><<<ctas4, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 8 * 8 * it3 * ctas4 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<8, 8, 1, 4><<<ctas4, 128>>>(out, it3, 1e-9); }); printf(" %6.2f", 8.0 * 8 * 8 * it3 * ctas4 * 128 / ms * 1e-9);
      ms = timeit([&] { ana_k<8, 8, 2, 4><<<ctas4, 128>>>(out, it3, 1e-9); }); printf(" %6.2f\n", 8.0 * 8 * 8 * it3 * ctas4 * 128 / ms * 1e-9);
    }
  }
  {
    printf("FP64 tensor (mma.m8n8k4.f64) vs FMA pipe, 16 warps/SM; TFLOP/s counted as 2*8*8*4 per warp MMA and 2*32 per warp FMA\n");
    const int itd = 1 << 14; float ms;
    const double wm = (double)sms * 4 * 4;   // warps in flight: 4 CTAs/SM x 4 warps
    ms = timeit([&] { dmma_k<8, 0><<<sms * 4, 128>>>(out, itd, 1e-9); });
    printf("DMMA only (8 independent accumulators): %6.2f\n", 8.0 * 512 * itd * wm / ms * 1e-9);
    ms = timeit([&] { dmma_k<0, 8><<<sms * 4, 128>>>(out, itd, 1e-9); });
    printf("DFMA only (8 chains):                   %6.2f\n", 8.0 * 64 * itd * wm / ms * 1e-9);
    ms = timeit([&] { dmma_k<8, 8><<<sms * 4, 128>>>(out, itd, 1e-9); });
    printf("DMMA + DFMA interleaved (8 + 8):        %6.2f total (%6.2f DMMA + %6.2f DFMA)\n",
           (8.0 * 512 + 8.0 * 64) * itd * wm / ms * 1e-9, 8.0 * 512 * itd * wm / ms * 1e-9, 8.0 * 64 * itd * wm / ms * 1e-9);
    ms = timeit([&] { dmma_k<2, 16><<<sms * 4, 128>>>(out, itd, 1e-9); });
    printf("DMMA + DFMA interleaved (2 + 16):       %6.2f total (%6.2f DMMA + %6.2f DFMA)\n",
           (2.0 * 512 + 16.0 * 64) * itd * wm / ms * 1e-9, 2.0 * 512 * itd * wm / ms * 1e-9, 16.0 * 64 * itd * wm / ms * 1e-9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
