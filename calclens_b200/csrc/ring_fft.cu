// calclens_b200/csrc/ring_fft.cu
// Batched HEALPix ring FFTs for sm_100a: one CTA per ring (and field), everything between the HBM read and the HBM
// write stays in shared memory.  Replaces the per-ring FFTW calls and the pack/unpack loops of the reference:
//   analysis : ring weights, r2c, alias pick + conjugate, half-pixel phase, write g_m
//              [map2alm_transpose_mpi.c:151-191 (weights + ring_analysis), :219-315 (pack); healpix_shtrans.c:549-571]
//   synthesis: alias fold of b_m into float half-complex bins, half-pixel phase, c2r, 1/sin(theta) scalings,
//              cot(theta) cross terms
//              [alm2allmaps_transpose_mpi.c:818-947 (unpack), :1034-1089 (FFT + scale), :1097-1147 (cot terms);
//               healpix_shtrans.c:168-205 ring_synthesis]
//
// Precision contract (SURVEY.md App. A.6 / DESIGN.md): the reference stores ring spectra as float and the FFT
// library it links is third-party, so the oracle uses an exactly-rounded float FFT.  Here the transform itself
// runs in FP64 and is rounded to float exactly where the reference stores a float; every other float rounding
// point of the reference is reproduced explicitly with __double2float_rn-style casts and un-contracted
// arithmetic (__dmul_rn/__dadd_rn) so that results agree to the last float bit except for FP64-level noise.
//
// Algorithm: a real ring has n = 4r pixels.  Radix-4 split into four real length-r sequences, packed pairwise
// into two complex length-r DFTs.  DFT_r: r a power of two -> in-place DIF (bit-reversed read-out), both transforms
// of a ring batched in the same passes; otherwise Bluestein with M = pow2 >= 2r-1: DIF forward, product with a
// precomputed bit-reversed chirp spectrum, DIT inverse (no bit reversal anywhere).  Passes are radix 16 in registers
// over XOR-swizzled shared memory; half-pixel phase factors come from plan-time tables.  Rings too long for shared
// memory (r > 4095) run the same code from an L2-resident global scratch slice.  (Index math prototyped in
// tools/fft_proto.py.)
#include "sht_internal.cuh"
#include "healpix.cuh"
#include <algorithm>
#include <math.h>

namespace clb {

constexpr int kFoldTile = 1024;   // degrees per shared-memory tile of the streamed alias fold (classes of short rings)

struct FftClass {   // one launch group: ring pairs that share a shared-memory footprint
  int logM;          // largest work length in the group, Mmax = 1 << logM
  int bluestein;     // group key only (groups with logM <= kSmallLogM mix both paths)
  int rmax;          // largest r in the group (sizes shared memory)
  int tail;          // float2 entries behind bufB (even)
  int tileT;         // degrees per tile of the streamed alias fold, stored behind the tail (0: class of long rings)
  int count;         // ring pairs in the class (local ones only)
  int *d_rp = nullptr;   // [count] global ring-pair indices
  int threads;
  size_t smem_ana, smem_syn;
};

struct FftTables {
  int logTW = 0;                 // twiddle table: exp(-2 pi i k / TW), k < TW/2
  double2 *d_tw = nullptr;
  // Bluestein tables, indexed by r (only non-power-of-two r < Nside are filled)
  long *d_chirp_off = nullptr;   // [nside+1] offset of chirp_r (r entries)
  long *d_bhat_off = nullptr;    // [nside+1] offset of bhat_r (M(r) entries, bit-reversed order, includes 1/M)
  double2 *d_chirp = nullptr;
  double2 *d_bhat = nullptr;
  // half-pixel phase factors exp(+i pi k / n), k = 0 .. n/2, one run per polar ring pair and one shared by all
  // equatorial ones; formed with exactly the expression the reference uses per call (healpix_shtrans.c:186-197)
  double2 *d_phase = nullptr;
  long *d_phase_off = nullptr;   // [nrp]
  signed char *d_rp_logM = nullptr;   // [nrp] work length per ring pair
  signed char *d_rp_blu = nullptr;    // [nrp] 1: Bluestein
  std::vector<FftClass> classes;
  // rings whose work buffers exceed an SM's shared memory (r > 4095) run from a global scratch buffer
  double2 *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  // the class launches of a stage are independent: they go round-robin to the caller's stream and three side streams, so that
  // the CTAs of the next class fill the SMs the tail of the previous one leaves idle (fft_fork / fft_join)
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  int depth = 0, rr = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b)   // a * conj(b)
{
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }   // a * (+i)

// Work buffers of the CTA FFTs are stored XOR-swizzled: logical element i lives at i ^ ((i >> 4) & 15).  The last
// radix-16 pass has every thread touch 16 consecutive elements (thread stride 256 bytes); unswizzled, the eight
// threads of a quarter warp would hit the same four banks (8-way conflict on every 16-byte access).  The swizzle
// permutes elements inside aligned runs of 16, so passes whose threads walk consecutive elements stay conflict free.
// The power-of-two path also reads its result in bit-reversed order (consecutive threads differ in the TOP bits of
// the index), so the top three bits are folded into the swizzle as well (only bits >= 4 feed the XOR: a bijection).
__device__ __forceinline__ int swz(int i, int logM)
{
  int x = (i >> 4) & 15;
  if (logM >= 7) x ^= (i >> (logM - 3)) & 7;
  return i ^ x;
}

// exp(i * pi * num / den) for integers, exact argument reduction
__device__ __forceinline__ double2 unit_pi(long num, long den)
{
  long twoden = 2 * den;
  num %= twoden;
  if (num < 0) num += twoden;
  double s, c;
  sincospi((double)num / (double)den, &s, &c);
  return make_double2(c, s);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-cooperative in-place FFTs on shared memory (all threads of the block must call)
// ---------------------------------------------------------------------------------------------------------------
// Both directions run K = 4 radix-2 stages per pass in registers (radix 16): a thread gathers the 16 elements of one
// butterfly group, applies the four stages with twiddles W^j, W^2j, W^4j, W^8j formed from ONE table look-up by
// squaring (times compile-time 16th roots of unity), and scatters them back -- a quarter of the shared-memory
// passes and an eighth of the twiddle look-ups of a radix-2 sweep.  A pass over fewer stages (K = 1..3) takes up the
// remainder of log2(M).  Any split into consecutive radix-2 stages leaves the same bit-reversed ordering.

// exp(-2 pi i q / 2^n), q < 2^(n-1), n <= 4
__device__ __forceinline__ double2 root16(int n, int q)
{
  constexpr double c8 = 0.92387953251128675613, s8 = 0.38268343236508977173, h = 0.70710678118654752440;
  const int k = q << (4 - n);   // index into the 16th roots, k < 8
  switch (k) {
    case 0: return make_double2(1.0, 0.0);
    case 1: return make_double2(c8, -s8);
    case 2: return make_double2(h, -h);
    case 3: return make_double2(s8, -c8);
    case 4: return make_double2(0.0, -1.0);
    case 5: return make_double2(-s8, -c8);
    case 6: return make_double2(-h, -h);
    default: return make_double2(-c8, -s8);
  }
}

// Twiddles.  A pass over stages lg .. lg-K+1 needs W_L^j = exp(-2 pi i j / 2^lg) for j < 2^(lg-K), i.e. W_M^i with
// i = j << (logM - lg) < M / 2^K.  The passes are ordered so that every pass with non-trivial twiddles is a radix-16 one
// (i < M/16), and W_M^i is formed as hi[i >> 5] * lo[i & 31] from two 32-entry tables in shared memory (hi[a] = W_M^(32a),
// lo[b] = W_M^b) -- no global twiddle loads inside the passes, whose L2 latency was the largest single stall of the
// Bluestein rings (profiles/r02_ring_fft.txt).  The tables alias the block-reduction scratch (s_red), which is idle
// while the transforms run; cta_twiddle_tables() rebuilds them after every block reduction.
constexpr int kTwEntries = 64;   // hi[32] | lo[32]
__device__ __forceinline__ void cta_twiddle_tables(double2 *stw, int logM, const double2 *__restrict__ tw, int logTW)
{
  // (the caller guarantees a barrier between the last use of the aliased scratch and this call)
  for (int t = threadIdx.x; t < kTwEntries; t += blockDim.x) {   // (small classes run CTAs of 32 threads)
    const long i = (t < 32) ? ((long)t << 5) : (long)(t - 32);            // exponent of W_M
    double2 w = make_double2(1.0, 0.0);
    if (logM >= 1 && i < (1L << (logM - 1))) w = __ldg(&tw[i << (logTW - logM)]);   // table holds exp(-2 pi i k / TW), k < TW/2
    stw[t] = w;                                                            // (exponents >= M/2 are never requested)
  }
  __syncthreads();
}
__device__ __forceinline__ double2 tw_get(const double2 *stw, int i)     // W_M^i, i < M/16 <= 1024
{
  return cmul(stw[i >> 5], stw[32 + (i & 31)]);
}

// forward (sign -1) pass over the K stages with block lengths 2^lg, 2^(lg-1), .., 2^(lg-K+1)
template <int K>
__device__ __forceinline__ void cta_dif_pass(double2 *a, int M, int lg, const double2 *stw, double2 *a2 = nullptr)
{
  constexpr int RR = 1 << K;
  const int s = 1 << (lg - K);
  const int logM = 31 - __clz(M);
  const int sh = logM - lg;
  const int nbf = M >> K;                      // butterflies per array; a2 (optional) is a second, independent array
  for (int it = threadIdx.x; it < (a2 ? 2 * nbf : nbf); it += blockDim.x) {
    double2 *arr = (it >= nbf) ? a2 : a;
    const int idx = (it >= nbf) ? it - nbf : it;
    const int j = idx & (s - 1);
    const int base = ((idx >> (lg - K)) << lg) + j;
    double2 x[RR];
#pragma unroll
    for (int q = 0; q < RR; ++q) x[q] = arr[swz(base + q * s, logM)];
    double2 w = (s > 1) ? tw_get(stw, j << sh) : make_double2(1.0, 0.0);   // W_L^j
#pragma unroll
    for (int t = 0; t < K; ++t) {
      const int half = RR >> (t + 1);
#pragma unroll
      for (int g = 0; g < RR; g += 2 * half) {
#pragma unroll
        for (int q = 0; q < half; ++q) {
          const double2 u = x[g + q], v = x[g + q + half];
          x[g + q] = cadd(u, v);
          const double2 d = csub(u, v);
          x[g + q + half] = cmul(d, q == 0 ? w : cmul(w, root16(K - t, q)));
        }
      }
      w = cmul(w, w);
    }
#pragma unroll
    for (int q = 0; q < RR; ++q) arr[swz(base + q * s, logM)] = x[q];
  }
  __syncthreads();
}

// inverse (sign +1) pass: the same K stages in reverse order, conjugate twiddles applied before the butterflies
template <int K>
__device__ __forceinline__ void cta_dit_pass(double2 *a, int M, int lg, const double2 *stw,
                                             const double2 *__restrict__ premul = nullptr)
{
  constexpr int RR = 1 << K;
  const int s = 1 << (lg - K);
  const int logM = 31 - __clz(M);
  const int sh = logM - lg;
  for (int idx = threadIdx.x; idx < (M >> K); idx += blockDim.x) {
    const int j = idx & (s - 1);
    const int base = ((idx >> (lg - K)) << lg) + j;
    double2 x[RR];
#pragma unroll
    for (int q = 0; q < RR; ++q) x[q] = a[swz(base + q * s, logM)];
    if (premul) {   // pointwise product with a table in the same (bit-reversed) order, fused into the first pass
#pragma unroll
      for (int q = 0; q < RR; ++q) x[q] = cmul(x[q], __ldg(&premul[base + q * s]));
    }
    double2 wp[K];                             // W_L^(j 2^t)
    wp[0] = (s > 1) ? tw_get(stw, j << sh) : make_double2(1.0, 0.0);
#pragma unroll
    for (int t = 1; t < K; ++t) wp[t] = cmul(wp[t - 1], wp[t - 1]);
#pragma unroll
    for (int t = K - 1; t >= 0; --t) {
      const int half = RR >> (t + 1);
#pragma unroll
      for (int g = 0; g < RR; g += 2 * half) {
#pragma unroll
        for (int q = 0; q < half; ++q) {
          const double2 u = x[g + q];
          const double2 v = cmulc(x[g + q + half], q == 0 ? wp[t] : cmul(wp[t], root16(K - t, q)));
          x[g + q] = cadd(u, v);
          x[g + q + half] = csub(u, v);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RR; ++q) a[swz(base + q * s, logM)] = x[q];
  }
  __syncthreads();
}

// forward, sign -1, natural order in -> bit-reversed order out: radix-16 passes from the top stage down, the remaining
// logM mod 4 stages last (block lengths <= 8: unit twiddle W_L^0 only)
// a2 (optional): a second array of the same length transformed in the same passes
__device__ void cta_fft_dif(double2 *a, int logM, const double2 *stw, double2 *a2 = nullptr)
{
  const int M = 1 << logM;
  int lg = logM;
  for (; lg >= 4; lg -= 4) cta_dif_pass<4>(a, M, lg, stw, a2);
  switch (lg) {
    case 1: cta_dif_pass<1>(a, M, 1, stw, a2); break;
    case 2: cta_dif_pass<2>(a, M, 2, stw, a2); break;
    case 3: cta_dif_pass<3>(a, M, 3, stw, a2); break;
    default: break;
  }
}

// inverse (unnormalised), sign +1, bit-reversed order in -> natural order out (the mirror image of cta_fft_dif)
// premul (optional): the spectrum is multiplied by this table (same order) on its way into the first pass; used when the
// first pass is the short remainder one, whose threads read consecutive elements (coalesced table reads)
__device__ void cta_fft_dit_inv(double2 *a, int logM, const double2 *stw, const double2 *__restrict__ premul = nullptr)
{
  const int M = 1 << logM;
  const int rem = logM & 3;
  switch (rem) {
    case 1: cta_dit_pass<1>(a, M, 1, stw, premul); premul = nullptr; break;
    case 2: cta_dit_pass<2>(a, M, 2, stw, premul); premul = nullptr; break;
    case 3: cta_dit_pass<3>(a, M, 3, stw, premul); premul = nullptr; break;
    default: break;
  }
  for (int lg = rem + 4; lg <= logM; lg += 4) { cta_dit_pass<4>(a, M, lg, stw, premul); premul = nullptr; }
}

// block-wide sum of NV doubles per thread; result valid in every thread (red: shared scratch of NV*32 doubles)
template <int NV>
__device__ void cta_sum(double (&v)[NV], double *red)
{
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    for (int o = 16; o; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + w] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = 0.0;
    for (int k = 0; k < nw; ++k) t += red[i * 32 + k];
    v[i] = t;
  }
  __syncthreads();
}

__device__ __forceinline__ int bitrev(int v, int bits) { return bits ? (int)(__brev((unsigned)v) >> (32 - bits)) : 0; }

// Forward DFT of length r held in a[0..r) (natural order).  Bluestein path: caller has ALREADY multiplied by the
// chirp and zero-filled a[r..M).  On return element k of the spectrum is dft_get(a, k, ...).
// final_chirp = false leaves the last chirp product of the Bluestein path to the caller (dft_get_chirp).
__device__ void cta_dft_r(double2 *a, int r, int logM, int bluestein, const double2 *__restrict__ chirp,
                          const double2 *__restrict__ bhat, const double2 *stw, bool final_chirp = true)
{
  cta_fft_dif(a, logM, stw);
  if (!bluestein) return;
  const int M = 1 << logM;
  if ((logM & 3) != 0) {
    // the inverse starts with the short remainder pass (consecutive elements per thread): the product with the chirp
    // spectrum rides on it
    cta_fft_dit_inv(a, logM, stw, bhat);
  } else {
    // (a radix-16 first pass reads 16 consecutive elements per thread, which would turn the coalesced table read into 32
    // wavefronts per load: keep the separate product pass)
#pragma unroll 4
    for (int k = threadIdx.x; k < M; k += blockDim.x) a[swz(k, logM)] = cmul(a[swz(k, logM)], __ldg(&bhat[k]));
    __syncthreads();
    cta_fft_dit_inv(a, logM, stw);
  }
  if (!final_chirp) return;
#pragma unroll 4
  for (int k = threadIdx.x; k < r; k += blockDim.x) a[swz(k, logM)] = cmul(a[swz(k, logM)], __ldg(&chirp[k]));
  __syncthreads();
}
__device__ __forceinline__ double2 dft_get_chirp(const double2 *a, int k, int logM, int bluestein,
                                                 const double2 *__restrict__ chirp)
{
  return bluestein ? cmul(a[swz(k, logM)], __ldg(&chirp[k])) : a[swz(bitrev(k, logM), logM)];
}
__device__ __forceinline__ double2 dft_get(const double2 *a, int k, int logM, int bluestein)
{
  return bluestein ? a[swz(k, logM)] : a[swz(bitrev(k, logM), logM)];
}

// ---------------------------------------------------------------------------------------------------------------
// plan-time: Bluestein chirp spectra
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) bluestein_table_kernel(const int *__restrict__ rlist, int nr, const long *__restrict__ chirp_off,
                                       const long *__restrict__ bhat_off, double2 *__restrict__ chirp_all,
                                       double2 *__restrict__ bhat_all, const double2 *__restrict__ tw, int logTW,
                                       double2 *scratch, long stride)
{
  extern __shared__ double2 smem_dyn[];
  __shared__ double2 s_tw[kTwEntries];
  double2 *smem = scratch ? scratch + (long)blockIdx.x * stride : smem_dyn;   // large rings: global scratch
  for (int item = blockIdx.x; item < nr; item += gridDim.x) {
    const int r = rlist[item];
    int logM = 0;
    while ((1 << logM) < 2 * r - 1) ++logM;
    const int M = 1 << logM;
    double2 *chirp = chirp_all + chirp_off[r];
    double2 *bhat = bhat_all + bhat_off[r];
    for (int k = threadIdx.x; k < M; k += blockDim.x) smem[k] = make_double2(0.0, 0.0);   // (all of it: order irrelevant)
    __syncthreads();
    for (int j = threadIdx.x; j < r; j += blockDim.x) {
      long j2 = ((long)j * j) % (2L * r);
      double2 w = unit_pi(-j2, r);       // exp(-i pi j^2 / r)
      chirp[j] = w;
      double2 c = cconj(w);
      smem[swz(j, logM)] = c;
      if (j) smem[swz(M - j, logM)] = c;
    }
    __syncthreads();
    cta_twiddle_tables(s_tw, logM, tw, logTW);
    cta_fft_dif(smem, logM, s_tw);
    const double inv = 1.0 / (double)M;   // fold the inverse-FFT normalisation into the table (exact power of two)
    for (int k = threadIdx.x; k < M; k += blockDim.x) bhat[k] = make_double2(smem[swz(k, logM)].x * inv, smem[swz(k, logM)].y * inv);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// plan-time: phase tables
// ---------------------------------------------------------------------------------------------------------------
__global__ void phase_table_kernel(const int *__restrict__ nphi, const long *__restrict__ phase_off, int nrp, int nside,
                                   double2 *__restrict__ phase)
{
  const int rp = blockIdx.x;
  const int n = nphi[rp];
  if (n == 4 * nside && rp != nside - 1) return;   // equatorial rings share the run of the first one
  double2 *t = phase + phase_off[rp];
  for (int k = threadIdx.x; k <= n / 2; k += blockDim.x) {
    const double ang = __ddiv_rn(__dmul_rn((double)k, CLB_PI), (double)n);
    t[k] = make_double2(cos(ang), sin(ang));
  }
}
// exp(-i pi x / n) for any integer x >= 0 from the table T[k] = exp(+i pi k / n), k <= n/2 (exact symmetries)
__device__ __forceinline__ double2 phase_neg(const double2 *__restrict__ T, long x, int n)
{
  x %= 2L * n;
  if (2 * x <= n) return cconj(__ldg(&T[x]));
  if (x <= n) { const double2 t = __ldg(&T[n - x]); return make_double2(-t.x, -t.y); }
  if (2 * x <= 3L * n) { const double2 t = __ldg(&T[x - n]); return make_double2(-t.x, t.y); }
  return __ldg(&T[2L * n - x]);
}

// ---------------------------------------------------------------------------------------------------------------
// analysis kernel: one CTA per (ring pair in class, hemisphere)
// ---------------------------------------------------------------------------------------------------------------
struct RingGeomDev {
  const double *cth, *sth, *weight;
  const int *nphi, *shifted;
  const long *startN, *startS;
};

struct AnaArgs {
  const float *map; double2 *g_send; RingGeomDev geo;
  const int *class_rp, *rp_to_local; const long *m_goff; int lmax;
  const signed char *rp_logM, *rp_blu; int Mmax;
  const long *chirp_off, *bhat_off; const double2 *chirp_all, *bhat_all, *tw; int logTW;
  const double2 *phase_all; const long *phase_off;
};

// work = 2 * (index into the class's ring-pair list) + hemisphere; smem = the CTA's work buffers (shared memory, or a
// private slice of a global scratch buffer for rings whose buffers exceed an SM's shared memory)
__device__ __forceinline__ void ring_analysis_body(const AnaArgs &A, double2 *smem, int work)
{
  const float *__restrict__ map = A.map; double2 *__restrict__ g_send = A.g_send; const RingGeomDev &geo = A.geo;
  const int *__restrict__ class_rp = A.class_rp, *__restrict__ rp_to_local = A.rp_to_local;
  const long *__restrict__ m_goff = A.m_goff; const int lmax = A.lmax;
  const signed char *__restrict__ rp_logM = A.rp_logM, *__restrict__ rp_blu = A.rp_blu; const int Mmax = A.Mmax;
  const long *__restrict__ chirp_off = A.chirp_off, *__restrict__ bhat_off = A.bhat_off;
  const double2 *__restrict__ chirp_all = A.chirp_all, *__restrict__ bhat_all = A.bhat_all, *__restrict__ tw = A.tw;
  const int logTW = A.logTW; const double2 *__restrict__ phase_all = A.phase_all;
  const long *__restrict__ phase_off = A.phase_off;
  const int rp = class_rp[work >> 1];
  const int hemi = work & 1;
  const int n = geo.nphi[rp];
  const int r = n >> 2;
  const int logM = rp_logM[rp], bluestein = rp_blu[rp];
  const int M = 1 << logM;
  const double2 *PT = phase_all + phase_off[rp];
  const long start = hemi ? geo.startS[rp] : geo.startN[rp];
  const int slot = 2 * rp_to_local[rp] + hemi;
  if (start < 0) {   // equator has no southern partner: its slot carries zeros
    for (int m = threadIdx.x; m <= lmax; m += blockDim.x) g_send[m_goff[m] + slot] = make_double2(0.0, 0.0);
    return;
  }
  double2 *bufA = smem;          // [M]
  double2 *bufB = smem + Mmax;   // [r]  first spectrum, natural order
  const double w = geo.weight[rp];
  const double2 *chirp = bluestein ? chirp_all + chirp_off[r] : nullptr;
  const double2 *bhat = bluestein ? bhat_all + bhat_off[r] : nullptr;
  const float4 *ring4 = reinterpret_cast<const float4 *>(map + start);   // start is a multiple of 4 for every ring

  // Bins whose twiddles are rational (k = 0, n/4, n/2; real parts of n/6, n/3) are exact sums of floats and sit on
  // float rounding ties with probability ~1/n, where FFT round-off would flip a coin.  They are formed from the
  // exact class sums T[c] = sum_{j = c mod 12} x_j instead (the oracle's exactly-rounded FFT does the same).
  __shared__ __align__(16) double s_red[12 * 32];
  double T[12];
  {
    double t3[3][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) t3[a][q] = 0.0;
    for (int j = threadIdx.x; j < r; j += blockDim.x) {
      float4 x = __ldg(&ring4[j]);
      const double xw[4] = {(double)__double2float_rn(__dmul_rn((double)x.x, w)), (double)__double2float_rn(__dmul_rn((double)x.y, w)),
                            (double)__double2float_rn(__dmul_rn((double)x.z, w)), (double)__double2float_rn(__dmul_rn((double)x.w, w))};
      const int a = j % 3;   // pixel 4j+q is in class (4j+q) mod 12 = 4*(j mod 3) + q
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (a == 0) t3[0][q] += xw[q];
        else if (a == 1) t3[1][q] += xw[q];
        else t3[2][q] += xw[q];
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) T[4 * a + q] = t3[a][q];
    cta_sum<12>(T, s_red);
  }
  double2 *stw = reinterpret_cast<double2 *>(s_red);   // the reduction scratch is idle from here on: twiddle tables
  cta_twiddle_tables(stw, logM, tw, logTW);

  for (int pass = 0; pass < 2; ++pass) {
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      double2 z = make_double2(0.0, 0.0);
      if (j < r) {
        float4 x = __ldg(&ring4[j]);
        // map = (float)(map * w_r)                                  [map2alm_transpose_mpi.c:161-162]
        float xa = __double2float_rn(__dmul_rn((double)(pass ? x.z : x.x), w));
        float xb = __double2float_rn(__dmul_rn((double)(pass ? x.w : x.y), w));
        z = make_double2((double)xa, (double)xb);
        if (bluestein) z = cmul(z, __ldg(&chirp[j]));
      }
      bufA[swz(j, logM)] = z;
    }
    __syncthreads();
    cta_dft_r(bufA, r, logM, bluestein, chirp, bhat, stw, pass != 0);
    if (pass == 0) {
      for (int k = threadIdx.x; k < r; k += blockDim.x) bufB[k] = dft_get_chirp(bufA, k, logM, bluestein, chirp);
      __syncthreads();
    }
  }
  // bufB = Z1 (natural), bufA = Z2 (natural or bit-reversed).  F_k = X0 + W X1 + W^2 X2 + W^3 X3, W = exp(-2 pi i k/n)
  const int shifted = geo.shifted[rp];
  for (int m = threadIdx.x; m <= lmax; m += blockDim.x) {
    int mind = m % n;                                               // [map2alm_transpose_mpi.c:237-252]
    bool conj_it = false;
    if (mind > n / 2) { mind = n - mind; conj_it = true; }
    const int kk = mind % r;
    const int kc = (r - kk) % r;
    double2 z1 = bufB[kk], z1c = cconj(bufB[kc]);
    double2 z2 = dft_get(bufA, kk, logM, bluestein), z2c = cconj(dft_get(bufA, kc, logM, bluestein));
    double2 X0 = make_double2(0.5 * (z1.x + z1c.x), 0.5 * (z1.y + z1c.y));
    double2 d1 = csub(z1, z1c);
    double2 X1 = make_double2(0.5 * d1.y, -0.5 * d1.x);              // -i/2 * (Z - conj Z')
    double2 X2 = make_double2(0.5 * (z2.x + z2c.x), 0.5 * (z2.y + z2c.y));
    double2 d2 = csub(z2, z2c);
    double2 X3 = make_double2(0.5 * d2.y, -0.5 * d2.x);
    double2 W1 = phase_neg(PT, 2L * mind, n);
    double2 W2 = cmul(W1, W1);
    double2 W3 = cmul(W1, W2);
    double2 F = cadd(cadd(X0, cmul(W1, X1)), cadd(cmul(W2, X2), cmul(W3, X3)));
    if (mind == 0) {
      F.x = ((T[0] + T[1]) + (T[2] + T[3])) + ((T[4] + T[5]) + (T[6] + T[7])) + ((T[8] + T[9]) + (T[10] + T[11]));
      F.y = 0.0;
    } else if (2 * mind == n) {
      F.x = ((T[0] + T[2]) + (T[4] + T[6]) + (T[8] + T[10])) - ((T[1] + T[3]) + (T[5] + T[7]) + (T[9] + T[11]));
      F.y = 0.0;
    } else if (4 * mind == n) {
      F.x = (T[0] + T[4] + T[8]) - (T[2] + T[6] + T[10]);
      F.y = (T[3] + T[7] + T[11]) - (T[1] + T[5] + T[9]);
    } else if (6 * mind == n) {      // cos(pi j/3) = 1, 1/2, -1/2, -1, -1/2, 1/2
      F.x = ((T[0] + T[6]) - (T[3] + T[9])) + 0.5 * (((T[1] + T[7]) + (T[5] + T[11])) - ((T[2] + T[8]) + (T[4] + T[10])));
    } else if (3 * mind == n) {      // cos(2 pi j/3) = 1, -1/2, -1/2
      F.x = ((T[0] + T[3]) + (T[6] + T[9])) - 0.5 * (((T[1] + T[4]) + (T[7] + T[10])) + ((T[2] + T[5]) + (T[8] + T[11])));
    }
    // the reference keeps the spectrum as float                     [healpix_shtrans.c:549-571]
    double gr = (double)__double2float_rn(F.x);
    double gi = (double)__double2float_rn(F.y);
    if (conj_it) gi = -gi;
    if (shifted) {                                                  // [map2alm_transpose_mpi.c:255-271]
      const double2 ph = phase_neg(PT, m, n);                       // exp(-i m pi / n)
      double p0 = ph.x, p1 = ph.y;
      double t0 = __dsub_rn(__dmul_rn(gr, p0), __dmul_rn(gi, p1));
      double t1 = __dadd_rn(__dmul_rn(gr, p1), __dmul_rn(gi, p0));
      gr = t0; gi = t1;
    }
    g_send[m_goff[m] + slot] = make_double2(gr, gi);
  }
}

__global__ void __launch_bounds__(512) ring_analysis_kernel(AnaArgs A)
{
  extern __shared__ double2 smem[];
  ring_analysis_body(A, smem, blockIdx.x);
}
// persistent variant: work buffers in global scratch (L2 resident), a CTA walks several rings
__global__ void __launch_bounds__(512) ring_analysis_scratch_kernel(AnaArgs A, double2 *scratch, long stride, int nwork)
{
  double2 *buf = scratch + (long)blockIdx.x * stride;
  for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
    ring_analysis_body(A, buf, work);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// synthesis kernel: one CTA per (ring pair in class, hemisphere, field)
// ---------------------------------------------------------------------------------------------------------------
// float half-complex bin k of one ring, accumulated exactly like the reference's unpack loop
// (ascending m; positive-m contribution before the negative-m one; one float rounding per contribution).
// b of (m, row = local ring pair * 6 + field, hemisphere): blocks are ordered [ring pair][field][m][hemisphere]
struct BAddr {
  const double2 *b; const long *off; const int *str; long row; int hemi;
  __device__ __forceinline__ double2 operator()(long m) const { return __ldg(&b[off[m] + row * str[m] + hemi]); }
};
__device__ __forceinline__ float2 fold_bin(const BAddr &B, int k, int n, int lmax, int shifted)
{
  // The contributions to bin k come from m = k, n-k, n+k, 2n-k, 2n+k, ... (k > 0) or m = 0, n, n, 2n, 2n, ...
  // (k = 0): term t >= 0 has m_t = (t+1)/2 * n + (t odd ? -k : +k) for k > 0; positive-m terms (t even) add b,
  // negative-m terms (t odd) add conj(b).  Sign: skfact = -1 when the ring is shifted and the wrap count is odd
  // (wrap count = j for the positive term of round j, j+1 for the negative one).
  float re = 0.f, im = 0.f;
  // common case (every ring with more than lmax pixels, except near its Nyquist bin): only m = k lands in this bin
  if (n - k > lmax && (k > 0 || n > lmax)) {
    if (k > lmax) return make_float2(0.f, 0.f);
    const double2 b = B(k);
    return make_float2(__double2float_rn(__dadd_rn(0.0, b.x)), __double2float_rn(__dadd_rn(0.0, b.y)));
  }
  auto term_m = [&](int t) -> long {
    if (k == 0) return (long)((t + 1) >> 1) * n;
    return (long)((t + 1) >> 1) * n + ((t & 1) ? -k : k);
  };
  auto apply = [&](int t, double2 b) {
    const int wraps = (t + 1) >> 1;                                 // [alm2allmaps_transpose_mpi.c:836-881]
    const double sk = (shifted && (wraps & 1)) ? -1.0 : 1.0;
    re = __double2float_rn(__dadd_rn((double)re, __dmul_rn(b.x, sk)));
    if (t & 1) im = __double2float_rn(__dsub_rn((double)im, __dmul_rn(b.y, sk)));
    else im = __double2float_rn(__dadd_rn((double)im, __dmul_rn(b.y, sk)));
  };
  // k = 0 visits m = 0 once (positive term only), every later multiple of n twice (positive, then negative)
  int t = 0;
  if (k == 0) {
    apply(0, B(0));
    for (long j = 1; j * n <= lmax; j += 4) {
      double2 b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long m = (j + u) * n;
        b[u] = (m <= lmax) ? B(m) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if ((j + u) * n > lmax) break;
        // positive term of round j (wraps = j), then the negative term -j*n (wraps = j as well: l = (-m - 0)/n = -j)
        const double sk = (shifted && ((j + u) & 1)) ? -1.0 : 1.0;
        re = __double2float_rn(__dadd_rn((double)re, __dmul_rn(b[u].x, sk)));
        im = __double2float_rn(__dadd_rn((double)im, __dmul_rn(b[u].y, sk)));
        const double sk2 = (shifted && ((j + u - 1 + 1) & 1)) ? -1.0 : 1.0;
        re = __double2float_rn(__dadd_rn((double)re, __dmul_rn(b[u].x, sk2)));
        im = __double2float_rn(__dsub_rn((double)im, __dmul_rn(b[u].y, sk2)));
      }
    }
    return make_float2(re, im);
  }
  for (;; t += 8) {
    if (term_m(t) > lmax) break;
    double2 b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long m = term_m(t + u);
      b[u] = (m <= lmax) ? B(m) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (term_m(t + u) > lmax) return make_float2(re, im);   // terms are visited in ascending m: the first miss ends the bin
      apply(t + u, b[u]);
    }
  }
  return make_float2(re, im);
}

struct MapPtrs { float *p[6]; };

// Shared memory: bufA[M] | bufB[r+1] | tail.  tail = float2 park[r] (Bluestein) or float2 Y[2r+1] (power of two,
// reused as park).  On the Bluestein path the bins Y overlay bufA[M/2..M), which is unused until the zero fill.
struct SynArgs {
  const double2 *b_recv; MapPtrs maps; RingGeomDev geo;
  const int *class_rp, *rp_to_local; const long *m_boff; const int *m_bstr; int lmax;
  const signed char *rp_logM, *rp_blu; int Mmax, rmax;
  const long *chirp_off, *bhat_off; const double2 *chirp_all, *bhat_all, *tw; int logTW;
  const double2 *phase_all; const long *phase_off;
  int dbg;            // development aid (clb_set_tuning(8, bits)): 1 no b loads, 2 no transforms, 4 no stores -- wrong results, phase costs
  int nfg;             // field groups per ring (1: all six in one CTA; 3: {0,3},{1,5},{2,4})
  int tail, tileT;     // float2 entries of the tail buffer; degrees per tile of the streamed alias fold behind it (0: none)
};




// One CTA synthesises two (or all six) fields of one ring, one after the other: work = 2 * (index into the class list) +
// hemisphere, group = which fields.  Field 1 is always done before 5 and 2 before 4 by the same threads, so the
// cot(theta) cross terms (alm2allmaps_transpose_mpi.c:1097-1147) are applied on the way out of fields 4 and 5 -- no second
// kernel, no second pass over the maps.  The b_m of a (ring, field) are one contiguous run in m (hemispheres interleaved):
// streaming reads; the CTA of the other hemisphere, next in line, finds the other half of every sector in L2.
__device__ __forceinline__ void ring_synthesis_body(const SynArgs &A, double2 *smem, int work, int group)
{
  const double2 *__restrict__ b_recv = A.b_recv; const MapPtrs &maps = A.maps; const RingGeomDev &geo = A.geo;
  const int *__restrict__ class_rp = A.class_rp, *__restrict__ rp_to_local = A.rp_to_local;
  const int lmax = A.lmax;
  const signed char *__restrict__ rp_logM = A.rp_logM, *__restrict__ rp_blu = A.rp_blu;
  const int Mmax = A.Mmax, rmax = A.rmax;
  const long *__restrict__ chirp_off = A.chirp_off, *__restrict__ bhat_off = A.bhat_off;
  const double2 *__restrict__ chirp_all = A.chirp_all, *__restrict__ bhat_all = A.bhat_all, *__restrict__ tw = A.tw;
  const int logTW = A.logTW; const double2 *__restrict__ phase_all = A.phase_all;
  const long *__restrict__ phase_off = A.phase_off;
  const int rp = class_rp[work >> 1];
  const int hemi = work & 1;
  const long start = hemi ? geo.startS[rp] : geo.startN[rp];
  if (start < 0) return;
  const int n = geo.nphi[rp];
  const int r = n >> 2;
  const int logM = rp_logM[rp], bluestein = rp_blu[rp];
  const int M = 1 << logM;
  const double2 *PT = phase_all + phase_off[rp];
  const int shifted = geo.shifted[rp];
  const long rpl = rp_to_local[rp];
  double2 *bufA = smem;
  double2 *bufB = smem + Mmax;
  float2 *tailbuf = reinterpret_cast<float2 *>(smem + Mmax + rmax + 1);
  float2 *Y = bluestein ? reinterpret_cast<float2 *>(smem + (M >> 1)) : tailbuf;
  float2 *park = tailbuf;
  const double2 *chirp = bluestein ? chirp_all + chirp_off[r] : nullptr;
  const double2 *bhat = bluestein ? bhat_all + bhat_off[r] : nullptr;
  const double sth = geo.sth[rp];
  const double cot = __ddiv_rn(geo.cth[rp], sth);   // the north ring's cos(theta); signs flip in the south
  // x / sin(theta) for every pixel of three fields: the reciprocal is rounded once per ring and every quotient gets one
  // FMA correction step (q = x rs; q += (x - q s) rs), which yields the correctly rounded double quotient (Markstein),
  // i.e. the same float as the reference's (float)((double)x / sin(theta)), at 3 instructions instead of a division
  const double rsth = __ddiv_rn(1.0, sth);
  auto div_sth = [&](float x) {
    const double xd = (double)x, q = __dmul_rn(xd, rsth);
    return __double2float_rn(__fma_rn(__fma_rn(-q, sth, xd), rsth, q));
  };
  __shared__ __align__(16) double s_red[8 * 32];
  __shared__ float s_special[4];
  double2 *stw = reinterpret_cast<double2 *>(s_red);
  // a bin k receives m = k (term 0) and m = n - k (term 1, conjugated); further terms (n + k, 2n - k, ...) exist only
  // when n + k <= lmax.  Bins with at most two terms -- every bin of a ring with n > lmax / 2... -- take the batched path.
  const bool deep = (n <= lmax);                      // some bin has three or more terms: generic fold for all bins

  // the fields this CTA handles, in an order that puts 1 before 5 and 2 before 4 (cot terms)
  const int nf = (A.nfg == 1) ? 6 : 2;
#pragma unroll 1
  for (int fi = 0; fi < nf; ++fi) {
    const int field = (A.nfg == 1) ? fi : (fi == 0 ? group : (group == 0 ? 3 : group == 1 ? 5 : 4));
    const BAddr B{b_recv, A.m_boff, A.m_bstr, rpl * 6 + field, hemi};
    // S1 for short rings (n <= lmax, classes that carry a tile buffer): a bin of a ring with n pixels collects ~2 lmax / n
    // values of m -- thousands for the rings next to the poles -- and the reference adds them in ascending m with one float
    // rounding each (alm2allmaps_transpose_mpi.c:836-881), a serial chain.  Gathering them one memory round trip at a time
    // made the ring of FOUR pixels the critical path of the whole stage (3.5 ms, whatever the number of GPUs); instead the
    // CTA streams the ring's b_m (contiguous in m) through a shared-memory tile and every bin walks its own terms there.
    if (deep && A.tileT > 0 && !(A.dbg & 1)) {
      double2 *tile = reinterpret_cast<double2 *>(tailbuf + A.tail);
      const int T = A.tileT;
      for (int k = threadIdx.x; k <= 2 * r; k += blockDim.x) Y[k] = make_float2(0.f, 0.f);
      for (int m0 = 0; m0 <= lmax; m0 += T) {
        const int mend = min(m0 + T, lmax + 1);
        for (int t = threadIdx.x; m0 + t < mend; t += blockDim.x) tile[t] = B(m0 + t);
        __syncthreads();
        for (int k = threadIdx.x; k <= 2 * r; k += blockDim.x) {
          float2 y = Y[k];
          auto apply = [&](int t, double2 b) {                          // term t of the bin: see fold_bin
            const int wraps = (t + 1) >> 1;
            const double sk = (shifted && (wraps & 1)) ? -1.0 : 1.0;
            y.x = __double2float_rn(__dadd_rn((double)y.x, __dmul_rn(b.x, sk)));
            if (t & 1) y.y = __double2float_rn(__dsub_rn((double)y.y, __dmul_rn(b.y, sk)));
            else y.y = __double2float_rn(__dadd_rn((double)y.y, __dmul_rn(b.y, sk)));
          };
          if (k == 0) {   // m = 0 once, then every multiple of n twice: its positive term, then its negative one
            for (long j = (m0 + n - 1) / n; j * n < mend; ++j) {
              const double2 b = tile[j * n - m0];
              if (j == 0) apply(0, b);
              else { apply((int)(2 * j), b); apply((int)(2 * j - 1), b); }
            }
          } else {
            // terms below m0 are done: P positive ones (m = j n + k < m0) and Q negative ones (m = j n - k < m0, j >= 1)
            const int P = (m0 > k) ? (m0 - k + n - 1) / n : 0;
            const int Q = (m0 + k - 1) / n;
            for (int t = P + Q;; ++t) {
              const long mm = (long)((t + 1) >> 1) * n + ((t & 1) ? -k : k);
              if (mm >= mend) break;
              apply(t, tile[mm - m0]);
            }
          }
          Y[k] = y;
        }
        __syncthreads();
      }
      if (shifted)
        for (int k = threadIdx.x; k <= 2 * r; k += blockDim.x) {        // [healpix_shtrans.c:186-197]
          const double2 ph = __ldg(&PT[k]);
          const float2 y = Y[k];
          const double t0 = (double)y.x, t1 = (double)y.y;
          Y[k] = make_float2(__double2float_rn(__dsub_rn(__dmul_rn(t0, ph.x), __dmul_rn(t1, ph.y))),
                             __double2float_rn(__dadd_rn(__dmul_rn(t1, ph.x), __dmul_rn(t0, ph.y))));
        }
    } else {
    // S1: folded, phased float bins.  U bins per thread and trip, the global loads of both terms of all U bins (and of
    // the phase table) issued together: this loop is otherwise bound by one exposed memory round trip per bin
    constexpr int U = 4;
    for (int kb = threadIdx.x; kb <= 2 * r; kb += U * blockDim.x) {
      double2 b0[U], b1[U], ph[U];
      bool valid[U], two[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = kb + u * blockDim.x;
        valid[u] = k <= 2 * r;
        two[u] = valid[u] && !deep && k > 0 && (n - k <= lmax);     // the conjugated term m = n - k lands here too
        b0[u] = make_double2(0.0, 0.0); b1[u] = make_double2(0.0, 0.0); ph[u] = make_double2(1.0, 0.0);
        if (valid[u] && !deep && k <= lmax && !(A.dbg & 1)) b0[u] = B(k);
        if (two[u] && !(A.dbg & 1)) b1[u] = B(n - k);
        if (valid[u] && shifted) ph[u] = __ldg(&PT[k]);               // (cos, sin)(k pi / n), tabulated at plan time
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!valid[u]) continue;
        const int k = kb + u * blockDim.x;
        float2 y;
        if (deep) y = (A.dbg & 1) ? make_float2(1.f, 0.f) : fold_bin(B, k, n, lmax, shifted);
        else {
          // exactly the reference's float accumulation (ascending m; alm2allmaps_transpose_mpi.c:836-881): term 0 has
          // wrap count 0 (sign +1), term 1 wrap count 1 (sign -1 on shifted rings) and enters conjugated
          y = make_float2(__double2float_rn(__dadd_rn(0.0, b0[u].x)), __double2float_rn(__dadd_rn(0.0, b0[u].y)));
          if (two[u]) {
            const double sk = shifted ? -1.0 : 1.0;
            y.x = __double2float_rn(__dadd_rn((double)y.x, __dmul_rn(b1[u].x, sk)));
            y.y = __double2float_rn(__dsub_rn((double)y.y, __dmul_rn(b1[u].y, sk)));
          }
        }
        if (shifted) {                                                  // [healpix_shtrans.c:186-197]
          const double c = ph[u].x, s = ph[u].y;
          const double t0 = (double)y.x, t1 = (double)y.y;
          y.x = __double2float_rn(__dsub_rn(__dmul_rn(t0, c), __dmul_rn(t1, s)));
          y.y = __double2float_rn(__dadd_rn(__dmul_rn(t1, c), __dmul_rn(t0, s)));
        }
        Y[k] = y;
      }
    }
    }
    __syncthreads();
    // c2r samples 0, n/4, n/2, 3n/4 have rational twiddles: exact sums of the float bins (same reason and same
    // formulas as in the oracle's exactly-rounded FFT)
    {
      double cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // R[k&3], I[k&3] over 0 < k < n/2
      for (int k = 1 + threadIdx.x; k < 2 * r; k += blockDim.x) {
        const float2 f = Y[k];
        const int c4 = k & 3;
        if (c4 == 0) { cs[0] += (double)f.x; cs[4] += (double)f.y; }
        else if (c4 == 1) { cs[1] += (double)f.x; cs[5] += (double)f.y; }
        else if (c4 == 2) { cs[2] += (double)f.x; cs[6] += (double)f.y; }
        else { cs[3] += (double)f.x; cs[7] += (double)f.y; }
      }
      cta_sum<8>(cs, s_red);
      if (threadIdx.x == 0) {
        const double y0 = (double)Y[0].x, yh = (double)Y[2 * r].x, sr = (r & 1) ? -1.0 : 1.0;
        s_special[0] = __double2float_rn(y0 + yh + 2.0 * ((cs[0] + cs[1]) + (cs[2] + cs[3])));
        s_special[1] = __double2float_rn(y0 + sr * yh + 2.0 * ((cs[0] - cs[2]) - (cs[5] - cs[7])));
        s_special[2] = __double2float_rn(y0 + yh + 2.0 * ((cs[0] + cs[2]) - (cs[1] + cs[3])));
        s_special[3] = __double2float_rn(y0 + sr * yh + 2.0 * ((cs[0] - cs[2]) + (cs[5] - cs[7])));
      }
      cta_twiddle_tables(stw, logM, tw, logTW);   // (cta_sum ended on a barrier; this one ends on a barrier too)
    }
    // S2: x_{4j+q} = IDFT_r(U^(q))_j with U^(q)_{k'} = E^q sum_p i^{qp} Yfull_{k'+p r}, E = exp(2 pi i k'/n).
    // Two real outputs per complex transform: V1 = U0 + i U1, V2 = U2 + i U3; IDFT(V) = conj(DFT(conj V)).
    for (int k0 = threadIdx.x; k0 < r; k0 += blockDim.x) {
      double2 y[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int k = k0 + p * r;
        float2 f = (k <= 2 * r) ? Y[k] : Y[n - k];
        y[p] = make_double2((double)f.x, (k <= 2 * r) ? (double)f.y : -(double)f.y);
        if (k == 0 || k == 2 * r) y[p].y = 0.0;   // c2r ignores the imaginary parts of the DC and Nyquist bins
      }
      double2 s02 = cadd(y[0], y[2]), d02 = csub(y[0], y[2]), s13 = cadd(y[1], y[3]), d13 = csub(y[1], y[3]);
      double2 T0 = cadd(s02, s13), T2 = csub(s02, s13);
      double2 T1 = cadd(d02, mul_pi(d13)), T3 = csub(d02, mul_pi(d13));
      double2 E1 = __ldg(&PT[2 * k0]), E2 = cmul(E1, E1), E3 = cmul(E1, E2);   // exp(2 pi i k0 / n)
      double2 U1 = cmul(T1, E1), U2 = cmul(T2, E2), U3 = cmul(T3, E3);
      double2 v1 = cconj(cadd(T0, mul_pi(U1)));
      double2 v2 = cconj(cadd(U2, mul_pi(U3)));
      if (bluestein) { bufB[k0] = v2; bufA[swz(k0, logM)] = cmul(v1, __ldg(&chirp[k0])); }
      else { bufB[swz(k0, logM)] = v2; bufA[swz(k0, logM)] = v1; }
    }
    __syncthreads();
    if (A.dbg & 2) {
    } else if (!bluestein) {
      // power-of-two ring: both length-r transforms run in the same passes (all threads busy, half the barriers)
      cta_fft_dif(bufA, logM, stw, bufB);
    } else {
      for (int k = r + threadIdx.x; k < M; k += blockDim.x) bufA[swz(k, logM)] = make_double2(0.0, 0.0);
      __syncthreads();
      cta_dft_r(bufA, r, logM, bluestein, chirp, bhat, stw, false);
      // S5: x^(0)_j = Re(res_j), x^(1)_j = -Im(res_j), rounded to float like the reference's c2r output
      for (int j = threadIdx.x; j < r; j += blockDim.x) {
        double2 res = dft_get_chirp(bufA, j, logM, bluestein, chirp);
        park[j] = make_float2(__double2float_rn(res.x), __double2float_rn(-res.y));
      }
      __syncthreads();
      for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double2 z = make_double2(0.0, 0.0);
        if (j < r) z = cmul(bufB[j], __ldg(&chirp[j]));
        bufA[swz(j, logM)] = z;
      }
      __syncthreads();
      cta_dft_r(bufA, r, logM, bluestein, chirp, bhat, stw, false);
    }
    // S7: float4 of four consecutive pixels, 1/sin(theta) scalings        [alm2allmaps_transpose_mpi.c:1045-1051]
    float4 *out = reinterpret_cast<float4 *>(maps.p[field] + start);
    for (int j = threadIdx.x; j < r; j += blockDim.x) {
      double2 res;
      float2 a;
      if (bluestein) { res = dft_get_chirp(bufA, j, logM, bluestein, chirp); a = park[j]; }
      else {
        const int jr = swz(bitrev(j, logM), logM);
        const double2 r1 = bufA[jr];
        res = bufB[jr];
        a = make_float2(__double2float_rn(r1.x), __double2float_rn(-r1.y));
      }
      float v[4] = {a.x, a.y, __double2float_rn(res.x), __double2float_rn(-res.y)};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int pidx = 4 * j + q;   // pixel index within the ring
        if (pidx == 0) v[q] = s_special[0];
        else if (pidx == r) v[q] = s_special[1];
        else if (pidx == 2 * r) v[q] = s_special[2];
        else if (pidx == 3 * r) v[q] = s_special[3];
      }
      if (field == 2 || field == 4 || field == 5) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = div_sth(v[q]);
      }
      if (field == 5) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = div_sth(v[q]);
      }
      // cot(theta) cross terms [alm2allmaps_transpose_mpi.c:1097-1147]: map4 -= cot * map2, map5 += cot * map1 on northern
      // rings, opposite signs on southern ones.  Fields 1 and 2 of these pixels were stored by this very thread.
      if (field == 4 || field == 5) {
        const float4 o = reinterpret_cast<const float4 *>(maps.p[field == 4 ? 2 : 1] + start)[j];
        const float ov[4] = {o.x, o.y, o.z, o.w};
        const bool sub = (field == 4) != (hemi != 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double c = __dmul_rn(cot, (double)ov[q]);
          v[q] = __double2float_rn(sub ? __dsub_rn((double)v[q], c) : __dadd_rn((double)v[q], c));
        }
      }
      if (!(A.dbg & 4) || v[0] == 1.2345f) out[j] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();   // the work buffers (Y overlays bufA on the Bluestein path) are rewritten by the next field
  }
}

__global__ void __launch_bounds__(512) ring_synthesis_kernel(SynArgs A)
{
  extern __shared__ double2 smem[];
  ring_synthesis_body(A, smem, blockIdx.x, blockIdx.y);
}
__global__ void __launch_bounds__(512) ring_synthesis_scratch_kernel(SynArgs A, double2 *scratch, long stride, int nwork)
{
  double2 *buf = scratch + (long)blockIdx.x * stride;
  for (int item = blockIdx.x; item < nwork * A.nfg; item += gridDim.x) {
    ring_synthesis_body(A, buf, item / A.nfg, item % A.nfg);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t kMaxSmem = 227 * 1024;
int g_fft_field_groups = 0;     // clb_set_tuning(6, 0|1|3): CTAs per ring and hemisphere; 0 = default (3: fields {0,3}, {1,5}, {2,4})
int g_fft_debug = 0;
int g_fft_force_scratch = 0;    // clb_set_tuning(4, 1): run every ring FFT from global scratch (tests the large-ring path at small Nside)
int g_fft_threads_big = 512;   // threads per CTA for work lengths >= 4096 (clb_set_tuning(2, .))
int g_fft_streams = 1;         // clb_set_tuning(7, 0|1): class launches of a stage on parallel streams

// Everything enqueued between fft_fork and fft_join through class_stream() runs after what `st` holds at the fork and before
// what `st` receives after the join.  Calls nest (the solver forks once around the launches of both shells of a pass).
void fft_fork(const ShtPlan *p, cudaStream_t st)
{
  FftTables *t = p->fft;
  if (t->depth++ > 0 || !g_fft_streams) return;
  t->rr = 0;
  CLB_CUDA_CHECK(cudaEventRecord(t->ev_fork, st));
  for (int i = 0; i < 3; ++i) CLB_CUDA_CHECK(cudaStreamWaitEvent(t->aux[i], t->ev_fork, 0));
}
void fft_join(const ShtPlan *p, cudaStream_t st)
{
  FftTables *t = p->fft;
  if (--t->depth > 0 || !g_fft_streams) return;
  for (int i = 0; i < 3; ++i) {
    CLB_CUDA_CHECK(cudaEventRecord(t->ev_join[i], t->aux[i]));
    CLB_CUDA_CHECK(cudaStreamWaitEvent(st, t->ev_join[i], 0));
  }
}
static cudaStream_t class_stream(FftTables *t, cudaStream_t st)
{
  if (!g_fft_streams) return st;
  const int i = t->rr++ & 3;
  return i == 0 ? st : t->aux[i - 1];
}

template <typename T>
static T *to_dev(const std::vector<T> &v)
{
  T *d = nullptr;
  CLB_CUDA_CHECK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CLB_CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return d;
}

static int ilog2_ceil(long v) { int l = 0; while ((1L << l) < v) ++l; return l; }

static RingGeomDev geom_of(const ShtPlan *p)
{
  RingGeomDev g;
  g.cth = p->d_cth; g.sth = p->d_sth; g.weight = p->d_weight; g.nphi = p->d_nphi; g.shifted = p->d_shifted;
  g.startN = p->d_startN; g.startS = p->d_startS;
  return g;
}

static double2 *fft_scratch(FftTables *t, size_t bytes)
{
  if (bytes > t->scratch_bytes) {
    if (t->d_scratch) cudaFree(t->d_scratch);
    CLB_CUDA_CHECK(cudaMalloc(&t->d_scratch, bytes));
    t->scratch_bytes = bytes;
  }
  return t->d_scratch;
}
static int scratch_ctas() { return sm_count() * 2; }

void fft_tables_destroy(ShtPlan *p)
{
  FftTables *t = p->fft;
  if (!t) return;
  for (auto &c : t->classes) cudaFree(c.d_rp);
  for (int i = 0; i < 3; ++i) { if (t->aux[i]) cudaStreamDestroy(t->aux[i]); if (t->ev_join[i]) cudaEventDestroy(t->ev_join[i]); }
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  cudaFree(t->d_tw); cudaFree(t->d_chirp_off); cudaFree(t->d_bhat_off); cudaFree(t->d_chirp); cudaFree(t->d_bhat);
  cudaFree(t->d_phase); cudaFree(t->d_phase_off); cudaFree(t->d_rp_logM); cudaFree(t->d_rp_blu); cudaFree(t->d_scratch);
  delete t;
  p->fft = nullptr;
}

void fft_tables_create(ShtPlan *p)
{
  FftTables *t = new FftTables();
  p->fft = t;
  CLB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 3; ++i) {
    CLB_CUDA_CHECK(cudaStreamCreateWithFlags(&t->aux[i], cudaStreamNonBlocking));
    CLB_CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_join[i], cudaEventDisableTiming));
  }
  const long nside = p->nside;
  // classes over the local ring pairs
  struct Key { int logM, blu; };
  constexpr int kSmallLogM = 11;   // rings with M <= 2048 mix the power-of-two and Bluestein paths in one class per work length
  constexpr int kTinyLogM = 8;     // ... and everything up to M = 256 shares one launch
  std::vector<std::vector<int>> members;
  std::vector<Key> keys;
  int maxLogM = 1;
  std::vector<int> need_r;   // distinct non-power-of-two r among local rings
  std::vector<char> seen(nside + 1, 0);
  std::vector<signed char> rp_logM(p->nrp, 0), rp_blu(p->nrp, 0);
  for (int rp = 0; rp < p->nrp; ++rp) {
    int r = p->h_nphi[rp] / 4;
    int pow2 = (r & (r - 1)) == 0;
    rp_logM[rp] = (signed char)(pow2 ? ilog2_ceil(r) : ilog2_ceil(2L * r - 1));
    rp_blu[rp] = (signed char)!pow2;
  }
  for (int i = 0; i < p->nrp_loc; ++i) {
    int rp = p->rp_loc[i];
    int r = p->h_nphi[rp] / 4;
    int pow2 = !rp_blu[rp];
    int logM = rp_logM[rp];
    maxLogM = std::max(maxLogM, logM);
    // small rings are latency bound one by one: classes of their own (by work length) keep their shared-memory footprint and
    // CTA size small, so that many of them share an SM
    Key key = (logM <= kTinyLogM) ? Key{kTinyLogM, 2} : (logM <= kSmallLogM) ? Key{logM, 2} : Key{logM, !pow2};
    size_t k = 0;
    for (; k < keys.size(); ++k) if (keys[k].logM == key.logM && keys[k].blu == key.blu) break;
    if (k == keys.size()) { keys.push_back(key); members.emplace_back(); }
    members[k].push_back(rp);
    if (!pow2 && !seen[r]) { seen[r] = 1; need_r.push_back(r); }
  }
  t->d_rp_logM = to_dev(rp_logM); t->d_rp_blu = to_dev(rp_blu);
  // phase tables: polar ring pair rp (r = rp+1 < nside) owns entries [r*r-1, r*r+2r], the equatorial ones share a run
  {
    std::vector<long> poff(p->nrp);
    const long eq_off = nside * nside - 1;
    for (int rp = 0; rp < p->nrp; ++rp) { long r = p->h_nphi[rp] / 4; poff[rp] = (rp + 1 < nside) ? r * r - 1 : eq_off; }
    const long total = eq_off + 2 * nside + 1;
    t->d_phase_off = to_dev(poff);
    CLB_CUDA_CHECK(cudaMalloc(&t->d_phase, sizeof(double2) * total));
    phase_table_kernel<<<p->nrp, 256>>>(p->d_nphi, t->d_phase_off, p->nrp, (int)nside, t->d_phase);
    CLB_CUDA_CHECK(cudaGetLastError());
  }
  // twiddles exp(-2 pi i k / TW), k < TW/2
  t->logTW = std::max(maxLogM, 2);
  const long TW = 1L << t->logTW;
  std::vector<double2> tw(TW / 2);
  for (long k = 0; k < TW / 2; ++k) {
    // octant reduction keeps the argument of sin/cos in [0, pi/4]
    long q = (8 * k) / TW, r8 = 8 * k - q * TW;
    double a = (CLB_PI / 4.0) * ((double)r8 / (double)TW);
    if (q & 1) a = (CLB_PI / 4.0) - a;
    double c = cos(a), s = sin(a), re, im;
    switch (q) { case 0: re = c; im = s; break; case 1: re = s; im = c; break; case 2: re = -s; im = c; break;
                 default: re = -c; im = s; break; }
    tw[k] = make_double2(re, -im);
  }
  CLB_CUDA_CHECK(cudaMalloc(&t->d_tw, sizeof(double2) * tw.size()));
  CLB_CUDA_CHECK(cudaMemcpy(t->d_tw, tw.data(), sizeof(double2) * tw.size(), cudaMemcpyHostToDevice));
  // Bluestein tables
  std::vector<long> coff(nside + 1, -1), boff(nside + 1, -1);
  long ctot = 0, btot = 0;
  for (int r : need_r) { coff[r] = ctot; ctot += r; boff[r] = btot; btot += 1L << ilog2_ceil(2L * r - 1); }
  CLB_CUDA_CHECK(cudaMalloc(&t->d_chirp_off, sizeof(long) * (nside + 1)));
  CLB_CUDA_CHECK(cudaMalloc(&t->d_bhat_off, sizeof(long) * (nside + 1)));
  CLB_CUDA_CHECK(cudaMemcpy(t->d_chirp_off, coff.data(), sizeof(long) * (nside + 1), cudaMemcpyHostToDevice));
  CLB_CUDA_CHECK(cudaMemcpy(t->d_bhat_off, boff.data(), sizeof(long) * (nside + 1), cudaMemcpyHostToDevice));
  if (!need_r.empty()) {
    CLB_CUDA_CHECK(cudaMalloc(&t->d_chirp, sizeof(double2) * ctot));
    CLB_CUDA_CHECK(cudaMalloc(&t->d_bhat, sizeof(double2) * btot));
    int *d_rlist;
    CLB_CUDA_CHECK(cudaMalloc(&d_rlist, sizeof(int) * need_r.size()));
    CLB_CUDA_CHECK(cudaMemcpy(d_rlist, need_r.data(), sizeof(int) * need_r.size(), cudaMemcpyHostToDevice));
    size_t smem = sizeof(double2) << maxLogM;
    if (smem <= kMaxSmem && !g_fft_force_scratch) {
      static size_t attr_blu_dev[64] = {};   // per device: the attribute belongs to the function in one context
      int dev = 0;
      CLB_CUDA_CHECK(cudaGetDevice(&dev));
      size_t &attr_blu = attr_blu_dev[dev & 63];
      if (smem > attr_blu) {
        CLB_CUDA_CHECK(cudaFuncSetAttribute(bluestein_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_blu = smem;
      }
      bluestein_table_kernel<<<(unsigned)need_r.size(), 256, smem>>>(d_rlist, (int)need_r.size(), t->d_chirp_off, t->d_bhat_off,
                                                                     t->d_chirp, t->d_bhat, t->d_tw, t->logTW, nullptr, 0);
    } else {
      const int ctas = std::min<int>((int)need_r.size(), scratch_ctas());
      const long stride = 1L << maxLogM;
      double2 *scr = fft_scratch(t, (size_t)ctas * stride * sizeof(double2));
      bluestein_table_kernel<<<ctas, 512>>>(d_rlist, (int)need_r.size(), t->d_chirp_off, t->d_bhat_off, t->d_chirp, t->d_bhat,
                                           t->d_tw, t->logTW, scr, stride);
    }
    CLB_CUDA_CHECK(cudaGetLastError());
    CLB_CUDA_CHECK(cudaDeviceSynchronize());
    cudaFree(d_rlist);
  }
  size_t max_ana = 0, max_syn = 0;
  for (size_t k = 0; k < keys.size(); ++k) {
    FftClass c;
    c.bluestein = keys[k].blu; c.count = (int)members[k].size();
    c.rmax = 0; c.logM = 0; c.tail = 0;
    for (int rp : members[k]) {
      const int r = p->h_nphi[rp] / 4;
      c.rmax = std::max(c.rmax, r);
      c.logM = std::max<int>(c.logM, rp_logM[rp]);
      c.tail = std::max(c.tail, rp_blu[rp] ? r : 2 * r + 1);
    }
    const long M = 1L << c.logM;
    c.threads = (int)std::min<long>(256, std::max<long>(32, M / 4));
    if (M >= 4096) c.threads = g_fft_threads_big;
    c.tail += c.tail & 1;   // the tile behind it holds double2
    c.tileT = (c.logM <= 10) ? kFoldTile : 0;
    c.smem_ana = sizeof(double2) * (M + c.rmax);
    c.smem_syn = sizeof(double2) * (M + c.rmax + 1) + sizeof(float2) * c.tail + sizeof(double2) * c.tileT;
    max_ana = std::max(max_ana, c.smem_ana); max_syn = std::max(max_syn, c.smem_syn);
    // longest rings first: CTAs are dispatched in order, so the short ones fill the tail of the launch
    std::stable_sort(members[k].begin(), members[k].end(), [&](int x, int y) { return p->h_nphi[x] > p->h_nphi[y]; });
    CLB_CUDA_CHECK(cudaMalloc(&c.d_rp, sizeof(int) * c.count));
    CLB_CUDA_CHECK(cudaMemcpy(c.d_rp, members[k].data(), sizeof(int) * c.count, cudaMemcpyHostToDevice));
    t->classes.push_back(c);
  }
  // most work first: the short classes then fill the gaps the long ones leave (fft_fork)
  std::stable_sort(t->classes.begin(), t->classes.end(), [](const FftClass &a, const FftClass &b) {
    return ((long)a.count << a.logM) > ((long)b.count << b.logM);
  });
  // classes that do not fit an SM's shared memory use the persistent global-scratch kernels instead
  max_ana = max_syn = 0;
  for (const auto &c : t->classes) {
    if (c.smem_ana <= kMaxSmem) max_ana = std::max(max_ana, c.smem_ana);
    if (c.smem_syn <= kMaxSmem) max_syn = std::max(max_syn, c.smem_syn);
  }
  // the attribute is per function (and device), not per plan: several plans may be alive (one per emulated rank, or one
  // per resolution), so it is only ever raised
  static size_t attr_ana_dev[64] = {}, attr_syn_dev[64] = {};
  int cur_dev = 0;
  CLB_CUDA_CHECK(cudaGetDevice(&cur_dev));
  size_t &attr_ana = attr_ana_dev[cur_dev & 63], &attr_syn = attr_syn_dev[cur_dev & 63];
  if (max_ana > attr_ana) {
    CLB_CUDA_CHECK(cudaFuncSetAttribute(ring_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_ana));
    attr_ana = max_ana;
  }
  if (max_syn > attr_syn) {
    CLB_CUDA_CHECK(cudaFuncSetAttribute(ring_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_syn));
    attr_syn = max_syn;
  }
}

// defined in sht_plan.cu
int *plan_rp_to_local(const ShtPlan *p);

int launch_ring_analysis(const ShtPlan *p, const float *d_map, double2 *d_g_send, cudaStream_t st)
{
  FftTables *t = p->fft;
  int launches = 0;
  fft_fork(p, st);
  const cudaStream_t st0 = st;
  for (const auto &c : t->classes) {
    st = class_stream(t, st0);
    AnaArgs A{d_map, d_g_send, geom_of(p), c.d_rp, plan_rp_to_local(p), p->d_m_goff, (int)p->lmax, t->d_rp_logM, t->d_rp_blu,
              1 << c.logM, t->d_chirp_off, t->d_bhat_off, t->d_chirp, t->d_bhat, t->d_tw, t->logTW, t->d_phase, t->d_phase_off};
    if (c.smem_ana <= kMaxSmem && !g_fft_force_scratch) {
      ring_analysis_kernel<<<2 * c.count, c.threads, c.smem_ana, st>>>(A);
    } else {
      const int nwork = 2 * c.count, ctas = std::min(nwork, scratch_ctas());
      const long stride = (long)((c.smem_ana + 255) / 256) * 16;   // double2 elements, 256-byte aligned slices
      double2 *scr = fft_scratch(t, (size_t)ctas * stride * sizeof(double2));
      ring_analysis_scratch_kernel<<<ctas, 512, 0, st0>>>(A, scr, stride, nwork);   // (one scratch buffer: these launches stay in line)
    }
    ++launches;
  }
  CLB_CUDA_CHECK(cudaGetLastError());
  fft_join(p, st0);
  return launches;
}

int launch_ring_synthesis(const ShtPlan *p, const double2 *d_b_recv, float *const d_maps[6], cudaStream_t st)
{
  FftTables *t = p->fft;
  MapPtrs mp;
  for (int k = 0; k < 6; ++k) mp.p[k] = d_maps[k];
  int launches = 0;
  fft_fork(p, st);
  const cudaStream_t st0 = st;
  for (const auto &c : t->classes) {
    st = class_stream(t, st0);
    SynArgs A{d_b_recv, mp, geom_of(p), c.d_rp, plan_rp_to_local(p), p->d_m_boff, p->d_m_bstr, (int)p->lmax, t->d_rp_logM,
              t->d_rp_blu, 1 << c.logM, c.rmax, t->d_chirp_off, t->d_bhat_off, t->d_chirp, t->d_bhat, t->d_tw, t->logTW,
              t->d_phase, t->d_phase_off, g_fft_debug, 1, c.tail, c.tileT};
    A.nfg = g_fft_field_groups ? g_fft_field_groups : 3;
    if (c.smem_syn <= kMaxSmem && !g_fft_force_scratch) {
      dim3 grid(2 * c.count, A.nfg);
      ring_synthesis_kernel<<<grid, c.threads, c.smem_syn, st>>>(A);
    } else {
      const int nwork = 2 * c.count, ctas = std::min(nwork * A.nfg, scratch_ctas());
      const long stride = (long)((c.smem_syn + 255) / 256) * 16;
      double2 *scr = fft_scratch(t, (size_t)ctas * stride * sizeof(double2));
      ring_synthesis_scratch_kernel<<<ctas, 512, 0, st0>>>(A, scr, stride, nwork);   // (one scratch buffer: these launches stay in line)
    }
    ++launches;
  }
  CLB_CUDA_CHECK(cudaGetLastError());
  fft_join(p, st0);
  return launches;
}

// Density load: this rank's rings of a full-sky count map -> scaled density in a device map.  src may be device
// memory or PINNED HOST memory (read through the unified address space, so only the rings this rank owns cross PCIe).
//   v = (v * premul) * densmul - backdens, each step rounded to float        [shtpoissonsolve.c:426,468,478]
__global__ void load_density_kernel(const float *__restrict__ src, float *__restrict__ dst, RingGeomDev geo,
                                    const int *__restrict__ rp_loc, int nslots, float premul, float densmul, float backdens)
{
  // a few persistent CTAs walk the rings: enough loads in flight to fill a PCIe link, few enough SM slots that the
  // kernel can run behind the previous plane's compute kernels on a high-priority stream
  for (int slot = blockIdx.x; slot < nslots; slot += gridDim.x) {
    const int rp = rp_loc[slot >> 1];
    const int hemi = slot & 1;
    const long start = hemi ? geo.startS[rp] : geo.startN[rp];
    if (start < 0) continue;
    const int n4 = geo.nphi[rp] >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src + start);
    float4 *d4 = reinterpret_cast<float4 *>(dst + start);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 v = s4[i];
      v.x = __fsub_rn(__fmul_rn(__fmul_rn(v.x, premul), densmul), backdens);
      v.y = __fsub_rn(__fmul_rn(__fmul_rn(v.y, premul), densmul), backdens);
      v.z = __fsub_rn(__fmul_rn(__fmul_rn(v.z, premul), densmul), backdens);
      v.w = __fsub_rn(__fmul_rn(__fmul_rn(v.w, premul), densmul), backdens);
      d4[i] = v;
    }
  }
}
int launch_load_density(const ShtPlan *p, const float *src, float *dst, float premul, float densmul, float backdens,
                        cudaStream_t st)
{
  if (p->nrp_loc == 0) return 0;
  cudaPointerAttributes attr;
  const bool on_device = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  const int nslots = 2 * p->nrp_loc;
  const int grid = on_device ? std::min(nslots, sm_count() * 8) : std::min(nslots, 96);   // PCIe needs few CTAs, HBM many
  load_density_kernel<<<grid, 512, 0, st>>>(src, dst, geom_of(p), p->d_rp_loc, nslots, premul, densmul, backdens);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// Fused exchange of the derivative maps: every rank stores the rings it synthesised into every peer's map buffers
// (NVLink peer stores), replacing the ring -> domain shuffle of map_shuffle.c:22-631 by a broadcast of ring sets.
struct PeerMaps { float *p[8][6]; };
// need[c] (optional): bit q set when rank q's ray domain, grown by the halo margin, touches coarse NEST cell c
// (clb_domain_masks); a group of four pixels goes only to the ranks whose bit is set for its cell(s).
// Four consecutive ring pixels can touch up to four coarse cells (corner cuts): OR over all of them.
__device__ __forceinline__ unsigned group_need(const unsigned char *__restrict__ need, long pix, long order, int coarse_shift)
{
  return need[ring2nest(pix, order) >> coarse_shift] | need[ring2nest(pix + 1, order) >> coarse_shift] |
         need[ring2nest(pix + 2, order) >> coarse_shift] | need[ring2nest(pix + 3, order) >> coarse_shift];
}
// gmask[pix / 4] (optional): the same per group of four pixels, tabulated once (the four ring -> nest conversions per group
// were most of this kernel's time)
__global__ void ring_broadcast_kernel(MapPtrs local, PeerMaps peers, int nranks, int rank, RingGeomDev geo,
                                      const int *__restrict__ rp_loc, const unsigned char *__restrict__ need,
                                      const unsigned char *__restrict__ gmask, long order, int coarse_shift)
{
  const int rp = rp_loc[blockIdx.x >> 1];
  const int hemi = blockIdx.x & 1;
  const long start = hemi ? geo.startS[rp] : geo.startN[rp];
  if (start < 0) return;
  const int n4 = geo.nphi[rp] >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    unsigned m = 0xffu;
    if (gmask) m = gmask[(start >> 2) + i];      // (every ring starts at a multiple of four)
    else if (need) m = group_need(need, start + 4L * i, order, coarse_shift);
    m &= ~(1u << rank);
    if (!m) continue;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const float4 v = reinterpret_cast<const float4 *>(local.p[f] + start)[i];
      for (int q = 0; q < nranks; ++q)
        if ((m >> q) & 1u) reinterpret_cast<float4 *>(peers.p[q][f] + start)[i] = v;
    }
  }
}
// the table for this rank's rings (entries of other rings stay untouched)
__global__ void group_mask_kernel(RingGeomDev geo, const int *__restrict__ rp_loc, const unsigned char *__restrict__ need,
                                  long order, int coarse_shift, unsigned char *__restrict__ gmask)
{
  const int rp = rp_loc[blockIdx.x >> 1];
  const int hemi = blockIdx.x & 1;
  const long start = hemi ? geo.startS[rp] : geo.startN[rp];
  if (start < 0) return;
  const int n4 = geo.nphi[rp] >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) gmask[(start >> 2) + i] = (unsigned char)group_need(need, start + 4L * i, order, coarse_shift);
}
int launch_group_masks(const ShtPlan *p, const unsigned char *d_need, long coarse_order, unsigned char *d_gmask, cudaStream_t st)
{
  if (p->nrp_loc == 0) return 0;
  group_mask_kernel<<<2 * p->nrp_loc, 256, 0, st>>>(geom_of(p), p->d_rp_loc, d_need, p->order, (int)(2 * (p->order - coarse_order)), d_gmask);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_maps_broadcast(const ShtPlan *p, float *const local_maps[6], float *const *peer_maps,
                          const unsigned char *d_need, long coarse_order, cudaStream_t st, const unsigned char *d_gmask)
{
  if (p->nranks <= 1 || p->nrp_loc == 0) return 0;
  if (p->nranks > 8) { fprintf(stderr, "calclens_b200: map broadcast supports up to 8 ranks per node\n"); abort(); }
  MapPtrs loc; PeerMaps peers;
  for (int k = 0; k < 6; ++k) loc.p[k] = local_maps[k];
  for (int q = 0; q < 8; ++q)
    for (int k = 0; k < 6; ++k) peers.p[q][k] = (q < p->nranks) ? peer_maps[q * 6 + k] : nullptr;
  if (d_need && coarse_order > p->order) { d_need = nullptr; d_gmask = nullptr; }
  ring_broadcast_kernel<<<2 * p->nrp_loc, 256, 0, st>>>(loc, peers, p->nranks, p->rank, geom_of(p), p->d_rp_loc, d_need, d_gmask,
                                                        p->order, (int)(2 * (p->order - coarse_order)));
  CLB_CUDA_CHECK(cudaGetLastError());
  return 1;
}

}  // namespace clb
