// calclens_b200/csrc/legendre.cu
// FP64 Legendre transforms for sm_100a with lambda_lm generated on the fly (no P_lm table in HBM).
//   analysis  : a_lm = sum_rings lambda_lm(theta_r) [g_m^N(r) + (-1)^(l+m) g_m^S(r)], then the Poisson filter
//               [map2alm_transpose_mpi.c:427-536; shtpoissonsolve.c:526-550]
//   synthesis : the six ring-space fields phi, d_theta, d_phi, d_theta^2, d_theta d_phi, d_phi^2 per (m, ring)
//               [alm2allmaps_transpose_mpi.c:272-595]
//
// Both kernels put R ring pairs (north ring + its southern mirror, combined through the parity of l+m) on a thread,
// keep their whole state in registers and walk l with the two-instruction scaled recurrence
// mu_{l+1} = (x A_l) mu_l - mu_{l-1} (sht_plan.cu), started from the precomputed per-(m, ring) seeds, so no work is
// spent below the degree where lambda reaches 1e-30.  A warp owns 32*R adjacent ring pairs of one m and runs on its
// own (no block barrier in the l loop); recurrence coefficients and alm-derived coefficients are warp-uniform and reach
// the threads as broadcast reads of a private cp.async tile.  What this instruction mix can reach on the FP64 pipe is
// measured by tools/fp64_probe.cu.
//
// Synthesis uses three complex sums instead of the reference's per-l derivative formulas:
//   P = sum a_l lambda_l,  K = sum l(l+1) a_l lambda_l,
//   D = sin(theta) d_theta f = sum d_l lambda_l with d_l = (l-1) e_l a_{l-1} - (l+2) e_{l+1} a_{l+1},
//   e_l = sqrt((l^2-m^2)/(4l^2-1))   [from sin(t) d_t lambda_l = l e_{l+1} lambda_{l+1} - (l+1) e_l lambda_{l-1}]
//   d_theta^2 f = -cot(theta) d_theta f - K + m^2/sin^2(theta) P      (associated Legendre equation)
// each split by the parity of l+m so that the southern ring costs nothing extra.
#include "sht_internal.cuh"
#include <math.h>

namespace clb {

constexpr int kLB = 16;            // degrees per register block (== kSeedAlign)
constexpr int kLegWarps = 4;       // warps per CTA of both Legendre kernels
constexpr int kLegThreads = kLegWarps * 32;
static_assert(kLB == kSeedAlign, "l-blocks must line up with the seed alignment");

// Ring pairs are dealt in contiguous runs: warp W of the grid row owns the 32*R adjacent ring pairs starting at
// W*32*R, thread `lane` of it the pairs W*32*R + j*32 + lane (j < R).  Adjacent rings reach |lambda| > 1e-30 at almost
// the same degree, so one start degree per warp wastes ~1 % of the work and the whole warp runs
// branch-free: rings that start later carry mu = 0 until their seed is injected at their own (16-aligned) start.

constexpr int kAnaTile = 32;      // degrees of recurrence coefficients per shared-memory tile of the analysis kernel

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// analysis
// ---------------------------------------------------------------------------------------------------------------
// One warp = 32*R adjacent ring pairs of one m, running on its own (no block barrier) from the first degree where any
// of its rings is above 1e-30.  Per degree and ring pair: DMUL + DFMA (recurrence) and two DFMA (re, im accumulate);
// the ring state (mu_{l-1}, mu_l, cos theta, G+, G-) lives in registers for the whole degree range.  After every
// block of KB degrees the 2*KB per-thread partial sums are reduced over the warp with a transpose-reduce (one
// shuffle + add per value) and written to the warp's own partial-sum row; alm_finish_kernel adds the rows of all
// warps of an m in a fixed order (deterministic).
template <int R, int KB, int NB, bool SR>
__global__ void __launch_bounds__(kLegThreads, NB)
legendre_analysis_kernel(const double2 *__restrict__ g_recv, const double2 *const *__restrict__ rp_gsrc,
                         const long *__restrict__ g_off, const int *__restrict__ g_stride, const double *__restrict__ Atab,
                         const long *__restrict__ row_off, const int *__restrict__ ls_tab,
                         const double2 *__restrict__ seed_tab, const double *__restrict__ cth_rp,
                         const int *__restrict__ m_loc, const long *__restrict__ alm_off, double2 *__restrict__ part,
                         long alm_total, int nrp, int lmax)
{
  static_assert(KB == 8 || KB == 16, "block of 8 or 16 degrees");
  static_assert(R % 2 == 0 || R == 1, "start blocks are packed two per register");
  constexpr int V = 2 * KB;   // v[i] = re(l0+i), v[KB+i] = im(l0+i)
  constexpr int NP = (R + 1) / 2;
  __shared__ __align__(16) double s_A[kLegWarps][2][kAnaTile];
  constexpr bool kAnaSmemReduce = (KB == 8) && SR;
  __shared__ double s_red[kAnaSmemReduce ? kLegWarps : 1][kAnaSmemReduce ? 16 * 33 : 1];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int chunk = blockIdx.x * (blockDim.x >> 5) + w, mi = blockIdx.y;
  if (chunk * 32 * R >= nrp) return;     // whole warp beyond the last ring pair (no barriers in this kernel)
  const int m = m_loc[mi];
  unsigned sb[NP];            // start block (ls - m) / 16 of ring j in the (j & 1) half of sb[j / 2]; 0xffff = never
  double mp[R], mc[R], x[R], gpx[R], gpy[R], gmx[R], gmy[R];
  int lsw = kNoStart, lsmax = -1;
  const int rp0 = chunk * 32 * R + lane;
#pragma unroll
  for (int j = 0; j < NP; ++j) sb[j] = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int rp = rp0 + j * 32;
    mp[j] = 0.0; mc[j] = 0.0; x[j] = 0.0; gpx[j] = 0.0; gpy[j] = 0.0; gmx[j] = 0.0; gmy[j] = 0.0;
    int ls = kNoStart;
    if (rp < nrp) ls = ls_tab[(size_t)mi * nrp + rp];
    if (ls != kNoStart) {
      // source: this rank's receive buffer, or (fused exchange) the ring owner's send buffer over NVLink
      const double2 *g = (rp_gsrc ? rp_gsrc[rp] : g_recv + g_off[rp]) + (long)mi * g_stride[rp];
      const double2 gn = g[0], gs = g[1];
      gpx[j] = gn.x + gs.x; gpy[j] = gn.y + gs.y;     // G+ = gN + gS multiplies even l+m
      gmx[j] = gn.x - gs.x; gmy[j] = gn.y - gs.y;     // G- = gN - gS multiplies odd l+m
      x[j] = cth_rp[rp];
      const unsigned blk = (unsigned)(ls - m) / kSeedAlign;
      sb[j / 2] = (j & 1) ? ((sb[j / 2] & 0x0000ffffu) | (blk << 16)) : ((sb[j / 2] & 0xffff0000u) | blk);
      lsw = min(lsw, ls); lsmax = max(lsmax, ls);
    }
  }
  for (int o = 16; o; o >>= 1) {
    lsw = min(lsw, __shfl_xor_sync(0xffffffffu, lsw, o));
    lsmax = max(lsmax, __shfl_xor_sync(0xffffffffu, lsmax, o));
  }
  double2 *out = part + (size_t)chunk * alm_total + alm_off[mi];   // index l - m
  // degrees below the first active block get exact zeros
  {
    const int lz = (lsw == kNoStart) ? lmax + 1 : lsw;
    for (int l = m + lane; l < lz; l += 32) out[l - m] = make_double2(0.0, 0.0);
    if (lsw == kNoStart) return;
  }
  // recurrence coefficients A_l stream through a private double-buffered tile of kAnaTile degrees (cp.async);
  // rows are zero padded beyond lmax+1 (kRowPad), so whole tiles can be read and computed without bounds checks
  const double *Arow = Atab + row_off[mi] + (lsw - m);
  double *sA = &s_A[w][0][0];
  if (lane < kAnaTile / 2) cp_async16(sA + 2 * lane, Arow + 2 * lane);
  cp_async_commit();
  int cur = 0;
  for (int lt = lsw; lt <= lmax; lt += kAnaTile, cur ^= 1) {
    Arow += kAnaTile;
    if (lane < kAnaTile / 2) cp_async16(sA + (cur ^ 1) * kAnaTile + 2 * lane, Arow + 2 * lane);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const double *sa = sA + cur * kAnaTile;
#pragma unroll
    for (int b = 0; b < kAnaTile / KB; ++b) {
      const int l0 = lt + b * KB;
      if (l0 > lmax) break;
      if ((b * KB) % kSeedAlign == 0 && l0 <= lsmax) {   // start-up phase of this warp: inject the seeds of rings starting here
        const unsigned blk = (unsigned)(l0 - m) / kSeedAlign;
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (((sb[j / 2] >> (16 * (j & 1))) & 0xffffu) == blk) {
            const double2 sd = seed_tab[(size_t)mi * nrp + rp0 + j * 32];
            mp[j] = sd.x; mc[j] = sd.y;
          }
      }
      double v[V];
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = 0.0;
#pragma unroll
      for (int i = 0; i < KB; ++i) {
        const double a = sa[b * KB + i];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const double mu = mc[j];
          if (i & 1) { v[i] = fma(mu, gmx[j], v[i]); v[KB + i] = fma(mu, gmy[j], v[KB + i]); }
          else       { v[i] = fma(mu, gpx[j], v[i]); v[KB + i] = fma(mu, gpy[j], v[KB + i]); }
          const double mn = fma(x[j] * a, mu, -mp[j]);
          mp[j] = mu; mc[j] = mn;
        }
      }
      if (kAnaSmemReduce) {
        // warp reduction through shared memory: every lane parks its V partial sums ([value][lane], rows padded to 33 so
        // that the column reads below spread over all banks), then lane L adds 16 lanes' worth of value L % 16 and one
        // shuffle joins the two halves -- ~3x fewer instructions than the shuffle transpose-reduce, which matters because
        // the reduction shares the issue slots of the warps that feed the FP64 pipe
        static_assert(!kAnaSmemReduce || V == 16, "shared-memory reduction is written for blocks of 8 degrees");
        double *red = &s_red[w][0];
#pragma unroll
        for (int i = 0; i < V; ++i) red[i * 33 + lane] = v[i];
        __syncwarp();
        const double *col = red + (lane & 15) * 33 + (lane >> 4) * 16;
        double t0 = col[0], t1 = col[1], t2 = col[2], t3 = col[3];
#pragma unroll
        for (int k = 4; k < 16; k += 4) { t0 += col[k]; t1 += col[k + 1]; t2 += col[k + 2]; t3 += col[k + 3]; }
        double t = (t0 + t1) + (t2 + t3);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        const double ti = __shfl_down_sync(0xffffffffu, t, KB);   // imaginary part lives KB lanes up
        if (lane < KB && l0 + lane <= lmax) out[l0 - m + lane] = make_double2(t, ti);
        __syncwarp();   // the parked sums are consumed before the next block overwrites them
      } else {
      // warp transpose-reduce: afterwards lane L holds the warp total of v[L % V]
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
      const double ti = __shfl_down_sync(0xffffffffu, v[0], KB);   // imaginary part lives KB lanes up
      if (lane < KB && l0 + lane <= lmax) out[l0 - m + lane] = make_double2(v[0], ti);
      }
    }
    __syncwarp();   // everyone is done with this tile before the next iteration's copy overwrites it
  }
  cp_async_wait<0>();
}

// sum the ring chunks in a fixed order, undo the recurrence scaling (lambda = c mu) and apply the Poisson filter
// alm *= -1/(l(l+1)), a_00 = 0                                        [shtpoissonsolve.c:526-550]
__global__ void alm_finish_kernel(const double2 *__restrict__ part, int nchunk, long alm_total,
                                  const int *__restrict__ m_loc, int nm_loc, const long *__restrict__ alm_off,
                                  const long *__restrict__ row_off, const double *__restrict__ ctab, int lmax,
                                  int apply_filter, double *__restrict__ alm_re, double *__restrict__ alm_im)
{
  const int mi = blockIdx.y;
  const int m = m_loc[mi];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= lmax - m; k += gridDim.x * blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      double2 p = part[(size_t)c * alm_total + alm_off[mi] + k];
      re += p.x; im += p.y;
    }
    const double cl = ctab[row_off[mi] + k];
    re *= cl; im *= cl;
    if (apply_filter) {
      const int l = m + k;
      if (l == 0) { re = 0.0; im = 0.0; }
      else {
        const double f = -1.0 / ((double)l) / (((double)l) + 1.0);
        re *= f; im *= f;
      }
    }
    alm_re[alm_off[mi] + k] = re;
    alm_im[alm_off[mi] + k] = im;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// synthesis
// ---------------------------------------------------------------------------------------------------------------
// per-plane coefficient records, one per (m, l), l = m .. lmax+1: {A_l, c P, c D, c K} with the recurrence scaling
// folded in; rows are zero-padded so whole tiles can be processed without bounds checks
__global__ void synthesis_coef_kernel(const double *__restrict__ alm_re, const double *__restrict__ alm_im,
                                      const int *__restrict__ m_loc, const long *__restrict__ alm_off,
                                      const long *__restrict__ row_off, const double *__restrict__ Atab,
                                      const double *__restrict__ ctab, int lmax, double *__restrict__ coef)
{
  const int mi = blockIdx.y;
  const int m = m_loc[mi];
  const long len = row_off[mi + 1] - row_off[mi];
  const double *ar = alm_re + alm_off[mi], *ai = alm_im + alm_off[mi];
  for (long k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += (long)gridDim.x * blockDim.x) {
    const long l = m + k;
    double rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (l <= lmax + 1) {
      const double cl = ctab[row_off[mi] + k];
      double pr = 0, pi = 0, dr = 0, di = 0;
      if (l <= lmax) { pr = ar[k]; pi = ai[k]; }
      const double m2 = (double)m * (double)m;
      if (l - 1 >= m) {
        const double dl = (double)l;
        const double e = sqrt((dl * dl - m2) / (4.0 * dl * dl - 1.0));
        dr += (dl - 1.0) * e * ar[k - 1]; di += (dl - 1.0) * e * ai[k - 1];
      }
      if (l + 1 <= lmax) {
        const double dl1 = (double)(l + 1);
        const double e = sqrt((dl1 * dl1 - m2) / (4.0 * dl1 * dl1 - 1.0));
        dr -= ((double)l + 2.0) * e * ar[k + 1]; di -= ((double)l + 2.0) * e * ai[k + 1];
      }
      const double ll1 = (double)l * ((double)l + 1.0);
      rec[0] = Atab[row_off[mi] + k];
      rec[1] = cl * pr; rec[2] = cl * pi; rec[3] = cl * dr; rec[4] = cl * di;
      rec[5] = cl * ll1 * pr; rec[6] = cl * ll1 * pi;
    }
    double4 *o = reinterpret_cast<double4 *>(coef + 8 * (row_off[mi] + k));
    o[0] = make_double4(rec[0], rec[1], rec[2], rec[3]);
    o[1] = make_double4(rec[4], rec[5], rec[6], rec[7]);
  }
}

// One warp = 32*R adjacent ring pairs of one m, independent of the other warps of its CTA (no block barrier): it
// streams the per-degree coefficient records {A, P, D, K} of its own degree range through a private double-buffered
// shared-memory tile (cp.async, 1 KB per 16 degrees) and reads them back as broadcast LDS.128 -- two per degree for
// 8*R FP64 instructions, which keeps the shared-memory pipe at ~1/4 of the FP64 pipe's time for R = 4.
template <int R>
__global__ void __launch_bounds__(kLegThreads, (R >= 4) ? 3 : 4)
legendre_synthesis_kernel(const double *__restrict__ coef, const long *__restrict__ row_off,
                          const int *__restrict__ ls_tab, const double2 *__restrict__ seed_tab,
                          const double *__restrict__ cth_rp, const double *__restrict__ sth_rp,
                          const int *__restrict__ m_loc, const long *__restrict__ b_off,
                          const int *__restrict__ b_stride, double2 *__restrict__ b_send,
                          double2 *const *__restrict__ rp_bptr, int nrp, int lmax)
{
  __shared__ __align__(16) double s_tile[kLegWarps][2][kLB * 8];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c = blockIdx.x, mi = blockIdx.y;
  const int m = m_loc[mi];
  int ls[R];
  double mp[R], mc[R], x[R];
  double acc[R][12];   // [parity(0 even,1 odd)*6 + {Pre,Pim,Dre,Dim,Kre,Kim}]
  int lsmin = kNoStart;
  const int rp0 = (c * (blockDim.x >> 5) + w) * 32 * R + lane;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int rp = rp0 + j * 32;
    ls[j] = kNoStart; mp[j] = 0.0; mc[j] = 0.0; x[j] = 0.0;
    if (rp < nrp) {
      ls[j] = ls_tab[(size_t)mi * nrp + rp];
      if (ls[j] != kNoStart) x[j] = cth_rp[rp];
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[j][k] = 0.0;
    lsmin = min(lsmin, ls[j]);
  }
  for (int o = 16; o; o >>= 1) lsmin = min(lsmin, __shfl_xor_sync(0xffffffffu, lsmin, o));
  if (lsmin != kNoStart) {
    // tiles start at lsmin (a multiple of kLB above m) and run through degree lmax+1; rows are zero padded (kRowPad)
    const double *crow = coef + 8 * (row_off[mi] + (long)(lsmin - m));
    double *tile = &s_tile[w][0][0];
    // lane moves bytes [32*lane, 32*lane+32) of the 1 KB tile
    cp_async16(tile + 4 * lane, crow + 4 * lane);
    cp_async16(tile + 4 * lane + 2, crow + 4 * lane + 2);
    cp_async_commit();
    int buf = 0;
    for (int l0 = lsmin; l0 <= lmax + 1; l0 += kLB, buf ^= 1) {
      crow += kLB * 8;
      double *nxt = &s_tile[w][buf ^ 1][0];
      cp_async16(nxt + 4 * lane, crow + 4 * lane);
      cp_async16(nxt + 4 * lane + 2, crow + 4 * lane + 2);
      cp_async_commit();
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (ls[j] == l0) {
          const double2 sd = seed_tab[(size_t)mi * nrp + rp0 + j * 32];
          mp[j] = sd.x; mc[j] = sd.y;
        }
      cp_async_wait<1>();
      __syncwarp();
      const double *t = &s_tile[w][buf][0];
#pragma unroll
      for (int i = 0; i < kLB; ++i) {
        const double4 ra = *reinterpret_cast<const double4 *>(&t[i * 8]);
        const double4 rb = *reinterpret_cast<const double4 *>(&t[i * 8 + 4]);
        const int par = (i & 1) * 6;   // (l0 - m) is even, so the parity of l+m is the parity of i
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
      __syncwarp();   // everyone is done with this tile before the next iteration's copy overwrites it
    }
    cp_async_wait<0>();
  }
  // combine parities into north / south and form the six fields
  const double dm = (double)m, m2 = dm * dm;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int rp = rp0 + j * 32;
    if (rp >= nrp) continue;
    const double sth = sth_rp[rp], cth = cth_rp[rp];
    const double isth = 1.0 / sth, cot = cth * isth, m2s2 = m2 * isth * isth;
    // destination: this rank's send buffer, or (fused exchange) the ring owner's receive buffer over NVLink
    double2 *o = (rp_bptr ? rp_bptr[rp] : b_send + b_off[rp]) + (long)mi * 6 * b_stride[rp];
    const long fs = b_stride[rp];
    double2 q[6][2];   // [field][hemisphere]
#pragma unroll
    for (int hemi = 0; hemi < 2; ++hemi) {
      const double sg = hemi ? -1.0 : 1.0;
      const double Pr = acc[j][0] + sg * acc[j][6], Pi = acc[j][1] + sg * acc[j][7];
      const double Dr = acc[j][2] + sg * acc[j][8], Di = acc[j][3] + sg * acc[j][9];
      const double Kr = acc[j][4] + sg * acc[j][10], Ki = acc[j][5] + sg * acc[j][11];
      const double q1r = Dr * isth, q1i = Di * isth;
      const double q3r = -sg * cot * q1r - Kr + m2s2 * Pr, q3i = -sg * cot * q1i - Ki + m2s2 * Pi;
      q[0][hemi] = make_double2(Pr, Pi);                       // phi
      q[1][hemi] = make_double2(q1r, q1i);                     // d_theta
      q[2][hemi] = make_double2(-dm * Pi, dm * Pr);            // i m P          (still to be divided by sin)
      q[3][hemi] = make_double2(q3r, q3i);                     // d_theta^2
      q[4][hemi] = make_double2(-dm * q1i, dm * q1r);          // i m d_theta    (.. / sin)
      q[5][hemi] = make_double2(-m2 * Pr, -m2 * Pi);           // -m^2 P         (.. / sin^2)
    }
    // the (north, south) pair of a field is 32 contiguous, 32-byte aligned bytes: one 256-bit store each (full sectors
    // on the local path, full packets over NVLink)
#pragma unroll
    for (int f = 0; f < 6; ++f)
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o + f * fs), "d"(q[f][0].x), "d"(q[f][0].y),
                   "d"(q[f][1].x), "d"(q[f][1].y) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
int g_leg_warps_per_cta = 4;      // warps are independent in both Legendre kernels; clb_set_tuning(3, 1|2|4)
int g_syn_rings_per_thread = 4;   // tunable through clb_set_tuning(0, .)
int g_ana_rings_per_thread = 8;   // tunable through clb_set_tuning(1, .): 8 (blocks of 8 degrees) or 4, 2, 1 (16 degrees)

int g_ana_smem_reduce = 1;        // clb_set_tuning(5, 0|1): warp reduction of the analysis kernel through shared memory (blocks of 8 degrees)

template <int R, int KB, int NB>
static void launch_ana_t(const ShtPlan *p, const double2 *g_recv, const double2 *const *gsrc, int nchunk, cudaStream_t st)
{
  const int warps = g_leg_warps_per_cta;
  dim3 grid((nchunk + warps - 1) / warps, p->nm_loc);
  if (KB == 8 && g_ana_smem_reduce)
    legendre_analysis_kernel<R, KB, NB, true><<<grid, 32 * warps, 0, st>>>(
        g_recv, gsrc, p->d_g_off, p->d_g_stride, p->d_A, p->d_row_off, p->d_ls_ana, p->d_seed, p->d_cth, p->d_m_loc,
        p->d_alm_off, reinterpret_cast<double2 *>(p->d_part), p->alm_total, p->nrp, (int)p->lmax);
  else
    legendre_analysis_kernel<R, KB, NB, false><<<grid, 32 * warps, 0, st>>>(
        g_recv, gsrc, p->d_g_off, p->d_g_stride, p->d_A, p->d_row_off, p->d_ls_ana, p->d_seed, p->d_cth, p->d_m_loc,
        p->d_alm_off, reinterpret_cast<double2 *>(p->d_part), p->alm_total, p->nrp, (int)p->lmax);
}

int launch_legendre_analysis(ShtPlan *p, const double2 *d_g_recv, double *d_alm_re, double *d_alm_im, int apply_filter,
                             cudaStream_t st)
{
  if (p->nm_loc == 0) return 0;
  // rings per thread: the tuned value, stepped down through the instantiated set while most of a warp would be left
  // without rings (small maps); nchunk is derived from the final R only
  static const int kAnaR[] = {12, 10, 8, 6, 4, 2, 1};
  int ri = 0;
  while (kAnaR[ri] != g_ana_rings_per_thread) {
    if (++ri >= 7) { fprintf(stderr, "calclens_b200: analysis rings per thread %d has no instantiation\n", g_ana_rings_per_thread); abort(); }
  }
  while (kAnaR[ri] > 1 && p->nrp < 32 * kAnaR[ri]) ++ri;
  const int R = kAnaR[ri];
  const int nchunk = (p->nrp + 32 * R - 1) / (32 * R);        // one partial-sum row per warp
  if (!p->d_part || p->ana_nchunk != nchunk) {
    if (p->d_part) cudaFree(p->d_part);
    CLB_CUDA_CHECK(cudaMalloc(&p->d_part, sizeof(double2) * (size_t)nchunk * p->alm_total));
    p->ana_nchunk = nchunk;
  }
  // fused exchange (clb_sht_plan_set_peers): g is pulled from the ring owners' send buffers only when the caller passes
  // no receive buffer; a caller that hands in g_recv gets the plain local path
  const double2 *const *gsrc = d_g_recv ? nullptr : p->d_rp_gsrc;
  if (!d_g_recv && !gsrc) { fprintf(stderr, "calclens_b200: legendre analysis needs g_recv (no peer buffers are set)\n"); abort(); }
  switch (R) {
    case 12: launch_ana_t<12, 8, 2>(p, d_g_recv, gsrc, nchunk, st); break;
    case 10: launch_ana_t<10, 16, 2>(p, d_g_recv, gsrc, nchunk, st); break;
    case 8: launch_ana_t<8, 8, 3>(p, d_g_recv, gsrc, nchunk, st); break;
    case 6: launch_ana_t<6, 16, 3>(p, d_g_recv, gsrc, nchunk, st); break;
    case 4: launch_ana_t<4, 16, 3>(p, d_g_recv, gsrc, nchunk, st); break;
    case 2: launch_ana_t<2, 16, 4>(p, d_g_recv, gsrc, nchunk, st); break;
    case 1: launch_ana_t<1, 16, 4>(p, d_g_recv, gsrc, nchunk, st); break;
    default: fprintf(stderr, "calclens_b200: analysis rings per thread %d has no instantiation\n", R); abort();
  }
  CLB_CUDA_CHECK(cudaGetLastError());
  dim3 grid((unsigned)((p->lmax + 256) / 256), p->nm_loc);
  alm_finish_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const double2 *>(p->d_part), nchunk, p->alm_total, p->d_m_loc,
                                          p->nm_loc, p->d_alm_off, p->d_row_off, p->d_c, (int)p->lmax, apply_filter,
                                          d_alm_re, d_alm_im);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 2;
}

int launch_legendre_synthesis(ShtPlan *p, const double *d_alm_re, const double *d_alm_im, double2 *d_b_send,
                              cudaStream_t st)
{
  if (p->nm_loc == 0) return 0;
  dim3 cgrid((unsigned)((p->lmax + 2 + kRowPad + 32 + 255) / 256), p->nm_loc);
  synthesis_coef_kernel<<<cgrid, 256, 0, st>>>(d_alm_re, d_alm_im, p->d_m_loc, p->d_alm_off, p->d_row_off, p->d_A, p->d_c,
                                               (int)p->lmax, p->d_coef);
  CLB_CUDA_CHECK(cudaGetLastError());
  int R = g_syn_rings_per_thread;
  while (R > 1 && p->nrp < 32 * R) --R;   // 4, 3, 2, 1 are all instantiated
  // fused exchange: b is pushed into the ring owners' receive buffers only when the caller passes no send buffer
  double2 *const *bptr = d_b_send ? nullptr : p->d_rp_bptr;
  if (!d_b_send && !bptr) { fprintf(stderr, "calclens_b200: legendre synthesis needs b_send (no peer buffers are set)\n"); abort(); }
  const int warps = g_leg_warps_per_cta;
  const int nchunk = (p->nrp + R * 32 * warps - 1) / (R * 32 * warps);
  dim3 grid(nchunk, p->nm_loc);
#define CLB_SYN_LAUNCH(RR)                                                                                           \
  legendre_synthesis_kernel<RR><<<grid, 32 * warps, 0, st>>>(p->d_coef, p->d_row_off, p->d_ls_syn, p->d_seed, p->d_cth, \
                                                              p->d_sth, p->d_m_loc, p->d_b_off, p->d_b_stride, d_b_send, \
                                                              bptr, p->nrp, (int)p->lmax)
  switch (R) {
    case 4: CLB_SYN_LAUNCH(4); break;
    case 3: CLB_SYN_LAUNCH(3); break;
    case 2: CLB_SYN_LAUNCH(2); break;
    default: CLB_SYN_LAUNCH(1); break;
  }
#undef CLB_SYN_LAUNCH
  CLB_CUDA_CHECK(cudaGetLastError());
  return 2;
}

}  // namespace clb
