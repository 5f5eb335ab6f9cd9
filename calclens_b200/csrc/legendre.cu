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
#include <algorithm>
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
// One warp = 32*R adjacent ring pairs ("chunk") of one m at a time, running on its own (no block barrier) from the first
// degree where any of its rings is above 1e-30.  Per degree and ring pair: DMUL + DFMA (recurrence, shared by the NS
// shells of a batched pass) and two DFMA (re, im accumulate) per shell; the ring state (mu_{l-1}, mu_l, cos theta, G+, G-
// of every shell) lives in registers for the whole degree range.  After every block of KB = 8/NS degrees the 16
// per-thread partial sums are reduced over the warp through shared memory and land in the warp's partial-sum row.
// A warp walks several chunks one after the other (chunk = row, row + rows, ...: every warp gets the same mix of polar and
// equatorial rings, so the warps of a CTA finish together) and accumulates them into the same row, which it alone owns:
// the first chunk stores, the later ones add with red.global (same thread, same address: program order, so the sum is
// deterministic).  alm_finish_kernel adds the `rows` rows of an m in a fixed order.
template <int R, int NS, int NB, bool PIPE>
__global__ void __launch_bounds__(kLegThreads, NB)
legendre_analysis_kernel(const double2 *__restrict__ g_recv, long g_shell, const double2 *const *__restrict__ rp_gsrc,
                         const long *__restrict__ g_off, const int *__restrict__ g_stride, const double *__restrict__ Atab,
                         const long *__restrict__ row_off, const int *__restrict__ ls_tab,
                         const double2 *__restrict__ seed_tab, const double *__restrict__ cth_rp,
                         const int *__restrict__ m_loc, const long *__restrict__ alm_off, double2 *__restrict__ part,
                         long part_shell, long alm_total, int nrp, int lmax, int rows, int nchunk)
{
  static_assert(NS == 1 || NS == 2, "one or two shells per pass");
  static_assert(R % 2 == 0 || R == 1, "start blocks are packed two per register");
  constexpr int KB = 8 / NS;  // degrees per block: 2 * KB * NS = 16 partial sums per thread either way
  constexpr int V = 16;       // v[s*2*KB + i] = re(shell s, l0+i), v[s*2*KB + KB + i] = im(shell s, l0+i)
  constexpr int NP = (R + 1) / 2;
  // dynamic shared memory (more than the 48 KB static limit for the larger variants), per warp:
  //   A tiles [2][kAnaTile] | parked partial sums [PIPE ? 2 : 1][16 * 33] | the chunk's seeds [R][32] (double2), copied in
  //   with the first coefficient tile
  extern __shared__ __align__(16) double s_dyn[];
  constexpr int kRedBufs = PIPE ? 2 : 1;
  constexpr int kWarpDoubles = 2 * kAnaTile + kRedBufs * 16 * 33 + 2 * R * 32;

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = blockIdx.x * (blockDim.x >> 5) + w, mi = blockIdx.y;
  if (row >= rows) return;     // (no block barriers in this kernel)
  const int m = m_loc[mi];
  double2 *out = part + (size_t)row * alm_total + alm_off[mi];   // index l - m; shell s at + s * part_shell
  double *sA = s_dyn + (size_t)w * kWarpDoubles;
  double *red = sA + 2 * kAnaTile;
  double2 *seeds = reinterpret_cast<double2 *>(red + kRedBufs * 16 * 33);   // [R][32]; 16-byte aligned: all counts are even
  // which (shell, degree) pair of a block this lane delivers after the reduction
  const int osh = lane / (2 * KB), oi = lane % (2 * KB);
  const bool writer = lane < V && oi < KB;
  bool first = true;
  // most equatorial chunk first: it starts at the lowest degree, so everything a later chunk adds to is already there
  for (int chunk = row + ((nchunk - 1 - row) / rows) * rows; chunk >= 0; chunk -= rows) {
    unsigned sb[NP];            // start block (ls - m) / 16 of ring j in the (j & 1) half of sb[j / 2]; 0xffff = never
    double mp[R], mc[R], x[R], gpx[NS][R], gpy[NS][R], gmx[NS][R], gmy[NS][R];
    int lsw = kNoStart, lsmax = -1;
    const int rp0 = chunk * 32 * R + lane;
#pragma unroll
    for (int j = 0; j < NP; ++j) sb[j] = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int rp = rp0 + j * 32;
      mp[j] = 0.0; mc[j] = 0.0; x[j] = 0.0;
#pragma unroll
      for (int s = 0; s < NS; ++s) { gpx[s][j] = 0.0; gpy[s][j] = 0.0; gmx[s][j] = 0.0; gmy[s][j] = 0.0; }
      int ls = kNoStart;
      if (rp < nrp) ls = ls_tab[(size_t)mi * nrp + rp];
      if (ls != kNoStart) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          // source: this rank's receive buffer, or (fused exchange) the ring owner's send buffer over NVLink
          const double2 *g = (rp_gsrc ? rp_gsrc[rp + s * nrp] : g_recv + s * g_shell + g_off[rp]) + (long)mi * g_stride[rp];
          const double2 gn = g[0], gs = g[1];
          gpx[s][j] = gn.x + gs.x; gpy[s][j] = gn.y + gs.y;     // G+ = gN + gS multiplies even l+m
          gmx[s][j] = gn.x - gs.x; gmy[s][j] = gn.y - gs.y;     // G- = gN - gS multiplies odd l+m
        }
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
    if (lsw == kNoStart) continue;   // every ring of this chunk is cut for this m
    // (a global load at the moment a ring starts would stall the whole warp on HBM latency in the middle of the FP64 work)
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (rp0 + j * 32 < nrp) cp_async16(&seeds[j * 32 + lane], &seed_tab[(size_t)mi * nrp + rp0 + j * 32]);
    const bool rmw = !first;
    if (first) {                     // degrees below the first active block get exact zeros
      for (int l = m + lane; l < lsw; l += 32)
#pragma unroll
        for (int s = 0; s < NS; ++s) out[s * part_shell + l - m] = make_double2(0.0, 0.0);
      first = false;
      __syncwarp();                  // later chunks read these entries from other lanes
    }
    // recurrence coefficients A_l stream through a private double-buffered tile of kAnaTile degrees (cp.async);
    // rows are zero padded beyond lmax+1 (kRowPad), so whole tiles can be read and computed without bounds checks
    const double *Arow = Atab + row_off[mi] + (lsw - m);
    if (lane < kAnaTile / 2) cp_async16(sA + 2 * lane, Arow + 2 * lane);
    cp_async_commit();
    // PIPE: the warp sum of block b is taken while block b+1 is computed -- the block's partial sums wait in one of two
    // shared-memory buffers, a quarter of the column sum rides on every (second) degree of the next block, and the result
    // leaves at the end of that block: the FP64 pipe has work during the whole reduction.  Same values, same order of adds.
    auto deliver = [&](double2 *dst, double t, double ti) {
      if (!rmw) *dst = make_double2(t, ti);
      else {   // a later chunk of the same row: fire-and-forget adds in L2 (this thread alone touches the address, in program order)
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&dst->x), "d"(t) : "memory");
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&dst->y), "d"(ti) : "memory");
      }
    };
    bool pend = false;          // a parked block is waiting for its warp sum
    int pbuf = 0;               // buffer the next block parks in
    double2 *pdst = out;        // destination of the waiting block (valid when pok)
    bool pok = false;
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
        double2 *dst = out + osh * part_shell + (l0 - m + oi);
        if ((b * KB) % kSeedAlign == 0 && l0 <= lsmax) {   // start-up phase of this warp: inject the seeds of rings starting here
          const unsigned blk = (unsigned)(l0 - m) / kSeedAlign;
#pragma unroll
          for (int j = 0; j < R; ++j)
            if (((sb[j / 2] >> (16 * (j & 1))) & 0xffffu) == blk) {
              const double2 sd = seeds[j * 32 + lane];
              mp[j] = sd.x; mc[j] = sd.y;
            }
        }
        // columns of the waiting block (PIPE; stale data when nothing waits -- computed, never delivered)
        const double *pcol = red + (pbuf ^ 1) * (16 * 33) + (lane & 15) * 33 + (lane >> 4) * 16;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        double v[V];
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = 0.0;
#pragma unroll
        for (int i = 0; i < KB; ++i) {
          if (PIPE && i % (KB / 4) == 0) {
            const int k = 4 * (i / (KB / 4));
            if (k == 0) { t0 = pcol[0]; t1 = pcol[1]; t2 = pcol[2]; t3 = pcol[3]; }
            else { t0 += pcol[k]; t1 += pcol[k + 1]; t2 += pcol[k + 2]; t3 += pcol[k + 3]; }
          }
          const double a = sa[b * KB + i];
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const double mu = mc[j];
#pragma unroll
            for (int s = 0; s < NS; ++s) {   // (l0 - m) is even, so the parity of l+m is the parity of i
              if (i & 1) { v[s * 2 * KB + i] = fma(mu, gmx[s][j], v[s * 2 * KB + i]); v[s * 2 * KB + KB + i] = fma(mu, gmy[s][j], v[s * 2 * KB + KB + i]); }
              else       { v[s * 2 * KB + i] = fma(mu, gpx[s][j], v[s * 2 * KB + i]); v[s * 2 * KB + KB + i] = fma(mu, gpy[s][j], v[s * 2 * KB + KB + i]); }
            }
            const double mn = fma(x[j] * a, mu, -mp[j]);
            mp[j] = mu; mc[j] = mn;
          }
        }
        // warp reduction through shared memory: every lane parks its 16 partial sums ([value][lane], rows padded to 33 so
        // that the column reads spread over all banks), then lane L adds 16 lanes' worth of value L % 16 and one
        // shuffle joins the two halves -- ~3x fewer instructions than a shuffle transpose-reduce, which matters because
        // the reduction shares the issue slots of the warps that feed the FP64 pipe
        if (PIPE) {
          double t = (t0 + t1) + (t2 + t3);
          t += __shfl_xor_sync(0xffffffffu, t, 16);
          const double ti = __shfl_down_sync(0xffffffffu, t, KB);   // imaginary part lives KB lanes up
          if (pend && pok) deliver(pdst, t, ti);
          double *park = red + pbuf * (16 * 33);
#pragma unroll
          for (int i = 0; i < V; ++i) park[i * 33 + lane] = v[i];
          __syncwarp();   // parked for the next block to sum; and every lane is done reading the other buffer
          pend = true; pdst = dst; pok = writer && l0 + oi <= lmax; pbuf ^= 1;
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) red[i * 33 + lane] = v[i];
          __syncwarp();
          const double *col = red + (lane & 15) * 33 + (lane >> 4) * 16;
          t0 = col[0]; t1 = col[1]; t2 = col[2]; t3 = col[3];
#pragma unroll
          for (int k = 4; k < 16; k += 4) { t0 += col[k]; t1 += col[k + 1]; t2 += col[k + 2]; t3 += col[k + 3]; }
          double t = (t0 + t1) + (t2 + t3);
          t += __shfl_xor_sync(0xffffffffu, t, 16);
          const double ti = __shfl_down_sync(0xffffffffu, t, KB);   // imaginary part lives KB lanes up
          if (writer && l0 + oi <= lmax) deliver(dst, t, ti);
          __syncwarp();   // the parked sums are consumed before the next block overwrites them
        }
      }
      __syncwarp();   // everyone is done with this tile before the next iteration's copy overwrites it
    }
    if (PIPE && pend) {   // the last block of the chunk
      const double *col = red + (pbuf ^ 1) * (16 * 33) + (lane & 15) * 33 + (lane >> 4) * 16;
      double t0 = col[0], t1 = col[1], t2 = col[2], t3 = col[3];
#pragma unroll
      for (int k = 4; k < 16; k += 4) { t0 += col[k]; t1 += col[k + 1]; t2 += col[k + 2]; t3 += col[k + 3]; }
      double t = (t0 + t1) + (t2 + t3);
      t += __shfl_xor_sync(0xffffffffu, t, 16);
      const double ti = __shfl_down_sync(0xffffffffu, t, KB);
      if (pok) deliver(pdst, t, ti);
    }
    cp_async_wait<0>();
    __syncwarp();   // (also: every lane has read its seeds before the next chunk's copies overwrite them)
  }
  if (first) {   // no ring of this warp is active for this m: the row is all zeros
    for (int l = m + lane; l <= lmax; l += 32)
#pragma unroll
      for (int s = 0; s < NS; ++s) out[s * part_shell + l - m] = make_double2(0.0, 0.0);
  }
}

// sum the partial rows in a fixed order, undo the recurrence scaling (lambda = c mu) and apply the Poisson filter
// alm *= -1/(l(l+1)), a_00 = 0                                        [shtpoissonsolve.c:526-550]
// blockIdx.z = shell of a batched pass (partial rows at + z * part_shell, alm at + z * alm_total)
__global__ void alm_finish_kernel(const double2 *__restrict__ part, long part_shell, int nrows, long alm_total,
                                  const int *__restrict__ m_loc, int nm_loc, const long *__restrict__ alm_off,
                                  const long *__restrict__ row_off, const double *__restrict__ ctab, int lmax,
                                  int apply_filter, double *__restrict__ alm_re, double *__restrict__ alm_im)
{
  const int mi = blockIdx.y;
  const int m = m_loc[mi];
  part += (size_t)blockIdx.z * part_shell;
  alm_re += (size_t)blockIdx.z * alm_total; alm_im += (size_t)blockIdx.z * alm_total;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= lmax - m; k += gridDim.x * blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int c = 0; c < nrows; ++c) {
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
// per-plane coefficient records, one per (m, l), l = m .. lmax+1, with the recurrence scaling folded in; rows are
// zero-padded so whole tiles can be processed without bounds checks.  One shell: 8 doubles {A_l, c P, c D, c K, 0}.
// Two shells (batched pass): 16 doubles {A_l, c P0, c D0, c K0, c P1, c D1, c K1, 0, 0, 0} (P, D, K complex).
// blockIdx.z = shell.
__global__ void synthesis_coef_kernel(const double *__restrict__ alm_re, const double *__restrict__ alm_im, long alm_total,
                                      const int *__restrict__ m_loc, const long *__restrict__ alm_off,
                                      const long *__restrict__ row_off, const double *__restrict__ Atab,
                                      const double *__restrict__ ctab, int lmax, int nshell, double *__restrict__ coef)
{
  const int mi = blockIdx.y, sh = blockIdx.z;
  const int m = m_loc[mi];
  const long len = row_off[mi + 1] - row_off[mi];
  const double *ar = alm_re + (size_t)sh * alm_total + alm_off[mi], *ai = alm_im + (size_t)sh * alm_total + alm_off[mi];
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
    if (nshell == 1) {
      double4 *o = reinterpret_cast<double4 *>(coef + 8 * (row_off[mi] + k));
      o[0] = make_double4(rec[0], rec[1], rec[2], rec[3]);
      o[1] = make_double4(rec[4], rec[5], rec[6], rec[7]);
    } else {
      double *o = coef + 16 * (row_off[mi] + k);
      if (sh == 0) {
        reinterpret_cast<double4 *>(o)[0] = make_double4(rec[0], rec[1], rec[2], rec[3]);
        o[4] = rec[4]; o[5] = rec[5]; o[6] = rec[6];
      } else {
        o[7] = rec[1];
        reinterpret_cast<double4 *>(o)[2] = make_double4(rec[2], rec[3], rec[4], rec[5]);
        reinterpret_cast<double4 *>(o)[3] = make_double4(rec[6], 0.0, 0.0, 0.0);
      }
    }
  }
}

// One warp = 32*R adjacent ring pairs of one m, independent of the other warps of its CTA (no block barrier): it
// streams the per-degree coefficient records of its own degree range through a private double-buffered shared-memory
// tile (cp.async, 1 KB per 16 degrees and shell) and reads them back as broadcast LDS.128.  Per degree and ring pair:
// the recurrence (DMUL + DFMA, shared by the NS shells of a batched pass) and six DFMA per shell.
template <int R, int NS>
__global__ void __launch_bounds__(kLegThreads, (NS == 2) ? ((R <= 2) ? 3 : 2) : ((R >= 4) ? 3 : 4))
legendre_synthesis_kernel(const double *__restrict__ coef, const long *__restrict__ row_off,
                          const int *__restrict__ ls_tab, const double2 *__restrict__ seed_tab,
                          const double *__restrict__ cth_rp, const double *__restrict__ sth_rp,
                          const int *__restrict__ m_loc, const long *__restrict__ b_off,
                          long b_fs, double2 *__restrict__ b_send, long b_shell,
                          double2 *const *__restrict__ rp_bptr, int nrp, int lmax)
{
  constexpr int RW = 8 * NS;     // doubles per coefficient record
  __shared__ __align__(16) double s_tile[kLegWarps][2][kLB * RW];
  __shared__ __align__(16) double2 s_seed[kLegWarps][R][32];   // the warp's seeds, copied in with the first coefficient tile
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c = blockIdx.x, mi = blockIdx.y;
  const int m = m_loc[mi];
  int ls[R];
  double mp[R], mc[R], x[R];
  double acc[NS][R][12];   // [parity(0 even,1 odd)*6 + {Pre,Pim,Dre,Dim,Kre,Kim}]
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
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int k = 0; k < 12; ++k) acc[s][j][k] = 0.0;
    lsmin = min(lsmin, ls[j]);
  }
  for (int o = 16; o; o >>= 1) lsmin = min(lsmin, __shfl_xor_sync(0xffffffffu, lsmin, o));
  if (lsmin != kNoStart) {
    // tiles start at lsmin (a multiple of kLB above m) and run through degree lmax+1; rows are zero padded (kRowPad)
    const double *crow = coef + RW * (row_off[mi] + (long)(lsmin - m));
    double *tile = &s_tile[w][0][0];
    // lane moves bytes [32*NS*lane, 32*NS*(lane+1)) of the tile
#pragma unroll
    for (int q = 0; q < 2 * NS; ++q) cp_async16(tile + 4 * NS * lane + 2 * q, crow + 4 * NS * lane + 2 * q);
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (rp0 + j * 32 < nrp) cp_async16(&s_seed[w][j][lane], &seed_tab[(size_t)mi * nrp + rp0 + j * 32]);
    cp_async_commit();
    int buf = 0;
    for (int l0 = lsmin; l0 <= lmax + 1; l0 += kLB, buf ^= 1) {
      crow += kLB * RW;
      double *nxt = &s_tile[w][buf ^ 1][0];
#pragma unroll
      for (int q = 0; q < 2 * NS; ++q) cp_async16(nxt + 4 * NS * lane + 2 * q, crow + 4 * NS * lane + 2 * q);
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (ls[j] == l0) {
          const double2 sd = s_seed[w][j][lane];
          mp[j] = sd.x; mc[j] = sd.y;
        }
      const double *t = &s_tile[w][buf][0];
#pragma unroll
      for (int i = 0; i < kLB; ++i) {
        const int par = (i & 1) * 6;   // (l0 - m) is even, so the parity of l+m is the parity of i
        if constexpr (NS == 1) {
          const double4 ra = *reinterpret_cast<const double4 *>(&t[i * 8]);
          const double4 rb = *reinterpret_cast<const double4 *>(&t[i * 8 + 4]);
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const double mu = mc[j];
            acc[0][j][par + 0] = fma(mu, ra.y, acc[0][j][par + 0]);
            acc[0][j][par + 1] = fma(mu, ra.z, acc[0][j][par + 1]);
            acc[0][j][par + 2] = fma(mu, ra.w, acc[0][j][par + 2]);
            acc[0][j][par + 3] = fma(mu, rb.x, acc[0][j][par + 3]);
            acc[0][j][par + 4] = fma(mu, rb.y, acc[0][j][par + 4]);
            acc[0][j][par + 5] = fma(mu, rb.z, acc[0][j][par + 5]);
            const double mn = fma(x[j] * ra.x, mu, -mp[j]);
            mp[j] = mu; mc[j] = mn;
          }
        } else {
          // 14 coefficients per degree, read two at a time so that few of them are alive at once (the 24 R accumulators
          // leave little room): {A, P0r} {P0i, D0r} {D0i, K0r} {K0i, P1r} {P1i, D1r} {D1i, K1r} {K1i, -}
          const double2 *t2 = reinterpret_cast<const double2 *>(&t[i * 16]);
          double2 q;
          q = t2[0];
          double xa[R];
#pragma unroll
          for (int j = 0; j < R; ++j) { xa[j] = x[j] * q.x; acc[0][j][par + 0] = fma(mc[j], q.y, acc[0][j][par + 0]); }
          q = t2[1];
#pragma unroll
          for (int j = 0; j < R; ++j) { acc[0][j][par + 1] = fma(mc[j], q.x, acc[0][j][par + 1]); acc[0][j][par + 2] = fma(mc[j], q.y, acc[0][j][par + 2]); }
          q = t2[2];
#pragma unroll
          for (int j = 0; j < R; ++j) { acc[0][j][par + 3] = fma(mc[j], q.x, acc[0][j][par + 3]); acc[0][j][par + 4] = fma(mc[j], q.y, acc[0][j][par + 4]); }
          q = t2[3];
#pragma unroll
          for (int j = 0; j < R; ++j) { acc[0][j][par + 5] = fma(mc[j], q.x, acc[0][j][par + 5]); acc[NS - 1][j][par + 0] = fma(mc[j], q.y, acc[NS - 1][j][par + 0]); }
          q = t2[4];
#pragma unroll
          for (int j = 0; j < R; ++j) { acc[NS - 1][j][par + 1] = fma(mc[j], q.x, acc[NS - 1][j][par + 1]); acc[NS - 1][j][par + 2] = fma(mc[j], q.y, acc[NS - 1][j][par + 2]); }
          q = t2[5];
#pragma unroll
          for (int j = 0; j < R; ++j) { acc[NS - 1][j][par + 3] = fma(mc[j], q.x, acc[NS - 1][j][par + 3]); acc[NS - 1][j][par + 4] = fma(mc[j], q.y, acc[NS - 1][j][par + 4]); }
          q = t2[6];
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const double mu = mc[j];
            acc[NS - 1][j][par + 5] = fma(mu, q.x, acc[NS - 1][j][par + 5]);
            const double mn = fma(xa[j], mu, -mp[j]);
            mp[j] = mu; mc[j] = mn;
          }
        }
      }
      __syncwarp();   // everyone is done with this tile before the next iteration's copy overwrites it
    }
    cp_async_wait<0>();
  }
  // combine parities into north / south and form the six fields
  const double dm = (double)m, m2 = dm * dm;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int rp = rp0 + j * 32;
    if (rp >= nrp) continue;
    const double sth = sth_rp[rp], cth = cth_rp[rp];
    const double isth = 1.0 / sth, cot = cth * isth, m2s2 = m2 * isth * isth;
    // destination: this rank's send buffer, or (fused exchange) the ring owner's receive buffer over NVLink
    // (blocks are ordered [ring pair][field][m][hemisphere]: field stride = 2 * number of local m)
    double2 *o = (rp_bptr ? rp_bptr[rp + s * nrp] : b_send + s * b_shell + b_off[rp]) + 2L * mi;
    const long fs = b_fs;
    double2 q[6][2];   // [field][hemisphere]
#pragma unroll
    for (int hemi = 0; hemi < 2; ++hemi) {
      const double sg = hemi ? -1.0 : 1.0;
      const double Pr = acc[s][j][0] + sg * acc[s][j][6], Pi = acc[s][j][1] + sg * acc[s][j][7];
      const double Dr = acc[s][j][2] + sg * acc[s][j][8], Di = acc[s][j][3] + sg * acc[s][j][9];
      const double Kr = acc[s][j][4] + sg * acc[s][j][10], Ki = acc[s][j][5] + sg * acc[s][j][11];
      const double q1r = Dr * isth, q1i = Di * isth;
      const double q3r = -sg * cot * q1r - Kr + m2s2 * Pr, q3i = -sg * cot * q1i - Ki + m2s2 * Pi;
      q[0][hemi] = make_double2(Pr, Pi);                       // phi
      q[1][hemi] = make_double2(q1r, q1i);                     // d_theta
      q[2][hemi] = make_double2(-dm * Pi, dm * Pr);            // i m P          (still to be divided by sin)
      q[3][hemi] = make_double2(q3r, q3i);                     // d_theta^2
      q[4][hemi] = make_double2(-dm * q1i, dm * q1r);          // i m d_theta    (.. / sin)
      q[5][hemi] = make_double2(-m2 * Pr, -m2 * Pi);           // -m^2 P         (.. / sin^2)
    }
    // the (north, south) pair of a field is 32 contiguous, 32-byte aligned bytes: one 256-bit store each (a full sector;
    // the neighbouring m of the same ring arrive from the CTAs next door and the L2 merges them into full lines)
#pragma unroll
    for (int f = 0; f < 6; ++f)
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o + f * fs), "d"(q[f][0].x), "d"(q[f][0].y),
                   "d"(q[f][1].x), "d"(q[f][1].y) : "memory");
  }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
int g_leg_warps_per_cta = 4;      // warps are independent in both Legendre kernels; clb_set_tuning(3, 1|2|4)
int g_syn_rings_per_thread = 4;   // tunable through clb_set_tuning(0, .)
int g_ana_rings_per_thread = 8;   // tunable through clb_set_tuning(1, .): 8, 4, 2 or 1
int g_ana_rows = 0;               // clb_set_tuning(5, n): partial-sum rows per m (0 = automatic)
int g_syn2_rings_per_thread = 3;  // clb_set_tuning(9, .): rings per thread of the two-shell synthesis kernel (3: 244 registers, no spills)
int g_ana2_rings_per_thread = 8;  // clb_set_tuning(10, .): rings per thread of the two-shell analysis kernel

static int cur_device()
{
  int dev = 0;
  CLB_CUDA_CHECK(cudaGetDevice(&dev));
  return dev;
}

int g_ana_pipeline = 1;           // clb_set_tuning(12, 0|1|2): warp sum of a block overlapped with the next block's FP64 work: never,
                                  // in one-shell passes only (measured: 71.1 -> 70.1 ms there, 56.7 -> 61.0 ms with two shells at the
                                  // 255-register cap), always

template <int R, int NS, int NB>
static void launch_ana_t(const ShtPlan *p, const double2 *g_recv, const double2 *const *gsrc, int rows, int nchunk, cudaStream_t st)
{
  const int warps = g_leg_warps_per_cta;
  dim3 grid((rows + warps - 1) / warps, p->nm_loc);
#define CLB_ANA_ARGS g_recv, p->g_recv_total, gsrc, p->d_g_off, p->d_g_stride, p->d_A, p->d_row_off, p->d_ls_ana, p->d_seed, p->d_cth, \
      p->d_m_loc, p->d_alm_off, reinterpret_cast<double2 *>(p->d_part), (long)rows * p->alm_total, p->alm_total, p->nrp,           \
      (int)p->lmax, rows, nchunk
  // dynamic shared memory per warp (see the kernel): A tiles, parked partial sums (two buffers when pipelined), seeds
  auto warp_bytes = [](int bufs) { return sizeof(double) * (size_t)(2 * kAnaTile + bufs * 16 * 33 + 2 * R * 32); };
  if (g_ana_pipeline == 2 || (g_ana_pipeline == 1 && NS == 1)) {
    static bool done[64] = {};   // (per instantiation and device: the attribute belongs to the function in one context)
    bool &attr = done[cur_device() & 63];
    if (!attr) {
      CLB_CUDA_CHECK(cudaFuncSetAttribute(legendre_analysis_kernel<R, NS, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(warp_bytes(2) * kLegWarps)));
      attr = true;
    }
    legendre_analysis_kernel<R, NS, NB, true><<<grid, 32 * warps, warp_bytes(2) * warps, st>>>(CLB_ANA_ARGS);
  } else {
    static bool done[64] = {};
    bool &attr = done[cur_device() & 63];
    if (!attr) {
      CLB_CUDA_CHECK(cudaFuncSetAttribute(legendre_analysis_kernel<R, NS, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(warp_bytes(1) * kLegWarps)));
      attr = true;
    }
    legendre_analysis_kernel<R, NS, NB, false><<<grid, 32 * warps, warp_bytes(1) * warps, st>>>(CLB_ANA_ARGS);
  }
#undef CLB_ANA_ARGS
}

// nshell = 1: one plane.  nshell = 2: two planes in one pass: shell s reads g at
// g_recv + s * g_recv_total (or the peers' second buffers) and delivers alm at alm_* + s * alm_total.
int launch_legendre_analysis(ShtPlan *p, const double2 *d_g_recv, double *d_alm_re, double *d_alm_im, int apply_filter,
                             cudaStream_t st, int nshell)
{
  if (p->nm_loc == 0) return 0;
  if (nshell != 1 && nshell != 2) { fprintf(stderr, "calclens_b200: legendre analysis of %d shells per pass is not supported\n", nshell); abort(); }
  // rings per thread: the tuned value, stepped down through the instantiated set while most of a warp would be left
  // without rings (small maps); the chunk count is derived from the final R only
  static const int kAnaR[] = {8, 6, 4, 2, 1};
  const int want = nshell == 2 ? g_ana2_rings_per_thread : g_ana_rings_per_thread;
  int ri = 0;
  while (kAnaR[ri] != want) {
    if (++ri >= 5) { fprintf(stderr, "calclens_b200: analysis rings per thread %d has no instantiation\n", want); abort(); }
  }
  while (kAnaR[ri] > 1 && p->nrp < 32 * kAnaR[ri]) ++ri;
  const int R = kAnaR[ri];
  const int nchunk = (p->nrp + 32 * R - 1) / (32 * R);
  // partial-sum rows per m: a warp walks nchunk / rows chunks and owns one row.  Few rows keep the partial sums small
  // (rows x alm); enough of them keep the grid at >= ~16 waves of CTAs so that the tail of the launch stays short
  int rows = g_ana_rows;
  if (rows <= 0) {
    const long want_warps = 16L * 3 * sm_count() * g_leg_warps_per_cta;
    rows = (int)std::min<long>(nchunk, std::max<long>(16, (want_warps + p->nm_loc - 1) / p->nm_loc));
    // ... and no more than ~24 GB of partial sums (lmax = 3 Nside - 1 at Nside 4096 would take 39 GB with 16 rows and two shells)
    const double row_gb = 16e-9 * (double)p->alm_total * nshell;
    while (rows > 8 && rows * row_gb > 24.0) rows = (rows + 1) / 2;
  }
  rows = std::max(1, std::min(rows, nchunk));
  rows = (nchunk + (nchunk / rows) - 1) / (nchunk / rows);     // chunks per warp = floor(nchunk / rows); rows = what that needs
  if (!p->d_part || p->ana_nchunk < rows * nshell) {   // grow only: one- and two-shell passes alternate in a run of planes
    if (p->d_part) cudaFree(p->d_part);
    CLB_CUDA_CHECK(cudaMalloc(&p->d_part, sizeof(double2) * (size_t)rows * nshell * p->alm_total));
    p->ana_nchunk = rows * nshell;
  }
  // fused exchange (clb_sht_plan_set_peers): g is pulled from the ring owners' send buffers only when the caller passes
  // no receive buffer; a caller that hands in g_recv gets the plain local path
  const double2 *const *gsrc = d_g_recv ? nullptr : p->d_rp_gsrc;
  if (!d_g_recv && !gsrc) { fprintf(stderr, "calclens_b200: legendre analysis needs g_recv (no peer buffers are set)\n"); abort(); }
  if (gsrc && nshell > p->peer_shells) { fprintf(stderr, "calclens_b200: the peer buffers of this plan hold one shell only\n"); abort(); }
  if (nshell == 1) {
    switch (R) {
      case 8: launch_ana_t<8, 1, 3>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      case 6: launch_ana_t<6, 1, 3>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      case 4: launch_ana_t<4, 1, 3>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      case 2: launch_ana_t<2, 1, 4>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      default: launch_ana_t<1, 1, 4>(p, d_g_recv, gsrc, rows, nchunk, st); break;
    }
  } else {
    switch (R) {
      case 8: launch_ana_t<8, 2, 2>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      case 6: launch_ana_t<6, 2, 2>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      case 4: launch_ana_t<4, 2, 3>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      case 2: launch_ana_t<2, 2, 4>(p, d_g_recv, gsrc, rows, nchunk, st); break;
      default: launch_ana_t<1, 2, 4>(p, d_g_recv, gsrc, rows, nchunk, st); break;
    }
  }
  CLB_CUDA_CHECK(cudaGetLastError());
  dim3 grid((unsigned)((p->lmax + 256) / 256), p->nm_loc, nshell);
  alm_finish_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const double2 *>(p->d_part), (long)rows * p->alm_total, rows,
                                          p->alm_total, p->d_m_loc, p->nm_loc, p->d_alm_off, p->d_row_off, p->d_c,
                                          (int)p->lmax, apply_filter, d_alm_re, d_alm_im);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 2;
}

int launch_legendre_synthesis(ShtPlan *p, const double *d_alm_re, const double *d_alm_im, double2 *d_b_send,
                              cudaStream_t st, int nshell)
{
  if (p->nm_loc == 0) return 0;
  if (nshell != 1 && nshell != 2) { fprintf(stderr, "calclens_b200: legendre synthesis of %d shells per pass is not supported\n", nshell); abort(); }
  if (p->coef_shells < nshell) {    // the two-shell records are twice as wide: grow the buffer on first use
    if (p->d_coef) cudaFree(p->d_coef);
    CLB_CUDA_CHECK(cudaMalloc(&p->d_coef, sizeof(double) * 8 * nshell * std::max<long>(p->rows_total, 1)));
    p->coef_shells = nshell;
  }
  dim3 cgrid((unsigned)((p->lmax + 2 + kRowPad + 32 + 255) / 256), p->nm_loc, nshell);
  synthesis_coef_kernel<<<cgrid, 256, 0, st>>>(d_alm_re, d_alm_im, p->alm_total, p->d_m_loc, p->d_alm_off, p->d_row_off, p->d_A,
                                               p->d_c, (int)p->lmax, nshell, p->d_coef);
  CLB_CUDA_CHECK(cudaGetLastError());
  int R = nshell == 2 ? g_syn2_rings_per_thread : g_syn_rings_per_thread;
  while (R > 1 && p->nrp < 32 * R) --R;   // 4, 3, 2, 1 are all instantiated
  // fused exchange: b is pushed into the ring owners' receive buffers only when the caller passes no send buffer
  double2 *const *bptr = d_b_send ? nullptr : p->d_rp_bptr;
  if (!d_b_send && !bptr) { fprintf(stderr, "calclens_b200: legendre synthesis needs b_send (no peer buffers are set)\n"); abort(); }
  if (bptr && nshell > p->peer_shells) { fprintf(stderr, "calclens_b200: the peer buffers of this plan hold one shell only\n"); abort(); }
  const int warps = g_leg_warps_per_cta;
  const int nchunk = (p->nrp + R * 32 * warps - 1) / (R * 32 * warps);
  dim3 grid(nchunk, p->nm_loc);
#define CLB_SYN_LAUNCH(RR, NS)                                                                                          \
  legendre_synthesis_kernel<RR, NS><<<grid, 32 * warps, 0, st>>>(p->d_coef, p->d_row_off, p->d_ls_syn, p->d_seed, p->d_cth, \
                                                                  p->d_sth, p->d_m_loc, p->d_b_off, 2L * p->nm_loc, d_b_send, \
                                                                  p->b_send_total, bptr, p->nrp, (int)p->lmax)
  if (nshell == 1) {
    switch (R) {
      case 4: CLB_SYN_LAUNCH(4, 1); break;
      case 3: CLB_SYN_LAUNCH(3, 1); break;
      case 2: CLB_SYN_LAUNCH(2, 1); break;
      default: CLB_SYN_LAUNCH(1, 1); break;
    }
  } else {
    switch (R) {
      case 4: CLB_SYN_LAUNCH(4, 2); break;
      case 3: CLB_SYN_LAUNCH(3, 2); break;
      case 2: CLB_SYN_LAUNCH(2, 2); break;
      default: CLB_SYN_LAUNCH(1, 2); break;
    }
  }
#undef CLB_SYN_LAUNCH
  CLB_CUDA_CHECK(cudaGetLastError());
  return 2;
}

}  // namespace clb
