// calclens_b200/csrc/legendre.cu
// FP64 Legendre transforms for sm_100a with lambda_lm generated on the fly (no P_lm table in HBM).
//   analysis  : a_lm = sum_rings lambda_lm(theta_r) [g_m^N(r) + (-1)^(l+m) g_m^S(r)], then the Poisson filter
//               [map2alm_transpose_mpi.c:427-536; shtpoissonsolve.c:526-550]
//   synthesis : the six ring-space fields phi, d_theta, d_phi, d_theta^2, d_theta d_phi, d_phi^2 per (m, ring)
//               [alm2allmaps_transpose_mpi.c:272-595]
//
// Both kernels put one ring pair (north ring + its southern mirror, combined through the parity of l+m) on a
// thread and walk l with the two-instruction scaled recurrence mu_{l+1} = (x A_l) mu_l - mu_{l-1} (sht_plan.cu),
// started from the precomputed per-(m, ring) seeds, so no work is spent below the degree where lambda reaches
// 1e-30.  All threads of a CTA share m, so recurrence coefficients and alm-derived coefficients are block-uniform.
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
constexpr int kAnaThreads = 256;
constexpr int kSynThreads = 256;
constexpr int kSynTile = 32;       // degrees per shared-memory coefficient tile
static_assert(kLB == kSeedAlign, "l-blocks must line up with the seed alignment");

// ring pair handled by (chunk c, warp w, slot j, lane): warp-sized groups of adjacent rings are dealt round-robin
// to chunks so that every CTA sees all latitudes (balanced start degrees)
__device__ __forceinline__ int ring_of(int c, int nchunk, int w, int nwarp, int j, int lane)
{
  return (((j * nwarp + w) * nchunk + c) << 5) + lane;
}

// ---------------------------------------------------------------------------------------------------------------
// analysis
// ---------------------------------------------------------------------------------------------------------------
template <int RPT>
__global__ void __launch_bounds__(kAnaThreads)
legendre_analysis_kernel(const double2 *__restrict__ g_recv, const long *__restrict__ g_off,
                         const int *__restrict__ g_stride, const double *__restrict__ Atab,
                         const long *__restrict__ row_off, const int *__restrict__ ls_tab,
                         const double2 *__restrict__ seed_tab, const double *__restrict__ cth_rp,
                         const int *__restrict__ m_loc, const long *__restrict__ alm_off, double2 *__restrict__ part,
                         long alm_total, int nrp, int lmax, int nchunk)
{
  extern __shared__ unsigned char smem_raw[];
  double2 *s_state = reinterpret_cast<double2 *>(smem_raw);                 // [RPT][T]
  double2 *s_gp = s_state + RPT * kAnaThreads;                                // [RPT][T] G+ = gN + gS
  double2 *s_gm = s_gp + RPT * kAnaThreads;                                   // [RPT][T] G- = gN - gS
  double *s_x = reinterpret_cast<double *>(s_gm + RPT * kAnaThreads);         // [RPT][T]
  int *s_ls = reinterpret_cast<int *>(s_x + RPT * kAnaThreads);               // [RPT][T]
  __shared__ double s_red[kAnaThreads / 32][32];
  __shared__ int s_lsmin;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nwarp = kAnaThreads / 32;
  const int c = blockIdx.x, mi = blockIdx.y;
  const int m = m_loc[mi];
  if (tid == 0) s_lsmin = kNoStart;
  __syncthreads();
  int lsmin = kNoStart;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int rp = ring_of(c, nchunk, w, nwarp, j, lane);
    int ls = kNoStart;
    double2 st = make_double2(0.0, 0.0), gp = st, gm = st;
    double x = 0.0;
    if (rp < nrp) {
      ls = ls_tab[(size_t)mi * nrp + rp];
      if (ls != kNoStart) {
        st = seed_tab[(size_t)mi * nrp + rp];
        const double2 *g = g_recv + g_off[rp] + (long)mi * g_stride[rp];
        double2 gn = g[0], gs = g[1];
        gp = make_double2(gn.x + gs.x, gn.y + gs.y);
        gm = make_double2(gn.x - gs.x, gn.y - gs.y);
        x = cth_rp[rp];
      }
    }
    s_state[j * kAnaThreads + tid] = st; s_gp[j * kAnaThreads + tid] = gp; s_gm[j * kAnaThreads + tid] = gm;
    s_x[j * kAnaThreads + tid] = x; s_ls[j * kAnaThreads + tid] = ls;
    lsmin = min(lsmin, ls);
  }
  for (int o = 16; o; o >>= 1) lsmin = min(lsmin, __shfl_xor_sync(0xffffffffu, lsmin, o));
  if (lane == 0) atomicMin(&s_lsmin, lsmin);
  __syncthreads();
  lsmin = s_lsmin;
  double2 *out = part + (size_t)c * alm_total + alm_off[mi];   // index l - m
  // degrees below the first active block get exact zeros
  {
    const int lz = (lsmin == kNoStart) ? lmax + 1 : lsmin;
    for (int l = m + tid; l < lz; l += kAnaThreads) out[l - m] = make_double2(0.0, 0.0);
    if (lsmin == kNoStart) return;
  }
  const double *Arow = Atab + row_off[mi];
  for (int l0 = lsmin; l0 <= lmax; l0 += kLB) {
    double A[kLB];
#pragma unroll
    for (int i = 0; i < kLB; ++i) A[i] = __ldg(&Arow[l0 - m + i]);
    double v[2 * kLB];   // v[i] = re(l0+i), v[kLB+i] = im(l0+i)
#pragma unroll
    for (int i = 0; i < 2 * kLB; ++i) v[i] = 0.0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (l0 >= s_ls[j * kAnaThreads + tid]) {
        double2 st = s_state[j * kAnaThreads + tid];
        const double2 gp = s_gp[j * kAnaThreads + tid], gm = s_gm[j * kAnaThreads + tid];
        const double x = s_x[j * kAnaThreads + tid];
        double mp = st.x, mc = st.y;
#pragma unroll
        for (int i = 0; i < kLB; i += 2) {
          v[i] = fma(mc, gp.x, v[i]); v[kLB + i] = fma(mc, gp.y, v[kLB + i]);
          double mn = fma(x * A[i], mc, -mp); mp = mc; mc = mn;
          v[i + 1] = fma(mc, gm.x, v[i + 1]); v[kLB + i + 1] = fma(mc, gm.y, v[kLB + i + 1]);
          mn = fma(x * A[i + 1], mc, -mp); mp = mc; mc = mn;
        }
        s_state[j * kAnaThreads + tid] = make_double2(mp, mc);
      }
    }
    // warp transpose-reduce: afterwards lane L holds the warp total of v[L]
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool upper = (lane & s) != 0;
#pragma unroll
      for (int k = 0; k < s; ++k) {
        double send = upper ? v[k] : v[k + s];
        double keep = upper ? v[k + s] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    s_red[w][lane] = v[0];
    __syncthreads();
    if (w == 0) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < kAnaThreads / 32; ++k) t += s_red[k][lane];   // fixed order: deterministic
      double ti = __shfl_down_sync(0xffffffffu, t, kLB);                 // imaginary part lives kLB lanes up
      if (lane < kLB && l0 + lane <= lmax) out[l0 - m + lane] = make_double2(t, ti);
    }
    __syncthreads();
  }
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

template <int R>
__global__ void __launch_bounds__(kSynThreads)
legendre_synthesis_kernel(const double *__restrict__ coef, const long *__restrict__ row_off,
                          const int *__restrict__ ls_tab, const double2 *__restrict__ seed_tab,
                          const double *__restrict__ cth_rp, const double *__restrict__ sth_rp,
                          const int *__restrict__ m_loc, const long *__restrict__ b_off,
                          const int *__restrict__ b_stride, double2 *__restrict__ b_send, int nrp, int lmax, int nchunk)
{
  __shared__ __align__(16) double s_tile[kSynTile * 8];
  __shared__ int s_lsmin;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nwarp = kSynThreads / 32;
  const int c = blockIdx.x, mi = blockIdx.y;
  const int m = m_loc[mi];
  if (tid == 0) s_lsmin = kNoStart;
  __syncthreads();
  int ls[R], rpv[R];
  double mp[R], mc[R], x[R];
  double acc[R][12];   // [parity(0 even,1 odd)*6 + {Pre,Pim,Dre,Dim,Kre,Kim}]
  int lsmin = kNoStart;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int rp = ring_of(c, nchunk, w, nwarp, j, lane);
    rpv[j] = rp; ls[j] = kNoStart; mp[j] = 0.0; mc[j] = 0.0; x[j] = 0.0;
    if (rp < nrp) {
      ls[j] = ls_tab[(size_t)mi * nrp + rp];
      if (ls[j] != kNoStart) {
        double2 st = seed_tab[(size_t)mi * nrp + rp];
        mp[j] = st.x; mc[j] = st.y; x[j] = cth_rp[rp];
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[j][k] = 0.0;
    lsmin = min(lsmin, ls[j]);
  }
  for (int o = 16; o; o >>= 1) lsmin = min(lsmin, __shfl_xor_sync(0xffffffffu, lsmin, o));
  if (lane == 0) atomicMin(&s_lsmin, lsmin);
  __syncthreads();
  lsmin = s_lsmin;
  if (lsmin != kNoStart) {
    const double *crow = coef + 8 * row_off[mi];
    // tiles start at lsmin (a multiple of kLB above m) and run through degree lmax+1
    const int per_thread = (kSynTile * 8) / kSynThreads;   // doubles of the tile each thread moves
    static_assert((kSynTile * 8) % kSynThreads == 0 && per_thread == 1, "tile copy assumes one double per thread");
    double nxt = __ldg(&crow[8 * (long)(lsmin - m) + tid]);
    for (int l0 = lsmin; l0 <= lmax + 1; l0 += kSynTile) {
      __syncthreads();
      s_tile[tid] = nxt;
      __syncthreads();
      if (l0 + kSynTile <= lmax + 1) nxt = __ldg(&crow[8 * (long)(l0 + kSynTile - m) + tid]);
#pragma unroll
      for (int h = 0; h < kSynTile / kLB; ++h) {
        const int lh = l0 + h * kLB;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (lh >= ls[j]) {
#pragma unroll
            for (int i = 0; i < kLB; ++i) {
              const double4 ra = *reinterpret_cast<const double4 *>(&s_tile[(h * kLB + i) * 8]);
              const double4 rb = *reinterpret_cast<const double4 *>(&s_tile[(h * kLB + i) * 8 + 4]);
              const int par = (i & 1) * 6;   // (lh - m) is even, so parity of l+m is parity of i
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
      }
    }
  }
  // combine parities into north / south and form the six fields
  const double dm = (double)m, m2 = dm * dm;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int rp = rpv[j];
    if (rp >= nrp) continue;
    const double sth = sth_rp[rp], cth = cth_rp[rp];
    const double isth = 1.0 / sth, cot = cth * isth, m2s2 = m2 * isth * isth;
    double2 *o = b_send + b_off[rp] + (long)mi * 6 * b_stride[rp];
    const long fs = b_stride[rp];
#pragma unroll
    for (int hemi = 0; hemi < 2; ++hemi) {
      const double sg = hemi ? -1.0 : 1.0;
      const double Pr = acc[j][0] + sg * acc[j][6], Pi = acc[j][1] + sg * acc[j][7];
      const double Dr = acc[j][2] + sg * acc[j][8], Di = acc[j][3] + sg * acc[j][9];
      const double Kr = acc[j][4] + sg * acc[j][10], Ki = acc[j][5] + sg * acc[j][11];
      const double q1r = Dr * isth, q1i = Di * isth;
      const double q3r = -sg * cot * q1r - Kr + m2s2 * Pr, q3i = -sg * cot * q1i - Ki + m2s2 * Pi;
      o[0 * fs + hemi] = make_double2(Pr, Pi);                       // phi
      o[1 * fs + hemi] = make_double2(q1r, q1i);                     // d_theta
      o[2 * fs + hemi] = make_double2(-dm * Pi, dm * Pr);            // i m P          (still to be divided by sin)
      o[3 * fs + hemi] = make_double2(q3r, q3i);                     // d_theta^2
      o[4 * fs + hemi] = make_double2(-dm * q1i, dm * q1r);          // i m d_theta    (.. / sin)
      o[5 * fs + hemi] = make_double2(-m2 * Pr, -m2 * Pi);           // -m^2 P         (.. / sin^2)
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
static int pick_rpt(int nrp)
{
  int per = (nrp + kAnaThreads - 1) / kAnaThreads;
  if (per >= 8) return 8;
  if (per >= 4) return 4;
  if (per >= 2) return 2;
  return 1;
}

template <int RPT>
static void launch_ana_t(const ShtPlan *p, const double2 *g_recv, int nchunk, cudaStream_t st)
{
  const size_t smem = (size_t)RPT * kAnaThreads * (3 * sizeof(double2) + sizeof(double) + sizeof(int));
  static bool attr_set = false;
  if (!attr_set) {
    CLB_CUDA_CHECK(cudaFuncSetAttribute(legendre_analysis_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(nchunk, p->nm_loc);
  legendre_analysis_kernel<RPT><<<grid, kAnaThreads, smem, st>>>(
      g_recv, p->d_g_off, p->d_g_stride, p->d_A, p->d_row_off, p->d_ls_ana, p->d_seed, p->d_cth, p->d_m_loc,
      p->d_alm_off, reinterpret_cast<double2 *>(p->d_part), p->alm_total, p->nrp, (int)p->lmax, nchunk);
}

int launch_legendre_analysis(ShtPlan *p, const double2 *d_g_recv, double *d_alm_re, double *d_alm_im, int apply_filter,
                             cudaStream_t st)
{
  if (p->nm_loc == 0) return 0;
  const int rpt = pick_rpt(p->nrp);
  const int nchunk = (p->nrp + rpt * kAnaThreads - 1) / (rpt * kAnaThreads);
  if (!p->d_part || p->ana_nchunk != nchunk) {
    if (p->d_part) cudaFree(p->d_part);
    CLB_CUDA_CHECK(cudaMalloc(&p->d_part, sizeof(double2) * (size_t)nchunk * p->alm_total));
    p->ana_nchunk = nchunk;
  }
  switch (rpt) {
    case 8: launch_ana_t<8>(p, d_g_recv, nchunk, st); break;
    case 4: launch_ana_t<4>(p, d_g_recv, nchunk, st); break;
    case 2: launch_ana_t<2>(p, d_g_recv, nchunk, st); break;
    default: launch_ana_t<1>(p, d_g_recv, nchunk, st); break;
  }
  CLB_CUDA_CHECK(cudaGetLastError());
  dim3 grid((unsigned)((p->lmax + 256) / 256), p->nm_loc);
  alm_finish_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const double2 *>(p->d_part), nchunk, p->alm_total, p->d_m_loc,
                                          p->nm_loc, p->d_alm_off, p->d_row_off, p->d_c, (int)p->lmax, apply_filter,
                                          d_alm_re, d_alm_im);
  CLB_CUDA_CHECK(cudaGetLastError());
  return 2;
}

int g_syn_rings_per_thread = 2;   // tunable through clb_set_tuning

int launch_legendre_synthesis(ShtPlan *p, const double *d_alm_re, const double *d_alm_im, double2 *d_b_send,
                              cudaStream_t st)
{
  if (p->nm_loc == 0) return 0;
  dim3 cgrid((unsigned)((p->lmax + 2 + kRowPad + 32 + 255) / 256), p->nm_loc);
  synthesis_coef_kernel<<<cgrid, 256, 0, st>>>(d_alm_re, d_alm_im, p->d_m_loc, p->d_alm_off, p->d_row_off, p->d_A, p->d_c,
                                               (int)p->lmax, p->d_coef);
  CLB_CUDA_CHECK(cudaGetLastError());
  int R = g_syn_rings_per_thread;
  if (p->nrp <= kSynThreads) R = 1;
  const int nchunk = (p->nrp + R * kSynThreads - 1) / (R * kSynThreads);
  dim3 grid(nchunk, p->nm_loc);
#define CLB_SYN_LAUNCH(RR)                                                                                           \
  legendre_synthesis_kernel<RR><<<grid, kSynThreads, 0, st>>>(p->d_coef, p->d_row_off, p->d_ls_syn, p->d_seed, p->d_cth, \
                                                              p->d_sth, p->d_m_loc, p->d_b_off, p->d_b_stride, d_b_send, \
                                                              p->nrp, (int)p->lmax, nchunk)
  switch (R) {
    case 4: CLB_SYN_LAUNCH(4); break;
    case 2: CLB_SYN_LAUNCH(2); break;
    default: CLB_SYN_LAUNCH(1); break;
  }
#undef CLB_SYN_LAUNCH
  CLB_CUDA_CHECK(cudaGetLastError());
  return 2;
}

}  // namespace clb
