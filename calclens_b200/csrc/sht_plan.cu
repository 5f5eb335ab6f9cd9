// calclens_b200/csrc/sht_plan.cu  (compiled with -fmad=false: the seed kernel mirrors the reference's
// un-contracted recurrence arithmetic)
// Plan for the HEALPix SHT on one GPU of an N-GPU job: ring geometry, ring/m ownership and exchange layouts,
// Legendre recurrence tables and per-(m, ring) start seeds.
//   ring geometry           [healpix_utils.c:907-953 get_ring_info2, healpix_shtrans.c:54-160 healpixsht_plan]
//   quadrature weights      [map2alm_transpose_mpi.c:108-124]
//   lambda_lm recurrence    [healpix_plmgen.c:73-183 plmgen, :185-243 plmgen_init, :245-260 plmgen_recalc_recfac]
//   small-Ylm cut           [healpix_shtrans.c:533-544 get_lmin_ylm; called with double sin(theta) in
//                            map2alm_transpose_mpi.c:457 and with (float) sin(theta) in alm2allmaps_transpose_mpi.c:308]
#include "sht_internal.cuh"
#include "healpix.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>

namespace clb {

void fft_tables_create(ShtPlan *p);
void fft_tables_destroy(ShtPlan *p);

template <typename T>
static T *to_device(const std::vector<T> &v)
{
  T *d = nullptr;
  CLB_CUDA_CHECK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CLB_CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return d;
}

// ---------------------------------------------------------------------------------------------------------------
// recurrence tables.  Reference recurrence: lam_{l+1} = cth*lam_l*rf0[l] - lam_{l-1}*rf1[l],
// rf0[l] = t1fac[l]*t2fac[l+m]*t2fac[l-m], rf1[l] = rf0[l]/rf0[l-1].  With lam_l = c_l mu_l and
// c_{l+1} = rf1[l] c_{l-1} (c_m = 1, c_{m+1} = rf0[m]) it becomes mu_{l+1} = (cth*A_l) mu_l - mu_{l-1},
// A_l = rf0[l] c_l / c_{l+1}: two FP64 instructions per degree instead of three.
// ---------------------------------------------------------------------------------------------------------------
__global__ void recurrence_table_kernel(const int *__restrict__ m_loc, int nm_loc, int lmax,
                                        const long *__restrict__ row_off, const double *__restrict__ t1fac,
                                        const double *__restrict__ t2fac, double *__restrict__ A, double *__restrict__ c)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nm_loc) return;
  const int m = m_loc[i];
  double *Ar = A + row_off[i], *cr = c + row_off[i];
  const long len = row_off[i + 1] - row_off[i];
  double c_prev = 0.0, c_cur = 1.0, f_old = 1.0;
  long k = 0;
  for (int l = m; l <= lmax + 1; ++l, ++k) {
    double rf0 = t1fac[l] * t2fac[l + m] * t2fac[l - m];
    double rf1 = rf0 / f_old;
    f_old = rf0;
    double c_next = (l == m) ? rf0 : rf1 * c_prev;
    cr[k] = c_cur;
    Ar[k] = rf0 * c_cur / c_next;
    c_prev = c_cur; c_cur = c_next;
  }
  for (; k < len; ++k) { cr[k] = 0.0; Ar[k] = 0.0; }
}

// One thread per (local m, ring pair): the reference's scaled start-up loop until |lambda| exceeds 1e-30
// (healpix_plmgen.c:99-155), keeping the state at the last kSeedAlign-aligned degree at or below firstl.
__global__ void seed_kernel(const int *__restrict__ m_loc, int nm_loc, int nrp, int lmax,
                            const double *__restrict__ cth_rp, const double *__restrict__ sth_rp,
                            const double *__restrict__ logsth_rp, const double *__restrict__ mfac,
                            const double *__restrict__ t1fac, const double *__restrict__ t2fac,
                            const double *__restrict__ cf, const long *__restrict__ row_off,
                            const double *__restrict__ ctab, double inv_ln2, double ln2, double fbig, double fsmall,
                            int *__restrict__ ls_ana, int *__restrict__ ls_syn, double2 *__restrict__ seed)
{
  const int rp = blockIdx.x * blockDim.x + threadIdx.x;
  const int mi = blockIdx.y;
  if (rp >= nrp) return;
  const int m = m_loc[mi];
  const size_t o = (size_t)mi * nrp + rp;
  const double cth = cth_rp[rp], sth = sth_rp[rp];
  // get_lmin_ylm: (long)((m-40)/1.35/sintheta), double argument (analysis) and float-rounded argument (synthesis)
  long cutA = (long)((m - 40) / 1.35 / sth);
  long cutS = (long)((m - 40) / 1.35 / (double)((float)sth));
  const bool skipA = ((cutA > m ? cutA : m) > lmax), skipS = ((cutS > m ? cutS : m) > lmax);
  int ls = kNoStart;
  double2 sd = make_double2(0.0, 0.0);
  if (!(skipA && skipS) && !(m > 0 && sth == 0.0)) {
    const double eps = 1e-30;   // plmeps, map2alm_transpose_mpi.c:102; inv_ln2, ln2, fbig, fsmall come from host libm
    double logval = mfac[m];
    if (m > 0) logval += m * inv_ln2 * logsth_rp[rp];
    long scale = (long)((logval / 90) - (-4));
    double corfac = (scale < 0) ? 0.0 : cf[scale];
    double lam_prev = 0.0;
    double lam_cur = exp(ln2 * (logval - (scale + (-4)) * 90));
    if (m & 1) lam_cur = -lam_cur;
    double f_old = 1.0;
    int l = m;
    int bl = m; double b_prev = 0.0, b_cur = lam_cur * corfac;   // state at the last aligned degree
    int firstl = -1;
    while (true) {
      if (((l - m) % kSeedAlign) == 0) { bl = l; b_prev = lam_prev * corfac; b_cur = lam_cur * corfac; }
      if (fabs(lam_cur * corfac) > eps) { firstl = l; break; }
      if (l + 1 > lmax) break;
      double rf0 = t1fac[l] * t2fac[l + m] * t2fac[l - m];
      double rf1 = rf0 / f_old;
      f_old = rf0;
      double lam_next = cth * lam_cur * rf0 - lam_prev * rf1;
      lam_prev = lam_cur; lam_cur = lam_next; ++l;
      while (fabs(lam_cur) > fbig) {
        lam_prev *= fsmall; lam_cur *= fsmall; ++scale;
        corfac = (scale < 0) ? 0.0 : cf[scale];
      }
    }
    if (firstl >= 0) {
      if (b_cur == 0.0 && b_prev == 0.0) {
        // the aligned state underflowed the scaled representation: advance in true scale to the next aligned
        // degree above firstl (drops < kSeedAlign terms of magnitude ~1e-30)
        lam_prev *= corfac; lam_cur *= corfac;
        while (((l - m) % kSeedAlign) != 0 && l <= lmax) {
          double rf0 = t1fac[l] * t2fac[l + m] * t2fac[l - m];
          double rf1 = rf0 / f_old;
          f_old = rf0;
          double lam_next = cth * lam_cur * rf0 - lam_prev * rf1;
          lam_prev = lam_cur; lam_cur = lam_next; ++l;
        }
        bl = l; b_prev = lam_prev; b_cur = lam_cur;
      }
      if (bl <= lmax) {
        ls = bl;
        const double *cr = ctab + row_off[mi];
        double cprev = (bl > m) ? cr[bl - m - 1] : 1.0;
        sd = make_double2(b_prev / cprev, b_cur / cr[bl - m]);
      }
    }
  }
  ls_ana[o] = skipA ? kNoStart : ls;
  ls_syn[o] = skipS ? kNoStart : ls;
  seed[o] = sd;
}

// ---------------------------------------------------------------------------------------------------------------


ShtPlan *sht_plan_create(long order, long lmax, const double *ring_weights, int nranks, int rank,
                         const int *rp_owner, const int *m_owner)
{
  ShtPlan *p = new ShtPlan();
  p->order = order; p->nside = 1L << order; p->npix = 12L << (2 * order); p->lmax = lmax;
  p->nranks = nranks; p->rank = rank;
  p->nrp = (int)(2 * p->nside);
  const int nrp = p->nrp;
  // geometry
  p->h_cth.resize(nrp); p->h_sth.resize(nrp); p->h_weight.resize(nrp); p->h_nphi.resize(nrp);
  p->h_shifted.resize(nrp); p->h_startN.resize(nrp); p->h_startS.resize(nrp);
  std::vector<double> logsth(nrp);
  const double quadweight = 4.0 * CLB_PI / p->npix;
  for (int rp = 0; rp < nrp; ++rp) {
    RingInfo ri = ring_info(rp + 1, order);
    p->h_cth[rp] = ri.costheta; p->h_sth[rp] = ri.sintheta; p->h_nphi[rp] = (int)ri.ringpix;
    p->h_shifted[rp] = (int)ri.shifted; p->h_startN[rp] = ri.startpix;
    p->h_startS[rp] = (rp == nrp - 1) ? -1 : p->npix - ri.startpix - ri.ringpix;
    double w = ring_weights ? ring_weights[rp] : 0.0;
    w += 1.0; w *= quadweight;
    p->h_weight[rp] = w;
    logsth[rp] = log(ri.sintheta);   // host libm, same as the reference's log(sth) (healpix_plmgen.c:101)
  }
  // ownership
  p->rp_owner.assign(nrp, 0); p->m_owner.assign(lmax + 1, 0);
  for (int rp = 0; rp < nrp; ++rp) p->rp_owner[rp] = rp_owner ? rp_owner[rp] : 0;
  for (long m = 0; m <= lmax; ++m) p->m_owner[m] = m_owner ? m_owner[m] : 0;
  for (int rp = 0; rp < nrp; ++rp)
    if (p->rp_owner[rp] < 0 || p->rp_owner[rp] >= nranks) { fprintf(stderr, "calclens_b200: rp_owner[%d] = %d outside [0, %d)\n", rp, p->rp_owner[rp], nranks); abort(); }
  for (long m = 0; m <= lmax; ++m)
    if (p->m_owner[m] < 0 || p->m_owner[m] >= nranks) { fprintf(stderr, "calclens_b200: m_owner[%ld] = %d outside [0, %d)\n", m, p->m_owner[m], nranks); abort(); }
  p->nrp_of_rank.assign(nranks, 0); p->nm_of_rank.assign(nranks, 0);
  std::vector<int> rp_local_idx(nrp), m_local_idx(lmax + 1);
  for (int rp = 0; rp < nrp; ++rp) rp_local_idx[rp] = p->nrp_of_rank[p->rp_owner[rp]]++;
  for (long m = 0; m <= lmax; ++m) m_local_idx[m] = p->nm_of_rank[p->m_owner[m]]++;
  for (int rp = 0; rp < nrp; ++rp) if (p->rp_owner[rp] == rank) p->rp_loc.push_back(rp);
  for (long m = 0; m <= lmax; ++m) if (p->m_owner[m] == rank) p->m_loc.push_back((int)m);
  p->nrp_loc = (int)p->rp_loc.size(); p->nm_loc = (int)p->m_loc.size();
  const long nslot_mine = 2L * p->nrp_loc;
  // exchange layouts
  p->g_send_count.assign(nranks, 0); p->g_recv_count.assign(nranks, 0);
  p->b_send_count.assign(nranks, 0); p->b_recv_count.assign(nranks, 0);
  std::vector<long> g_sbase(nranks), g_rbase(nranks), b_sbase(nranks), b_rbase(nranks);
  for (int q = 0; q < nranks; ++q) {
    p->g_send_count[q] = (long)p->nm_of_rank[q] * nslot_mine;           // my rings, q's m
    p->g_recv_count[q] = (long)p->nm_loc * 2L * p->nrp_of_rank[q];      // q's rings, my m
    p->b_send_count[q] = (long)p->nm_loc * 6L * 2L * p->nrp_of_rank[q]; // my m, q's rings
    p->b_recv_count[q] = (long)p->nm_of_rank[q] * 6L * nslot_mine;      // q's m, my rings
  }
  long a = 0, b = 0, c = 0, d = 0;
  for (int q = 0; q < nranks; ++q) {
    g_sbase[q] = a; a += p->g_send_count[q];
    g_rbase[q] = b; b += p->g_recv_count[q];
    b_sbase[q] = c; c += p->b_send_count[q];
    b_rbase[q] = d; d += p->b_recv_count[q];
  }
  p->g_send_total = a; p->g_recv_total = b; p->b_send_total = c; p->b_recv_total = d;
  std::vector<long> m_goff(lmax + 1), m_boff(lmax + 1), g_off(nrp), b_off(nrp);
  std::vector<int> g_stride(nrp), m_bstr(lmax + 1);
  for (long m = 0; m <= lmax; ++m) {
    int q = p->m_owner[m];
    m_goff[m] = g_sbase[q] + (long)m_local_idx[m] * nslot_mine;
    // b blocks are ordered [ring pair][field][m][hemisphere]: the block from rank q holds its nm_of_rank[q] values of m
    m_boff[m] = b_rbase[q] + 2L * m_local_idx[m]; m_bstr[m] = 2 * p->nm_of_rank[q];
  }
  for (int rp = 0; rp < nrp; ++rp) {
    int q = p->rp_owner[rp];
    int ns = 2 * p->nrp_of_rank[q];
    g_off[rp] = g_rbase[q] + 2L * rp_local_idx[rp]; g_stride[rp] = ns;
    b_off[rp] = b_sbase[q] + (long)rp_local_idx[rp] * 6L * p->nm_loc * 2L;
  }
  // upload geometry and layouts
  p->d_cth = to_device(p->h_cth); p->d_sth = to_device(p->h_sth); p->d_logsth = to_device(logsth);
  p->d_weight = to_device(p->h_weight); p->d_nphi = to_device(p->h_nphi); p->d_shifted = to_device(p->h_shifted);
  p->d_startN = to_device(p->h_startN); p->d_startS = to_device(p->h_startS);
  p->d_rp_loc = to_device(p->rp_loc); p->d_m_loc = to_device(p->m_loc);
  p->d_m_goff = to_device(m_goff); p->d_m_boff = to_device(m_boff);
  p->d_g_off = to_device(g_off); p->d_b_off = to_device(b_off);
  p->d_g_stride = to_device(g_stride); p->d_m_bstr = to_device(m_bstr);
  {
    std::vector<int> r2l(nrp, -1);
    for (int i = 0; i < p->nrp_loc; ++i) r2l[p->rp_loc[i]] = i;
    p->d_rp_to_local = to_device(r2l);
  }
  // (m, l) rows: degrees m .. lmax+1, padded
  p->h_row_off.resize(p->nm_loc + 1); p->h_alm_off.resize(p->nm_loc + 1);
  long off = 0, aoff = 0;
  for (int i = 0; i < p->nm_loc; ++i) {
    p->h_row_off[i] = off; p->h_alm_off[i] = aoff;
    long len = lmax + 2 - p->m_loc[i];
    len = ((len + 31) / 32) * 32 + kRowPad;
    off += len; aoff += lmax + 1 - p->m_loc[i];
  }
  p->h_row_off[p->nm_loc] = off; p->h_alm_off[p->nm_loc] = aoff;
  p->rows_total = off; p->alm_total = aoff;
  p->d_row_off = to_device(p->h_row_off); p->d_alm_off = to_device(p->h_alm_off);
  // plmgen_init tables (healpix_plmgen.c:221-235), host libm
  std::vector<double> mfac(lmax + 1), t1fac(lmax + 2), t2fac(2 * lmax + 3), cf(15);
  {
    const double inv_sqrt4pi = 1.0 / sqrt(4.0 * CLB_PI), inv_ln2 = 1.0 / log(2.0);
    mfac[0] = 1;
    for (long m = 1; m < lmax + 1; ++m) mfac[m] = mfac[m - 1] * sqrt((2 * m + 1.0) / (2 * m));
    for (long m = 0; m < lmax + 1; ++m) mfac[m] = inv_ln2 * log(inv_sqrt4pi * mfac[m]);
    for (long m = 0; m < lmax + 2; ++m) t1fac[m] = sqrt(4.0 * (m + 1) * (m + 1) - 1.0);
    for (long m = 0; m < 2 * lmax + 3; ++m) t2fac[m] = 1. / sqrt(m + 1.0);
    for (int m = 0; m < 15; ++m) cf[m] = ldexp(1.0, (m - 4) * 90);
  }
  double *d_mfac = to_device(mfac), *d_t1 = to_device(t1fac), *d_t2 = to_device(t2fac), *d_cf = to_device(cf);
  CLB_CUDA_CHECK(cudaMalloc(&p->d_A, sizeof(double) * std::max<long>(off, 1)));
  CLB_CUDA_CHECK(cudaMalloc(&p->d_c, sizeof(double) * std::max<long>(off, 1)));
  CLB_CUDA_CHECK(cudaMalloc(&p->d_coef, sizeof(double) * 8 * std::max<long>(off, 1)));
  const size_t npair = (size_t)std::max(p->nm_loc, 1) * nrp;
  CLB_CUDA_CHECK(cudaMalloc(&p->d_ls_ana, sizeof(int) * npair));
  CLB_CUDA_CHECK(cudaMalloc(&p->d_ls_syn, sizeof(int) * npair));
  CLB_CUDA_CHECK(cudaMalloc(&p->d_seed, sizeof(double2) * npair));
  if (p->nm_loc > 0) {
    recurrence_table_kernel<<<(p->nm_loc + 63) / 64, 64>>>(p->d_m_loc, p->nm_loc, (int)lmax, p->d_row_off, d_t1, d_t2,
                                                           p->d_A, p->d_c);
    CLB_CUDA_CHECK(cudaGetLastError());
    dim3 grid((nrp + 127) / 128, p->nm_loc);
    seed_kernel<<<grid, 128>>>(p->d_m_loc, p->nm_loc, nrp, (int)lmax, p->d_cth, p->d_sth, p->d_logsth, d_mfac, d_t1,
                               d_t2, d_cf, p->d_row_off, p->d_c, 1.0 / log(2.0), log(2.0), ldexp(1.0, 90), ldexp(1.0, -90),
                               p->d_ls_ana, p->d_ls_syn, p->d_seed);
    CLB_CUDA_CHECK(cudaGetLastError());
  }
  CLB_CUDA_CHECK(cudaDeviceSynchronize());
  cudaFree(d_mfac); cudaFree(d_t1); cudaFree(d_t2); cudaFree(d_cf);
  fft_tables_create(p);
  return p;
}

void sht_plan_destroy(ShtPlan *p)
{
  if (!p) return;
  fft_tables_destroy(p);
  void *ptrs[] = {p->d_cth, p->d_sth, p->d_logsth, p->d_weight, p->d_nphi, p->d_shifted, p->d_startN, p->d_startS,
                  p->d_rp_loc, p->d_m_loc, p->d_m_goff, p->d_m_boff, p->d_g_off, p->d_b_off, p->d_g_stride,
                  p->d_m_bstr, p->d_rp_to_local, p->d_row_off, p->d_alm_off, p->d_A, p->d_c, p->d_coef,
                  p->d_ls_ana, p->d_ls_syn, p->d_seed, p->d_part, (void *)p->d_rp_gsrc, p->d_rp_bptr};
  for (void *q : ptrs) if (q) cudaFree(q);
  delete p;
}

int *plan_rp_to_local(const ShtPlan *p) { return p->d_rp_to_local; }

// Fused exchange over peer memory.  g (analysis): every rank's ring FFT fills its own send buffer with coalesced
// local stores, and the Legendre stage of the m owner READS the (north, south) pairs of 32 adjacent ring pairs -- 1 KB
// contiguous per warp -- straight out of the ring owner's buffer over NVLink.  b (synthesis): the Legendre epilogue
// WRITES its 32-byte (north, south) pairs straight into the ring owner's receive buffer.  g_send_ptrs[q] / b_recv_ptrs[q] are the
// base addresses of rank q's buffers as seen from this process (peer mappings; this rank's own buffers for q == rank).
void sht_plan_set_peers(ShtPlan *p, void *const *g_send_ptrs, void *const *b_recv_ptrs, int nshell)
{
  // nshell = 2: every rank's buffers hold two shells back to back (second shell at + that rank's g_send_total /
  // b_recv_total elements); the second half of the pointer tables addresses it (batched two-plane pass)
  const int nrp = p->nrp, me = p->rank;
  std::vector<int> rp_local_idx(nrp), cnt_rp(p->nranks, 0);
  for (int rp = 0; rp < nrp; ++rp) rp_local_idx[rp] = cnt_rp[p->rp_owner[rp]]++;
  std::vector<const double2 *> gsrc(2 * (size_t)nrp, nullptr);
  std::vector<double2 *> bptr(2 * (size_t)nrp, nullptr);
  for (int rp = 0; rp < nrp; ++rp) {
    const int q = p->rp_owner[rp];
    const long nslot_q = 2L * p->nrp_of_rank[q];
    long sbase = 0, rbase = 0;
    for (int r = 0; r < me; ++r) {
      sbase += (long)p->nm_of_rank[r] * nslot_q;        // q's send block for rank r: nm_of_rank[r] * nslot_q elements
      rbase += (long)p->nm_of_rank[r] * 6L * nslot_q;   // q's receive block from rank r
    }
    gsrc[rp] = reinterpret_cast<const double2 *>(g_send_ptrs[q]) + sbase + 2L * rp_local_idx[rp];
    bptr[rp] = reinterpret_cast<double2 *>(b_recv_ptrs[q]) + rbase + (long)rp_local_idx[rp] * 6L * p->nm_loc * 2L;
    if (nshell >= 2) {
      gsrc[nrp + rp] = gsrc[rp] + (p->lmax + 1) * nslot_q;          // rank q's g_send_total
      bptr[nrp + rp] = bptr[rp] + (p->lmax + 1) * 6L * nslot_q;     // rank q's b_recv_total
    }
  }
  if (p->d_rp_gsrc) cudaFree(p->d_rp_gsrc);
  if (p->d_rp_bptr) cudaFree(p->d_rp_bptr);
  p->d_rp_gsrc = to_device(gsrc);
  p->d_rp_bptr = to_device(bptr);
  p->peer_shells = nshell >= 2 ? 2 : 1;
}

}  // namespace clb
