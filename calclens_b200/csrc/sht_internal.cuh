// calclens_b200/csrc/sht_internal.cuh -- internal plan structure shared by the SHT translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CLB_CUDA_CHECK(call)                                                                              \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess) {                                                                             \
      fprintf(stderr, "calclens_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e__), __FILE__,  \
              __LINE__, #call);                                                                           \
      abort(); /* the reference's failure mode on this path is MPI_Abort (map2alm_transpose_mpi.c:129-138) */ \
    }                                                                                                     \
  } while (0)

namespace clb {

// SM count of the current device (cached per device)
inline int sm_count()
{
  static int cache[64] = {};
  int dev = 0;
  CLB_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) CLB_CUDA_CHECK(cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev));
  return cache[dev];
}

constexpr int kSeedAlign = 16;    // Legendre l-blocks start at m + k*kSeedAlign
constexpr int kRowPad = 64;       // zero padding at the end of every (m, l) table row
constexpr int kNoStart = 0x7fffffff;

// One GPU's view of the transform.  Ring pairs rp = 0 .. 2*Nside-1 (north ring rp+1, south ring 4*Nside-1-rp,
// the last pair is the equator with no southern partner).  Every ring pair owns two "slots" (N, S).
struct ShtPlan {
  long order, nside, npix, lmax;
  int nranks, rank;
  int nrp;                 // 2*Nside ring pairs on the sphere
  int nrp_loc;             // ring pairs whose pixels (FFT stage) this rank owns
  int nm_loc;              // m values whose Legendre stage this rank owns
  // host copies
  std::vector<int> rp_owner, m_owner;      // owner rank per ring pair / per m
  std::vector<int> rp_loc, m_loc;          // ascending lists owned by this rank
  std::vector<int> nrp_of_rank, nm_of_rank;
  std::vector<double> h_cth, h_sth, h_weight;
  std::vector<int> h_nphi, h_shifted;
  std::vector<long> h_startN, h_startS;
  // device geometry (all ring pairs)
  double *d_cth = nullptr, *d_sth = nullptr, *d_logsth = nullptr, *d_weight = nullptr;
  int *d_nphi = nullptr, *d_shifted = nullptr;
  long *d_startN = nullptr, *d_startS = nullptr;
  int *d_rp_loc = nullptr;   // [nrp_loc] global ring-pair index of local slot pair
  int *d_rp_to_local = nullptr;   // [nrp] local index of a ring pair, -1 if not owned
  int *d_m_loc = nullptr;    // [nm_loc]
  // exchange layouts (in double2 elements)
  //  analysis:  FFT side writes  g_send[m_goff[m] + slot_loc]             (slot_loc = 2*rp_local + hemi)
  //             Legendre side reads g_recv[g_off[rp] + m_idx*g_stride[rp] + hemi]
  //  synthesis: blocks ordered [ring pair][field][m][hemisphere] (a ring's b_m are one contiguous run for its FFT)
  //             Legendre side writes b_send[b_off[rp] + (f*nm_loc + m_idx)*2 + hemi]
  //             FFT side reads   b_recv[m_boff[m] + (rp_local*6 + f)*m_bstr[m] + hemi]
  long *d_m_goff = nullptr, *d_m_boff = nullptr;   // [lmax+1]
  long *d_g_off = nullptr, *d_b_off = nullptr;     // [nrp]
  int *d_g_stride = nullptr;                       // [nrp]
  int *d_m_bstr = nullptr;                         // [lmax+1]
  std::vector<long> g_send_count, g_recv_count, b_send_count, b_recv_count;  // per peer, in double2 elements
  long g_send_total = 0, g_recv_total = 0, b_send_total = 0, b_recv_total = 0;
  // recurrence tables for the local m rows: row(m_idx) starts at row_off[m_idx], index l-m, length row_len
  std::vector<long> h_row_off;
  long rows_total = 0;
  long *d_row_off = nullptr;
  double *d_A = nullptr;     // mu_{l+1} = (x*A_l)*mu_l - mu_{l-1}
  double *d_c = nullptr;     // lambda_l = c_l * mu_l
  // seeds per (m_idx, rp): block-aligned start degree and the two scaled values there
  int *d_ls_ana = nullptr, *d_ls_syn = nullptr;
  double2 *d_seed = nullptr;   // (mu_{ls-1}, mu_{ls})
  // plane-dependent scratch
  double *d_coef = nullptr;    // [rows_total][8 * coef_shells]: A, Pre, Pim, Dre, Dim, Kre, Kim, 0 (one shell; legendre.cu for two)
  int coef_shells = 1;
  double *d_part = nullptr;    // analysis partial sums [shell][row][alm_total][2] (ana_nchunk = rows * shells allocated)
  int ana_nchunk = 0, ana_chunk = 0, syn_chunk = 0;
  long alm_total = 0;          // sum over local m of (lmax-m+1)
  std::vector<long> h_alm_off;
  long *d_alm_off = nullptr;
  // fused exchange over peer memory (clb_sht_plan_set_peers): destination of every m's g block / every ring pair's b
  // block inside the owning rank's receive buffer (NVLink peer pointers, or local ones for this rank's own share)
  const double2 **d_rp_gsrc = nullptr;   // [2][nrp] (second half: second shell of a batched pass) analysis PULLS g: where ring pair rp's block lives in its owner's send buffer
  double2 **d_rp_bptr = nullptr;         // [2][nrp] synthesis PUSHES b: ring pair rp's block in its owner's receive buffer
  int peer_shells = 1;                   // shells the peer buffers hold (clb_sht_plan_set_peers_shells)
  // FFT tables
  struct FftTables *fft = nullptr;
};

}  // namespace clb
