// calclens_b200/csrc/launch.cuh -- host launchers of the kernel translation units, shared by api.cu and solver.cu.
#pragma once
#include "sht_internal.cuh"
#include "raymath.cuh"

namespace clb {

ShtPlan *sht_plan_create(long order, long lmax, const double *ring_weights, int nranks, int rank, const int *rp_owner,
                         const int *m_owner);
void sht_plan_destroy(ShtPlan *p);
void sht_plan_set_peers(ShtPlan *p, void *const *g_send_ptrs, void *const *b_recv_ptrs, int nshell = 1);
int launch_ring_analysis(const ShtPlan *p, const float *d_map, double2 *d_g_send, cudaStream_t st);
int launch_ring_synthesis(const ShtPlan *p, const double2 *d_b_recv, float *const d_maps[6], cudaStream_t st);
void fft_fork(const ShtPlan *p, cudaStream_t st);   // the ring-FFT class launches up to fft_join may run on parallel streams
void fft_join(const ShtPlan *p, cudaStream_t st);
int launch_legendre_analysis(ShtPlan *p, const double2 *d_g_recv, double *d_alm_re, double *d_alm_im, int apply_filter,
                             cudaStream_t st, int nshell = 1);
int launch_legendre_synthesis(ShtPlan *p, const double *d_alm_re, const double *d_alm_im, double2 *d_b_send, cudaStream_t st,
                              int nshell = 1);
int launch_ray_step(Ray *d_rays, long nrays, const float *const d_maps[6], long order, double wp, double wpm1, double wpm2,
                    int mode, cudaStream_t st, const unsigned char *d_need = nullptr, long coarse_order = 0, int rank = 0,
                    int *d_err = nullptr, double *d_sum6 = nullptr, const unsigned char *d_safe = nullptr);
int launch_ray_init(Ray *d_rays, long nrays, long first_nest, long ray_order, double binL_2, cudaStream_t st);
int launch_ray_summary(const Ray *d_rays, long nrays, double *d_out6, cudaStream_t st);
int launch_ray_output(const Ray *d_rays, Ray *d_out, long nrays, long ray_order, cudaStream_t st);
int launch_deposit_ngp(const float *d_pos, const float *d_mass, long nparts, long order, float *d_ringmap, cudaStream_t st);
void launch_healpix_index(int what, long order, long n, const long *in, const double *th, const double *ph, long *out, cudaStream_t st);
void launch_healpix_interpol(long order, long n, const double *vec, long *pix, double *wgt, int use_table, cudaStream_t st);
int launch_maps_broadcast(const ShtPlan *p, float *const local_maps[6], float *const *peer_maps,
                          const unsigned char *d_need, long coarse_order, cudaStream_t st, const unsigned char *d_gmask = nullptr);
int launch_group_masks(const ShtPlan *p, const unsigned char *d_need, long coarse_order, unsigned char *d_gmask, cudaStream_t st);
int launch_load_density(const ShtPlan *p, const float *src, float *dst, float premul, float densmul, float backdens,
                        cudaStream_t st);
void domain_masks(long ray_order, int nranks, long coarse_order, double margin_rad, unsigned char *mask);
void safe_masks(long coarse_order, double neighbour_rad, const unsigned char *mask, unsigned char *safe);
void count_launches(int n);

}  // namespace clb

// the opaque plan handle of the C ABI
struct clb_sht_plan { clb::ShtPlan *p; };
