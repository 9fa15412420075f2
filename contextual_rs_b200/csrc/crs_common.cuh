// Shared device helpers for libcrs_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include <utility>
#include "../../include/crs_b200.h"

namespace crs {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// gpytorch.settings.min_variance: 1e-10 for double, 1e-6 for float.
template <typename T> __host__ __device__ inline T variance_floor();
template <> __host__ __device__ inline double variance_floor<double>() { return 1e-10; }
template <> __host__ __device__ inline float variance_floor<float>() { return 1e-6f; }

// Stationary ARD correlation c(r) from the squared scaled distance r2 (outputscale folded into
// alpha / linv_t by crs_gp_prepare).  Matern-5/2: (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r); RBF: exp(-r2/2).
template <typename T> __device__ __forceinline__ T corr_matern52(T r2);
template <> __device__ __forceinline__ double corr_matern52<double>(double r2) {
  const double s5r = 2.23606797749978969641 * sqrt(r2);
  const double poly = (s5r + 1.0) + (5.0 / 3.0) * r2;
  return poly * exp(-s5r);
}
template <> __device__ __forceinline__ float corr_matern52<float>(float r2) {
  const float s5r = 2.2360679775f * sqrtf(r2);
  const float poly = (s5r + 1.0f) + (5.0f / 3.0f) * r2;
  return poly * expf(-s5r);
}
template <typename T> __device__ __forceinline__ T corr_rbf(T r2);
template <> __device__ __forceinline__ double corr_rbf<double>(double r2) { return exp(-0.5 * r2); }
template <> __device__ __forceinline__ float corr_rbf<float>(float r2) { return expf(-0.5f * r2); }

template <typename T> __device__ __forceinline__ T correlation(T r2, int kernel) {
  return kernel == CRS_MATERN52 ? corr_matern52<T>(r2) : corr_rbf<T>(r2);
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ---- layout of the prepared per-arm state (written by gp_prepare, bulk-copied by the grid kernel)
// arm_head [K][kHeadElems] T : sub[16] | mul[16] | (o, m, ybar, ysd) | pad[4]
constexpr int kHeadSub = 0, kHeadMul = CRS_MAX_DIM, kHeadScal = 2 * CRS_MAX_DIM;
constexpr int kHeadElems = 2 * CRS_MAX_DIM + 8;
// linv_t is stored panel-major: panel p = rows j in [0, min(n_cap, 32 (p+1))) x columns
// [32 p, 32 p + 32), row-major with a fixed row stride — exactly the block the grid kernel needs for
// its p-th chunk of v, so that one 1-D bulk copy stages it.  The row stride is 40 elements for fp64
// (a 16-bank skew per row makes the DMMA fragment loads of 4 consecutive rows conflict-free) and 32
// for fp32.
constexpr int kPanel = 32;
__host__ __device__ inline int panel_ld(int elem_size) { return elem_size == 8 ? 40 : 32; }
template <typename T> struct PanelLd { static constexpr int value = sizeof(T) == 8 ? 40 : 32; };
__host__ __device__ inline int panel_rows(int n_cap, int p) {
  const int r = (p + 1) * kPanel;
  return r < n_cap ? r : n_cap;
}
__host__ __device__ inline int num_panels(int n_cap) { return (n_cap + kPanel - 1) / kPanel; }
__host__ __device__ inline long long panel_offset(int n_cap, int p, int elem_size) {
  long long off = 0;
  for (int q = 0; q < p; ++q) off += (long long)panel_rows(n_cap, q) * panel_ld(elem_size);
  return off;
}
__host__ __device__ inline long long linv_stride(int n_cap, int elem_size) {
  return panel_offset(n_cap, num_panels(n_cap), elem_size);
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Programmatic dependent launch (prepare -> grid -> scan of one decision): a kernel launched with
// launch_pdl may become resident while its predecessor in the stream is still running (once every
// CTA of the predecessor has called griddep_launch_dependents or exited) and must call
// griddep_wait() before it touches anything the predecessor writes; the wait returns when the
// predecessor has completed and its writes are visible.  Without a PDL predecessor both are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CRS_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

inline int dim_pad_of(int d) {
  int p = 1;
  while (p < d) p <<= 1;
  return p;
}

// Host-side error plumbing shared by the C ABI.
void set_cuda_error(cudaError_t e, const char* where);
#define CRS_CUDA_TRY(expr, where)                 \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) {                      \
      (void)cudaGetLastError(); /* reported here: must not resurface as the next launch's error */ \
      crs::set_cuda_error(_e, where);             \
      return CRS_ERR_CUDA;                        \
    }                                             \
  } while (0)

// Kernel launchers (one translation unit each).
int launch_gp_prepare(const crs_gp_state* st, int arm_begin, int arm_end, cudaStream_t s);
int launch_gp_prepare_copy(const crs_gp_state* st, int arm_begin, int arm_end, const void* cp_src, void* cp_dst,
                           size_t cp_bytes, cudaStream_t s);
int launch_posterior_grid(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                          int arm_end, void* mean, void* var, int64_t ld, int col_offset,
                          cudaStream_t s);
int launch_ocba_scan(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                     int arm_offset, const int32_t* ext_best_arm, const void* ext_best_mean,
                     const void* ext_best_var, int32_t* best_arm, void* min_zeta, int32_t* min_arm,
                     int32_t* tie_cnt, crs_decision* decision, void* workspace, cudaStream_t s);
int launch_ocba_resolve_tie(const void* mean, const void* var, int64_t n_c, int K, int64_t ld,
                            int dtype, const int32_t* best_arm, const void* min_zeta,
                            const int32_t* tie_cnt, int64_t r, crs_decision* decision,
                            cudaStream_t s);
int launch_ocba_select(const crs_gp_state* st, const void* ctx, const void* mean, const void* var,
                       int64_t ld, int arm_offset, int p_mode, double kernel_scale,
                       crs_decision* decision, crs_decision* host_mirror, int seq, const crs_peer_exchange* px,
                       cudaStream_t s);
int launch_arm_pack_best(const void* mean, const void* var, int64_t n_c, int64_t ld, int dtype, int arm_offset,
                         const int32_t* best_arm, double* pack, cudaStream_t s);
int launch_arm_combine_best(const double* pack_all, int world, int64_t n_c, int dtype, int32_t* ext_arm, void* ext_mean,
                            void* ext_var, cudaStream_t s);
int launch_arm_record_mean(crs_decision* decision, const void* mean, int64_t ld, int dtype, int arm_offset, cudaStream_t s);
int launch_arm_fold(const crs_decision* records, int world, const int32_t* ext_arm, crs_decision* out, cudaStream_t s);
int launch_arm_finish(crs_decision* decision, const double* psi, const int32_t* bad, cudaStream_t s);
int launch_peer_sum(const uint64_t* values, int n, uint64_t* out, uint64_t* host_mirror, const crs_peer_exchange* px,
                    cudaStream_t s);
int launch_peer_exchange_only(crs_decision* decision, crs_decision* host_mirror, int seq, const crs_peer_exchange* px,
                              cudaStream_t s);
int launch_ocba_scan_select(const void* mean, const void* var, int64_t n_c, int K, int64_t ld,
                            int dtype, int arm_offset, const int32_t* ext_best_arm,
                            const void* ext_best_mean, const void* ext_best_var, int32_t* best_arm,
                            void* min_zeta, int32_t* min_arm, int32_t* tie_cnt,
                            crs_decision* decision, void* workspace, const crs_gp_state* st,
                            const void* ctx, int p_mode, double kernel_scale,
                            crs_decision* host_mirror, int seq, const crs_peer_exchange* px, cudaStream_t s);
int64_t scan_workspace_bytes();
int launch_gp_mll(const crs_gp_state* st, int arm_begin, int arm_end, double* value, double* grad,
                  cudaStream_t s);
int launch_min_zeta_grad(const crs_gp_state* st, const void* ctx, int64_t n_c, const void* mean,
                         const void* var, int64_t ld, const int32_t* best_arm, const int32_t* min_arm,
                         void* grad, cudaStream_t s);
int launch_levi_scan(const crs_gp_state* st, const void* mean, const void* var, int64_t n_c, int64_t ld,
                     const void* weights, void* ei_out, int64_t ld_ei, void* max_ei, int32_t* max_arm,
                     crs_levi_decision* decision, cudaStream_t s);
int launch_pcs_counts(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                      int64_t num_samples, uint64_t seed, uint32_t offset, int arm_offset,
                      int64_t ctx_offset, uint32_t* counts, uint64_t* stats, cudaStream_t s);
int launch_pcs_counts_tree(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                           int64_t num_samples, uint64_t seed, uint32_t offset, int64_t ctx_offset,
                           uint32_t* counts, uint64_t* stats, cudaStream_t s);
int64_t pcs_cdf_workspace_bytes(int64_t n_c);
int launch_pcs_counts_cdf(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                          int64_t num_samples, uint64_t seed, uint32_t offset, int64_t ctx_offset,
                          uint32_t* counts, uint64_t* stats, void* workspace, cudaStream_t s);
int launch_cdf_probe(int kind, int iters, int blocks, void* out, cudaStream_t s);
int launch_pcs_counts_base(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                           const void* base, int64_t num_samples, uint32_t* counts, cudaStream_t s);
int launch_tree_zfun(const float* t, int64_t n, float* out, cudaStream_t s);
int launch_tree_probe(int kind, int iters, int blocks, void* out, cudaStream_t s);
int launch_debug_corr(const double* r2, int64_t n, int kernel, int variant, double* out, cudaStream_t s);
int launch_peak_probe(int kind, int iters, int blocks, void* out, cudaStream_t s);
int launch_pcs_probe(int iters, int blocks, void* out, cudaStream_t s);
int launch_philox_words(const uint32_t* ctr, int64_t n, uint32_t k0, uint32_t k1, uint32_t* out,
                        cudaStream_t s);

}  // namespace crs
