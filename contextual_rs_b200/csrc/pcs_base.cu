// crs_pcs_counts_base: the sampling half of estimate_current_generalized_pcs for a ModelListGP when
// the CALLER supplies the standard normals (`base_samples`, contextual_rs/generalized_pcs.py:449-451;
// the shape the reference hands to `posterior.rsample`: num_samples x n_c x K).  Each normal is used
// as the marginal draw of its (sample, context, arm):
//     y[s, c, k] = mean[c, k] + sqrt(var[c, k]) * z[s, c, k]        (separate IEEE mul and add)
//     counts[c]  = #{ s : y[s, c, b_c] - max_{k != b_c} y[s, c, k] > 0 },  b_c = first arg-max of mean[c, :]
// — the sort / gather / max / delta > 0 of generalized_pcs.py:464-473 without the sort.  The reference's
// rsample correlates contexts within an arm (joint MVN); PCS(c) only depends on the marginals across
// arms at context c (generalized_pcs.py:382-395), so every per-context estimate has the same law.
// Reads S * n_c * K normals once: meant for the reference's call shape (tens of arms and contexts,
// 64 - 1,024 samples), HBM-bound beyond that.  One warp per (sample, context) pair, lanes over arms.
#include <math_constants.h>
#include <climits>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kBaseThreads = 256;

__device__ __forceinline__ double mul_add_rn(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
__device__ __forceinline__ float mul_add_rn(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
__device__ __forceinline__ double sqrt_rn(double x) { return __dsqrt_rn(x); }
__device__ __forceinline__ float sqrt_rn(float x) { return __fsqrt_rn(x); }

template <typename T>
__global__ void __launch_bounds__(kBaseThreads)
pcs_base_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K, int64_t ld,
                const T* __restrict__ base, int64_t num_samples, uint32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kBaseThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kBaseThreads) >> 5;
  // warps walk contexts; all samples of a context stay in one warp so that the count needs no atomics
  for (int64_t c = warp; c < n_c; c += nwarps) {
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    T bm = -(T)CUDART_INF;
    int bi = INT_MAX;
    for (int a = lane; a < K; a += 32) {
      const T m = mrow[a];
      if (m > bm) { bm = m; bi = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T om = __shfl_xor_sync(kFull, bm, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
    }
    if (bi == INT_MAX) bi = 0;
    const T sd_b = sqrt_rn(vrow[bi]);
    const T mu_b = mrow[bi];
    uint32_t hits = 0;
    for (int64_t s = 0; s < num_samples; ++s) {
      const T* z = base + (s * n_c + c) * (int64_t)K;
      T rest = -(T)CUDART_INF;
      for (int a = lane; a < K; a += 32)
        if (a != bi) rest = fmax(rest, mul_add_rn(sqrt_rn(vrow[a]), z[a], mrow[a]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rest = fmax(rest, __shfl_xor_sync(kFull, rest, o));
      const T yb = mul_add_rn(sd_b, z[bi], mu_b);
      hits += (yb - rest > (T)0) ? 1u : 0u;  // func_I = strict indicator of delta
    }
    if (lane == 0) counts[c] = hits;
  }
}

}  // namespace

int launch_pcs_counts_base(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                           const void* base, int64_t num_samples, uint32_t* counts, cudaStream_t s) {
  if (!mean || !var || !counts || n_c < 0 || K < 2 || ld < K || num_samples < 0) return CRS_ERR_INVALID_ARG;
  if (num_samples > 0 && !base) return CRS_ERR_INVALID_ARG;
  if (num_samples > 0xffffffffLL) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0) return CRS_OK;
  const int64_t want = (n_c * 32 + kBaseThreads - 1) / kBaseThreads;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  if (dtype == CRS_F64)
    pcs_base_kernel<double><<<blocks, kBaseThreads, 0, s>>>((const double*)mean, (const double*)var, n_c, K, ld,
                                                            (const double*)base, num_samples, counts);
  else if (dtype == CRS_F32)
    pcs_base_kernel<float><<<blocks, kBaseThreads, 0, s>>>((const float*)mean, (const float*)var, n_c, K, ld,
                                                           (const float*)base, num_samples, counts);
  else
    return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "pcs_counts_base launch");
  return CRS_OK;
}

}  // namespace crs
