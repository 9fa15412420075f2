// crs_levi_scan: the LEVI grid acquisition of Pearce & Branke as the reference implements it —
// replaces contextual_rs/levi.py:132-172 (`_compute_all_EI_reviewer`), :175-196 (`PredictiveEI`)
// and :199-227 (`discrete_levi`) on a materialised posterior grid:
//     Delta_a  = -| mu_a - max_{j != a} mu_j |
//     sigma~_a = var_a / sqrt(var_a + noise_a * stdv_a^2)        (var clamped at 1e-9)
//     EI_a     = sigma~_a * (pdf(u) + u * cdf(u)),  u = Delta_a / sigma~_a
// then the row maximum over arms (first index on ties), optional context weights, and the first
// arg-max over contexts.  The reference builds an n_c x K x K "expanded mean" to get the leave-one-
// out maximum; the two largest means of the row give the same thing.  One warp owns one context
// row: pass 1 keeps the top two means per lane and merges them by shuffle, pass 2 evaluates EI.
// Reads 2 K n_c w bytes once (the second pass hits L1/L2); the erf / exp per entry make it
// FP64-pipe-bound in fp64 (~70 FP64 instructions per entry) and HBM-bound in fp32.
#include <math_constants.h>
#include <climits>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kLeviThreads = 256;

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ double neg_inf<double>() { return -CUDART_INF; }
template <> __device__ __forceinline__ float neg_inf<float>() { return -CUDART_INF_F; }

// torch.distributions.Normal(0, 1): cdf = 0.5 (1 + erf(u / sqrt 2)); pdf = exp(-u^2 / 2 - log sqrt(2 pi))
__device__ __forceinline__ double std_cdf(double u) { return 0.5 * (1.0 + erf(u / 1.4142135623730951)); }
__device__ __forceinline__ float std_cdf(float u) { return 0.5f * (1.0f + erff(u / 1.4142135623730951f)); }
__device__ __forceinline__ double std_pdf(double u) { return exp(-(u * u) / 2.0 - 0.9189385332046727); }
__device__ __forceinline__ float std_pdf(float u) { return expf(-(u * u) / 2.0f - 0.9189385332046727f); }

template <typename T>
__device__ __forceinline__ T levi_ei(T mu, T vr, T other_max, T noise_orig) {
  const T delta = -fabs(mu - other_max);                 // levi.py:153
  const T v = vr < T(1e-9) ? T(1e-9) : vr;               // :155 clamp_min(1e-9)
  const T sigma = v / sqrt(v + noise_orig);              // :156-162
  const T u = delta / sigma;                             // :164
  return sigma * (std_pdf(u) + u * std_cdf(u));          // :165-168
}

template <typename T>
__global__ void __launch_bounds__(kLeviThreads)
levi_scan_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K, int64_t ld,
                 const double* __restrict__ noise, const double* __restrict__ y_std,
                 const T* __restrict__ weights, T* __restrict__ ei_out, int64_t ld_ei,
                 T* __restrict__ max_ei, int32_t* __restrict__ max_arm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t c = (int64_t)blockIdx.x * nwarps + warp; c < n_c; c += (int64_t)gridDim.x * nwarps) {
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    // pass 1: the two largest means of the row (equal to each other when the maximum is tied)
    T m1 = neg_inf<T>(), m2 = neg_inf<T>();
    for (int a = lane; a < K; a += 32) {
      const T m = mrow[a];
      if (m > m1) { m2 = m1; m1 = m; }
      else if (m > m2) m2 = m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T o1 = __shfl_xor_sync(kFull, m1, o), o2 = __shfl_xor_sync(kFull, m2, o);
      const T lo = fmin(m1, o1);
      m1 = fmax(m1, o1);
      m2 = fmax(lo, fmax(m2, o2));
    }
    // pass 2: EI per arm, running maximum (first index wins ties, like torch.max)
    T best = neg_inf<T>();
    int barm = INT_MAX;
    for (int a = lane; a < K; a += 32) {
      const T mu = mrow[a];
      const double sd = y_std[a];
      const T ei = levi_ei<T>(mu, vrow[a], mu == m1 ? m2 : m1, (T)(noise[a] * (sd * sd)));
      if (ei_out) ei_out[c * ld_ei + a] = ei;
      if (ei > best) { best = ei; barm = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ob = __shfl_xor_sync(kFull, best, o);
      const int oa = __shfl_xor_sync(kFull, barm, o);
      if (ob > best || (ob == best && oa < barm)) { best = ob; barm = oa; }
    }
    if (lane == 0) {
      if (barm == INT_MAX) barm = 0;  // all-NaN row
      max_ei[c] = weights ? best * weights[c] : best;  // levi.py:221-223
      max_arm[c] = barm;
    }
  }
}

// first arg-max over contexts (levi.py:224): one CTA, contiguous chunks so that "first" is kept
template <typename T>
__global__ void __launch_bounds__(1024)
levi_argmax_kernel(const T* __restrict__ max_ei, const int32_t* __restrict__ max_arm, int64_t n_c,
                   crs_levi_decision* decision) {
  __shared__ double s_v[32];
  __shared__ long long s_i[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t chunk = (n_c + 1023) / 1024;
  const int64_t c0 = (int64_t)tid * chunk, c1 = min(c0 + chunk, n_c);
  double bv = -CUDART_INF;
  long long bi = LLONG_MAX;
  for (int64_t c = c0; c < c1; ++c) {
    const double v = (double)max_ei[c];
    if (v > bv) { bv = v; bi = c; }
  }
  auto better = [](double v, long long i, double bv, long long bi) { return v > bv || (v == bv && i < bi); };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(kFull, bv, o);
    const long long oi = __shfl_xor_sync(kFull, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w)
      if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; }
    if (bi == LLONG_MAX) bi = 0;
    decision->value = bv;
    decision->context = (int32_t)bi;
    decision->arm = max_arm[bi];
    decision->reserved = 0;
  }
}

}  // namespace

int launch_levi_scan(const crs_gp_state* st, const void* mean, const void* var, int64_t n_c, int64_t ld,
                     const void* weights, void* ei_out, int64_t ld_ei, void* max_ei, int32_t* max_arm,
                     crs_levi_decision* decision, cudaStream_t s) {
  if (!st || !mean || !var || !max_ei || !max_arm || n_c < 0 || n_c > INT_MAX || ld < st->num_arms)
    return CRS_ERR_INVALID_ARG;
  if (st->num_arms < 2) return CRS_ERR_INVALID_ARG;  // the leave-one-out maximum needs another arm
  if (ei_out && ld_ei < st->num_arms) return CRS_ERR_INVALID_ARG;
  if (n_c == 0) return CRS_OK;
  const int K = st->num_arms;
  const int wpb = kLeviThreads / 32;
  const int64_t want = (n_c + wpb - 1) / wpb;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  if (st->dtype == CRS_F64) {
    levi_scan_kernel<double><<<blocks, kLeviThreads, 0, s>>>((const double*)mean, (const double*)var, n_c, K, ld,
        st->noise, st->y_std, (const double*)weights, (double*)ei_out, ld_ei, (double*)max_ei, max_arm);
    if (decision) levi_argmax_kernel<double><<<1, 1024, 0, s>>>((const double*)max_ei, max_arm, n_c, decision);
  } else if (st->dtype == CRS_F32) {
    levi_scan_kernel<float><<<blocks, kLeviThreads, 0, s>>>((const float*)mean, (const float*)var, n_c, K, ld,
        st->noise, st->y_std, (const float*)weights, (float*)ei_out, ld_ei, (float*)max_ei, max_arm);
    if (decision) levi_argmax_kernel<float><<<1, 1024, 0, s>>>((const float*)max_ei, max_arm, n_c, decision);
  } else {
    return CRS_ERR_INVALID_ARG;
  }
  CRS_CUDA_TRY(cudaGetLastError(), "levi_scan launch");
  return CRS_OK;
}

}  // namespace crs
