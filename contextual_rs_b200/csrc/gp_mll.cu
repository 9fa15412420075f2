// crs_gp_mll: exact marginal log-likelihood of every arm's GP and its gradient with respect to the
// hyper-parameters — the objective `fit_gpytorch_mll(ExactMarginalLogLikelihood(...))` climbs when
// the reference fits its per-arm SingleTaskGPs (contextual_rs/experiment_utils.py:15-48, 68-121;
// model family ibid. :30-40: Standardize + Normalize + ScaleKernel(ARD Matern-5/2 | RBF) +
// ConstantMean + homoskedastic noise).
//
// One CTA per arm, fp64, everything in shared memory.  With r = y~ - m 1, K^ = o C(l) + s2 I = L L^T:
//   log p   = -1/2 r^T alpha - sum_i log L_ii - n/2 log 2 pi,            alpha = K^^-1 r
//   d/dtheta = 1/2 tr( (alpha alpha^T - K^^-1) dK^/dtheta ),   d/dm = 1^T alpha
//   dK^/do = C,  dK^/ds2 = I,  dK^_ij/dl_k = o c'(r2_ij) (-2 D_ijk^2 / l_k^3),  c' = dc/d(r2)
// K^^-1 is never stored: L is inverted in place and (K^^-1)_ij = sum_{k >= max(i,j)} Linv_ki Linv_kj
// is formed pair by pair while the gradient is accumulated.  Transforms (Normalize bounds,
// Standardize mean / std) are the frozen ones in the state, i.e. what crs_gp_prepare learnt.
// Priors and the 1/n scaling of gpytorch's ExactMarginalLogLikelihood are elementwise and live on
// the host side (contextual_rs_b200/fit.py).  Latency-bound: O(K n^3) flops, n <= 128.
#include <math_constants.h>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kMllThreads = 256;
constexpr int kMllMaxObs = 128;

__device__ __forceinline__ void corr_and_slope(double r2, int kernel, double& c, double& cp) {
  if (kernel == CRS_MATERN52) {
    const double s = sqrt(5.0 * r2), e = exp(-s);
    c = (1.0 + s + (5.0 / 3.0) * r2) * e;
    cp = -(5.0 / 6.0) * (1.0 + s) * e;
  } else {
    c = exp(-0.5 * r2);
    cp = -0.5 * c;
  }
}

__global__ void __launch_bounds__(kMllThreads)
gp_mll_kernel(crs_gp_state st, int arm_begin, int arm_end, double* __restrict__ value,
              double* __restrict__ grad) {
  extern __shared__ double smem[];
  const int k = arm_begin + blockIdx.x;
  if (k >= arm_end) return;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int d = st.dim, ncap = st.n_cap;
  const int n = min(st.n_obs[k], ncap);
  const int lda = n + 1;
  double* A = smem;                          // [n][lda]  K^ -> L -> L^-1 (lower)
  double* xn = A + (size_t)n * lda;          // [n][d]    normalised inputs
  double* rhs = xn + (size_t)n * d;          // [n]       r -> z -> alpha
  double* red = rhs + n;                     // [8][CRS_MAX_DIM + 4] block-reduction scratch
  double* ls = red + 8 * (CRS_MAX_DIM + 4);  // [d] lengthscales
  __shared__ int s_fail;
  __shared__ double s_logdet;
  const int P = d + 3;
  const double o = st.outputscale[k], noise = st.noise[k], m = st.mean_const[k];
  const double ybar = st.y_mean[k], ysd = st.y_std[k];
  const double* X = st.train_x + (size_t)k * ncap * d;
  const double* Y = st.train_y + (size_t)k * ncap;
  if (tid == 0) { s_fail = 0; s_logdet = 0.0; }
  if (tid < d) ls[tid] = st.lengthscale[(size_t)k * d + tid];
  for (int idx = tid; idx < n * d; idx += nt) {
    const int dd = idx % d;
    xn[idx] = (X[idx] - st.norm_mins[(size_t)k * d + dd]) / st.norm_ranges[(size_t)k * d + dd];
  }
  for (int i = tid; i < n; i += nt) rhs[i] = (Y[i] - ybar) / ysd - m;
  __syncthreads();
  // ---- K^ (lower triangle) -----------------------------------------------------------------
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    if (j <= i) {
      double r2 = 0.0;
      for (int dd = 0; dd < d; ++dd) {
        const double df = (xn[i * d + dd] - xn[j * d + dd]) / ls[dd];
        r2 = fma(df, df, r2);
      }
      double c, cp;
      corr_and_slope(r2, st.kernel, c, cp);
      A[i * lda + j] = o * c + (i == j ? noise : 0.0);
    }
  }
  // ---- Cholesky, right-looking, in place -------------------------------------------------------
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const double ajj = A[j * lda + j];
    if (!(ajj > 0.0)) {
      if (tid == 0) s_fail = j + 1;
      break;  // uniform
    }
    const double ljj = sqrt(ajj);
    __syncthreads();
    if (tid == 0) { A[j * lda + j] = ljj; s_logdet += log(ljj); }
    for (int i = j + 1 + tid; i < n; i += nt) A[i * lda + j] /= ljj;
    __syncthreads();
    const int mrem = n - j - 1;
    for (int idx = tid; idx < mrem * mrem; idx += nt) {
      const int i = j + 1 + idx / mrem, c = j + 1 + idx % mrem;
      if (c <= i) A[i * lda + c] -= A[i * lda + j] * A[c * lda + j];
    }
  }
  __syncthreads();
  if (s_fail) {  // not positive definite at these hyper-parameters: -inf, zero gradient
    if (tid == 0) value[k] = -CUDART_INF;
    if (tid < P) grad[(size_t)k * P + tid] = 0.0;
    return;
  }
  // ---- alpha = K^^-1 r, keeping r^T alpha ----------------------------------------------------
  __shared__ double s_quad;
  if (tid == 0) s_quad = 0.0;
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const double zj = rhs[j] / A[j * lda + j];
    __syncthreads();
    if (tid == 0) { rhs[j] = zj; s_quad += zj * zj; }  // r^T K^^-1 r = ||L^-1 r||^2
    for (int i = j + 1 + tid; i < n; i += nt) rhs[i] -= A[i * lda + j] * zj;
  }
  for (int j = n - 1; j >= 0; --j) {
    __syncthreads();
    const double aj = rhs[j] / A[j * lda + j];
    __syncthreads();
    if (tid == 0) rhs[j] = aj;
    for (int i = tid; i < j; i += nt) rhs[i] -= A[j * lda + i] * aj;
  }
  __syncthreads();
  // ---- L <- L^-1 in place (column sweep from the right; same scheme as crs_gp_prepare) ----------
  for (int j = n - 1; j >= 0; --j) {
    const double inv = 1.0 / A[j * lda + j];
    double acc[(kMllMaxObs + kMllThreads - 1) / kMllThreads];
    int cnt = 0;
    for (int i = j + 1 + tid; i < n; i += nt, ++cnt) {
      double a = 0.0;
      for (int c = j + 1; c <= i; ++c) a += A[i * lda + c] * A[c * lda + j];
      acc[cnt] = a;
    }
    __syncthreads();
    cnt = 0;
    for (int i = j + 1 + tid; i < n; i += nt, ++cnt) A[i * lda + j] = -acc[cnt] * inv;
    if (tid == 0) A[j * lda + j] = inv;
    __syncthreads();
  }
  // ---- gradient: sum over pairs i >= j of w_ij W_ij dK^_ij ---------------------------------------
  double g[CRS_MAX_DIM + 3];
#pragma unroll
  for (int q = 0; q < CRS_MAX_DIM + 3; ++q) g[q] = 0.0;
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    if (j > i) continue;
    double kinv = 0.0;
    for (int c = i; c < n; ++c) kinv = fma(A[c * lda + i], A[c * lda + j], kinv);
    const double W = (i == j ? 0.5 : 1.0) * (rhs[i] * rhs[j] - kinv);
    double r2 = 0.0;
    for (int dd = 0; dd < d; ++dd) {
      const double df = (xn[i * d + dd] - xn[j * d + dd]) / ls[dd];
      r2 = fma(df, df, r2);
    }
    double c, cp;
    corr_and_slope(r2, st.kernel, c, cp);
    g[CRS_MAX_DIM] = fma(W, c, g[CRS_MAX_DIM]);             // d/do
    if (i == j) g[CRS_MAX_DIM + 1] += W;                    // d/ds2 (W already carries the 1/2)
    const double wc = W * o * cp * -2.0;
#pragma unroll
    for (int dd = 0; dd < CRS_MAX_DIM; ++dd) {
      if (dd < d) {
        const double df = xn[i * d + dd] - xn[j * d + dd];
        g[dd] = fma(wc, df * df / (ls[dd] * ls[dd] * ls[dd]), g[dd]);
      }
    }
  }
  for (int i = tid; i < n; i += nt) g[CRS_MAX_DIM + 2] += rhs[i];  // d/dm = 1^T alpha
  // block reduction of the d + 3 accumulators
#pragma unroll
  for (int q = 0; q < CRS_MAX_DIM + 3; ++q) g[q] = warp_sum(g[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < CRS_MAX_DIM + 3; ++q) red[warp * (CRS_MAX_DIM + 4) + q] = g[q];
  }
  __syncthreads();
  if (tid < CRS_MAX_DIM + 3) {
    double t = 0.0;
    for (int w = 0; w < (nt >> 5); ++w) t += red[w * (CRS_MAX_DIM + 4) + tid];
    if (tid < d) grad[(size_t)k * P + tid] = t;
    else if (tid >= CRS_MAX_DIM) grad[(size_t)k * P + d + (tid - CRS_MAX_DIM)] = t;
  }
  if (tid == 0) value[k] = -0.5 * s_quad - s_logdet - 0.5 * n * 1.8378770664093453;  // log 2 pi
}

}  // namespace

int launch_gp_mll(const crs_gp_state* st, int arm_begin, int arm_end, double* value, double* grad,
                  cudaStream_t s) {
  if (!st || !value || !grad || arm_begin < 0 || arm_end > st->num_arms || arm_begin > arm_end)
    return CRS_ERR_INVALID_ARG;
  if (st->dim < 1 || st->dim > CRS_MAX_DIM) return CRS_ERR_UNSUPPORTED;
  const int n = st->n_max > 0 && st->n_max <= st->n_cap ? st->n_max : st->n_cap;
  if (n > kMllMaxObs) return CRS_ERR_UNSUPPORTED;
  if (arm_begin == arm_end) return CRS_OK;
  const size_t smem = sizeof(double) * ((size_t)n * (n + 1) + (size_t)n * st->dim + n + 8 * (CRS_MAX_DIM + 4) + CRS_MAX_DIM);
  if (smem > 40 * 1024)
    CRS_CUDA_TRY(cudaFuncSetAttribute(gp_mll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "mll attr");
  gp_mll_kernel<<<arm_end - arm_begin, kMllThreads, smem, s>>>(*st, arm_begin, arm_end, value, grad);
  CRS_CUDA_TRY(cudaGetLastError(), "gp_mll launch");
  return CRS_OK;
}

}  // namespace crs
