// crs_min_zeta_grad: d(-min_j zeta_{c,j}) / d(context c) — the gradient `optimize_acqf` takes
// through `MinZeta.forward` by autograd in the reference's continuous-context driver
// (contextual_rs/continuous_context.py:57-74; experiments/continuous_experiments/main.py:252-263).
//
// zeta* = (mu_b - mu_j)^2 / (var_b + var_j) only involves two arms per context — the predicted
// best arm b and the minimiser j, both outputs of crs_ocba_scan — so the backward pass costs
// 2 / K of the forward.  Per (context, arm), with x~ the normalised / lengthscale-scaled context,
// c_i = corr(||x~ - X~_i||^2), c'_i = dc/d(r^2) (Matern-5/2: -(5/6)(1 + s) e^-s, RBF: -c/2):
//   d mu  / dx_k =  ysd       sum_i alpha_i c'_i 2 (x~_k - X~_ik) mul_k
//   d var / dx_k = -ysd^2 2   sum_i u_i     c'_i 2 (x~_k - X~_ik) mul_k,   u = M^T (M c), M = o L^-1
// (var clamped at the floor has zero gradient).  One CTA per context; latency-bound and tiny next to
// the grid kernel, so it reads the arm state straight from L2.
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kGradThreads = 128;

template <typename T>
__device__ __forceinline__ double linv_at(const T* __restrict__ lt, int ncap, int j, int i) {
  const int p = i / kPanel;
  return (double)lt[panel_offset(ncap, p, sizeof(T)) + (long long)j * PanelLd<T>::value + (i - p * kPanel)];
}

template <typename T>
__global__ void __launch_bounds__(kGradThreads)
min_zeta_grad_kernel(crs_gp_state st, const T* __restrict__ ctx, int64_t n_c, const T* __restrict__ mean,
                     const T* __restrict__ var, int64_t ld, const int32_t* __restrict__ best_arm,
                     const int32_t* __restrict__ min_arm, T* __restrict__ grad) {
  __shared__ double s_c[CRS_MAX_OBS], s_cp[CRS_MAX_OBS], s_v[CRS_MAX_OBS], s_u[CRS_MAX_OBS];
  __shared__ double s_x[CRS_MAX_DIM], s_dmu[2][CRS_MAX_DIM], s_dvar[2][CRS_MAX_DIM];
  const int tid = threadIdx.x;
  const int d = st.dim, dp = st.dim_pad, ncap = st.n_cap;
  const T* g_head = reinterpret_cast<const T*>(st.arm_head);
  const T* g_alpha = reinterpret_cast<const T*>(st.alpha);
  const T* g_xs = reinterpret_cast<const T*>(st.xs);
  const T* g_lt = reinterpret_cast<const T*>(st.linv_t);
  const long long lt_stride = linv_stride(ncap, sizeof(T));
  for (int64_t c = blockIdx.x; c < n_c; c += gridDim.x) {
    const int arms[2] = {best_arm[c], min_arm[c]};
    for (int which = 0; which < 2; ++which) {
      const int a = arms[which];
      __syncthreads();
      if (a < 0) {  // no competitor (uniform branch)
        if (tid < d) { s_dmu[which][tid] = 0.0; s_dvar[which][tid] = 0.0; }
        continue;
      }
      const T* head = g_head + (size_t)a * kHeadElems;
      const int n = min(st.n_obs[a], ncap);
      if (tid < d) s_x[tid] = ((double)ctx[c * d + tid] - (double)head[kHeadSub + tid]) * (double)head[kHeadMul + tid];
      __syncthreads();
      const T* xs = g_xs + (size_t)a * ncap * dp;
      for (int i = tid; i < n; i += kGradThreads) {
        double r2 = 0.0;
        for (int k = 0; k < d; ++k) {
          const double df = s_x[k] - (double)xs[i * dp + k];
          r2 = fma(df, df, r2);
        }
        double cv, cp;
        if (st.kernel == CRS_MATERN52) {
          const double s = sqrt(5.0 * r2), e = exp(-s);
          cv = (1.0 + s + (5.0 / 3.0) * r2) * e;
          cp = -(5.0 / 6.0) * (1.0 + s) * e;
        } else {
          cv = exp(-0.5 * r2);
          cp = -0.5 * cv;
        }
        s_c[i] = cv;
        s_cp[i] = cp;
      }
      __syncthreads();
      const T* lt = g_lt + (size_t)a * lt_stride;
      for (int i = tid; i < n; i += kGradThreads) {  // v = M c
        double acc = 0.0;
        for (int j = 0; j <= i; ++j) acc = fma(linv_at<T>(lt, ncap, j, i), s_c[j], acc);
        s_v[i] = acc;
      }
      __syncthreads();
      for (int j = tid; j < n; j += kGradThreads) {  // u = M^T v
        double acc = 0.0;
        for (int i = j; i < n; ++i) acc = fma(linv_at<T>(lt, ncap, j, i), s_v[i], acc);
        s_u[j] = acc;
      }
      __syncthreads();
      if (tid < d) {
        const double ysd = (double)head[kHeadScal + 3], mul = (double)head[kHeadMul + tid];
        const T* alpha = g_alpha + (size_t)a * ncap;
        double gm = 0.0, gq = 0.0;
        for (int i = 0; i < n; ++i) {
          const double t = s_cp[i] * 2.0 * (s_x[tid] - (double)xs[i * dp + tid]);
          gm = fma((double)alpha[i], t, gm);
          gq = fma(s_u[i], t, gq);
        }
        const bool clamped = !(var[c * ld + a] > variance_floor<T>());
        s_dmu[which][tid] = ysd * gm * mul;
        s_dvar[which][tid] = clamped ? 0.0 : -ysd * ysd * 2.0 * gq * mul;
      }
    }
    __syncthreads();
    if (tid < d) {
      double g = 0.0;
      if (arms[0] >= 0 && arms[1] >= 0) {
        const double mb = (double)mean[c * ld + arms[0]], mj = (double)mean[c * ld + arms[1]];
        const double s = (double)var[c * ld + arms[0]] + (double)var[c * ld + arms[1]];
        const double gap = mb - mj;
        const double dz = 2.0 * gap * (s_dmu[0][tid] - s_dmu[1][tid]) / s
                          - gap * gap * (s_dvar[0][tid] + s_dvar[1][tid]) / (s * s);
        g = -dz;  // MinZeta = -min zeta
      }
      grad[c * d + tid] = (T)g;
    }
  }
}

}  // namespace

int launch_min_zeta_grad(const crs_gp_state* st, const void* ctx, int64_t n_c, const void* mean,
                         const void* var, int64_t ld, const int32_t* best_arm, const int32_t* min_arm,
                         void* grad, cudaStream_t s) {
  if (!st || !ctx || !mean || !var || !best_arm || !min_arm || !grad || n_c < 0 || ld < st->num_arms)
    return CRS_ERR_INVALID_ARG;
  if (st->n_cap > CRS_MAX_OBS || st->dim > CRS_MAX_DIM) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0) return CRS_OK;
  const int blocks = (int)(n_c < 148 * 16 ? n_c : 148 * 16);
  if (st->dtype == CRS_F64)
    min_zeta_grad_kernel<double><<<blocks, kGradThreads, 0, s>>>(*st, (const double*)ctx, n_c, (const double*)mean,
        (const double*)var, ld, best_arm, min_arm, (double*)grad);
  else if (st->dtype == CRS_F32)
    min_zeta_grad_kernel<float><<<blocks, kGradThreads, 0, s>>>(*st, (const float*)ctx, n_c, (const float*)mean,
        (const float*)var, ld, best_arm, min_arm, (float*)grad);
  else
    return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "min_zeta_grad launch");
  return CRS_OK;
}

}  // namespace crs
