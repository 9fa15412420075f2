// crs_gp_prepare: per-arm exact-GP state (transforms, K^ = o*c(X~,X~) + noise*I, Cholesky, alpha,
// explicit L^-1) — the work GPyTorch caches after the first eval-mode `model.posterior` call
// (model family of contextual_rs/experiment_utils.py:29-40).
//
// One CTA per arm, everything in fp64 in shared memory, results cast to the compute dtype T.
// The grid kernel wants a multiply-only inner loop, so instead of L we store
//     linv_t[j][i] = o * (L^-1)[i][j]   (row j = column j of L^-1, zero above the diagonal;
//                                        stored in 32-column panels, see crs_common.cuh)
//     alpha[i]     = o * (K^^-1 (y~ - m))[i]
// which give  mu~ = m + c* . alpha  and  ||L^-1 k*||^2 = || sum_j linv_t[j][:] c*_j ||^2  for the
// unscaled correlation vector c* = c(x~*, X~).
#include <cstdio>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kPrepThreads = 256;
constexpr double kMinRange = 1e-8;  // botorch Normalize(min_range)
constexpr double kMinStdv = 1e-8;   // botorch Standardize(min_stdv)

__device__ __forceinline__ double group_sum(double v, double* red, int tid, int nt) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (nt >> 5); ++w) t += red[w];
  return t;
}

__host__ __device__ inline size_t prep_group_doubles(int ncap, int dp) {
  return (size_t)ncap * (ncap + 1) + (size_t)ncap * dp + ncap + 8 + 2 * CRS_MAX_DIM + 2;
}

// General variant: one 256-thread CTA per arm, any n <= CRS_MAX_OBS.
template <typename T>
__global__ void __launch_bounds__(kPrepThreads)
gp_prepare_kernel(crs_gp_state st, int arm_begin, int arm_end) {
  extern __shared__ double smem[];
  const int k = arm_begin + blockIdx.x;
  if (k >= arm_end) return;
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  const int d = st.dim, dp = st.dim_pad, ncap = st.n_cap;
  int n = st.n_obs[k];
  if (n > ncap) n = ncap;
  const int lda = ncap + 1;
  double* A = smem;                     // [ncap][lda] lower triangle
  double* xs = A + (size_t)ncap * lda;  // [ncap][dp]  normalised / lengthscale
  double* rhs = xs + (size_t)ncap * dp; // [ncap]
  double* red = rhs + ncap;             // [8] reduction scratch
  double* tr = red + 8;                 // [2*CRS_MAX_DIM] mins, ranges
  volatile int* failp = reinterpret_cast<volatile int*>(tr + 2 * CRS_MAX_DIM);
  if (tid == 0) *failp = 0;

  const double* X = st.train_x + (size_t)k * ncap * d;
  const double* Y = st.train_y + (size_t)k * ncap;

  // ---- Normalize bounds ------------------------------------------------------------------
  if (tid < d) {
    double mn, rg;
    if (st.flags & CRS_LEARN_NORMALIZE) {
      if (n > 0) {
        mn = X[tid];
        double mx = mn;
        for (int i = 1; i < n; ++i) {
          const double v = X[(size_t)i * d + tid];
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
        rg = mx - mn;
        if (rg < kMinRange) rg = 1.0;
      } else {
        mn = 0.0;
        rg = 1.0;
      }
      st.norm_mins[(size_t)k * d + tid] = mn;
      st.norm_ranges[(size_t)k * d + tid] = rg;
    } else {
      mn = st.norm_mins[(size_t)k * d + tid];
      rg = st.norm_ranges[(size_t)k * d + tid];
    }
    tr[tid] = mn;
    tr[CRS_MAX_DIM + tid] = rg;
  }
  __syncthreads();

  // ---- transformed inputs ----------------------------------------------------------------
  T* xn_out = reinterpret_cast<T*>(st.xn) + (size_t)k * ncap * d;
  T* xs_out = reinterpret_cast<T*>(st.xs) + (size_t)k * ncap * dp;
  for (int idx = tid; idx < ncap * dp; idx += nt) {
    const int i = idx / dp, dd = idx - i * dp;
    double v = 0.0;
    if (i < n && dd < d) {
      const double x = X[(size_t)i * d + dd];
      const double xn = (x - tr[dd]) / tr[CRS_MAX_DIM + dd];
      v = xn / st.lengthscale[(size_t)k * d + dd];
      // model.train_inputs: raw, or normalised in the model dtype's own arithmetic (eval-mode BoTorch >= 0.4)
      xn_out[(size_t)i * d + dd] = (st.flags & CRS_TRAIN_INPUTS_RAW) ? T(x) : (T(x) - T(tr[dd])) / T(tr[CRS_MAX_DIM + dd]);
    } else if (dd < d) {
      xn_out[(size_t)i * d + dd] = T(0);
    }
    xs[idx] = v;
    xs_out[idx] = T(v);
  }

  // ---- Standardize ------------------------------------------------------------------------
  double ybar, ysd;
  if (st.flags & CRS_LEARN_STANDARDIZE) {
    double s = 0.0;
    for (int i = tid; i < n; i += nt) s += Y[i];
    s = group_sum(s, red, tid, nt);
    ybar = n > 0 ? s / n : 0.0;
    double ss = 0.0;
    for (int i = tid; i < n; i += nt) {
      const double dv = Y[i] - ybar;
      ss += dv * dv;
    }
    ss = group_sum(ss, red, tid, nt);
    ysd = n > 1 ? sqrt(ss / (n - 1)) : 1.0;
    if (!(ysd >= kMinStdv)) ysd = 1.0;
    if (tid == 0) {
      st.y_mean[k] = ybar;
      st.y_std[k] = ysd;
    }
  } else {
    ybar = st.y_mean[k];
    ysd = st.y_std[k];
  }
  const double o = st.outputscale[k], noise = st.noise[k], m = st.mean_const[k];
  for (int i = tid; i < n; i += nt) rhs[i] = (Y[i] - ybar) / ysd - m;
  __syncthreads();

  // ---- K^ (lower triangle) -----------------------------------------------------------------
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    if (j <= i) {
      double r2 = 0.0;
      for (int dd = 0; dd < d; ++dd) {
        const double df = xs[i * dp + dd] - xs[j * dp + dd];
        r2 += df * df;
      }
      double v = o * correlation<double>(r2, st.kernel);
      if (i == j) v += noise;
      A[i * lda + j] = v;
    }
  }

  // ---- Cholesky, right-looking, in place ---------------------------------------------------
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const double ajj = A[j * lda + j];
    if (!(ajj > 0.0)) {
      if (tid == 0) *failp = j + 1;
      break;  // uniform: every thread read the same ajj
    }
    const double ljj = sqrt(ajj);
    __syncthreads();
    if (tid == 0) A[j * lda + j] = ljj;
    for (int i = j + 1 + tid; i < n; i += nt) A[i * lda + j] /= ljj;
    __syncthreads();
    const int mrem = n - j - 1;
    for (int idx = tid; idx < mrem * mrem; idx += nt) {
      const int i = j + 1 + idx / mrem, c = j + 1 + idx % mrem;
      if (c <= i) A[i * lda + c] -= A[i * lda + j] * A[c * lda + j];
    }
  }
  __syncthreads();
  if (*failp) {
    if (tid == 0) st.status[k] = *failp;
    return;
  }

  // ---- alpha = K^^-1 rhs: L z = rhs, L^T a = z ----------------------------------------------
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const double zj = rhs[j] / A[j * lda + j];
    __syncthreads();
    if (tid == 0) rhs[j] = zj;
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

  // ---- L <- L^-1 in place (column sweep from the right) -------------------------------------
  //   X[i][j] = -(sum_{c=j+1..i} X[i][c] L[c][j]) / L[j][j],  X[j][j] = 1 / L[j][j]
  for (int j = n - 1; j >= 0; --j) {
    const double inv = 1.0 / A[j * lda + j];
    double acc[(CRS_MAX_OBS + kPrepThreads - 1) / kPrepThreads];
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

  // ---- write the prepared state --------------------------------------------------------------
  T* alpha = reinterpret_cast<T*>(st.alpha) + (size_t)k * ncap;
  for (int i = tid; i < ncap; i += nt) alpha[i] = i < n ? T(o * rhs[i]) : T(0);
  // linv_t, panel-major (see crs_common.cuh): panel p holds rows j < min(ncap, 32 (p+1)) and
  // columns [32 p, 32 p + 32); entry (j, i) = o * (L^-1)[i][j] for i >= j, zero otherwise.
  T* lt = reinterpret_cast<T*>(st.linv_t) + (size_t)k * linv_stride(ncap, sizeof(T));
  const int ld = PanelLd<T>::value;
  for (int p = 0; p < num_panels(ncap); ++p) {
    T* pan = lt + panel_offset(ncap, p, sizeof(T));
    const int rows = panel_rows(ncap, p);
    for (int idx = tid; idx < rows * ld; idx += nt) {
      const int j = idx / ld, i = p * kPanel + (idx - j * ld);
      pan[idx] = (i >= j && i < n && i < (p + 1) * kPanel) ? T(o * A[i * lda + j]) : T(0);
    }
  }
  T* head = reinterpret_cast<T*>(st.arm_head) + (size_t)k * kHeadElems;
  if (tid < CRS_MAX_DIM) {
    if (tid < d) {
      head[kHeadSub + tid] = T(tr[tid]);
      head[kHeadMul + tid] = T(1.0 / (tr[CRS_MAX_DIM + tid] * st.lengthscale[(size_t)k * d + tid]));
    } else {
      head[kHeadSub + tid] = T(0);
      head[kHeadMul + tid] = T(0);
    }
  }
  if (tid == 0) {
    head[kHeadScal + 0] = T(o);
    head[kHeadScal + 1] = T(m);
    head[kHeadScal + 2] = T(ybar);
    head[kHeadScal + 3] = T(ysd);
    st.status[k] = 0;
  }
}

// Small-arm variant (row capacity <= 32, i.e. every BASELINE config): one 128-thread CTA per arm.
// The embarrassingly parallel phases (transforms, K^ fill, output) use all four warps; the
// sequential ones run in warp 0 alone with lane = row / column and warp shuffles, so the ~10 n
// CTA barriers of the general variant (ncu r01a: 52 % barrier stalls, 35 us) collapse to three:
//   Cholesky     left-looking, lane i owns row i:  s_i = A[i][j] - sum_{k<j} L[i][k] L[j][k]
//   alpha        column-oriented forward / backward substitution, r_i in a register per lane
//   L^-1         lane c solves L x = e_c for column c (no synchronisation at all)
#ifdef CRS_PREP_TIMING
__device__ long long g_prep_clk[8];
__device__ long long g_prep_start;
#define CRS_PREP_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_prep_clk[i] = clock64(); } while (0)
#else
#define CRS_PREP_MARK(i) do { } while (0)
#endif
constexpr int kSmallThreads = 128;
constexpr int kSmallCap = 32;

template <typename T>
__global__ void __launch_bounds__(kSmallThreads)
gp_prepare_small_kernel(crs_gp_state st, int arm_begin, int arm_end, const uint4* __restrict__ cp_src,
                        uint4* __restrict__ cp_dst, int cp_n16) {
  griddep_launch_dependents();  // the grid kernel may become resident now; it waits for this grid before reading
#ifdef CRS_PREP_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) g_prep_start = clock64();
#endif
  const int k = arm_begin + blockIdx.x;
  if (k >= arm_end) {
    // extra CTAs (crs_gao_decide_host): bring the caller's context set from mapped pinned host memory
    // into HBM while the arms are being prepared — the H2D copy of a decision overlaps prepare
    const int narms = arm_end - arm_begin;
    const int stride = ((int)gridDim.x - narms) * kSmallThreads;
    for (int i = ((int)blockIdx.x - narms) * kSmallThreads + (int)threadIdx.x; i < cp_n16; i += stride) cp_dst[i] = cp_src[i];
    return;
  }
  constexpr int lda = kSmallCap + 1;
  __shared__ double A[kSmallCap * lda];   // K^ then L (lower)
  __shared__ double B[kSmallCap * lda];   // L^-1 (lower)
  __shared__ double xs[kSmallCap * CRS_MAX_DIM];
  __shared__ double rhs[kSmallCap];
  __shared__ double tr[2 * CRS_MAX_DIM];
  __shared__ double s_stat[2];
  __shared__ int fail;
  const int tid = threadIdx.x, nt = kSmallThreads, lane = tid & 31, warp = tid >> 5;
  const int d = st.dim, dp = st.dim_pad, ncap = st.n_cap;
  const int n = min(st.n_obs[k], ncap);
  const double* X = st.train_x + (size_t)k * ncap * d;
  const double* Y = st.train_y + (size_t)k * ncap;
  if (tid == 0) fail = 0;
  // one coalesced round trip brings the arm's raw data on chip; every later phase reads shared
  // memory (the per-dimension min/max loop used to pay one global-memory latency per observation)
  __shared__ double s_x[kSmallCap * CRS_MAX_DIM];
  __shared__ double s_y[kSmallCap];
  __shared__ double s_dinv[kSmallCap];
  for (int idx = tid; idx < n * d; idx += nt) s_x[idx] = X[idx];
  if (tid < n) s_y[tid] = Y[tid];
  const double o = st.outputscale[k], noise = st.noise[k], m = st.mean_const[k];
  const double ls = tid < d ? st.lengthscale[(size_t)k * d + tid] : 1.0;
  __shared__ double s_ls[CRS_MAX_DIM];
  if (tid < d) s_ls[tid] = ls;
  __syncthreads();
  CRS_PREP_MARK(0);

  // ---- Normalize bounds (warp 0: lane = dimension) and Standardize constants (warp 1) ---------
  if (warp == 0 && lane < d) {
    double mn, rg;
    if (st.flags & CRS_LEARN_NORMALIZE) {
      if (n > 0) {
        mn = s_x[lane];
        double mx = mn;
        for (int i = 1; i < n; ++i) {
          const double v = s_x[i * d + lane];
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
        rg = mx - mn;
        if (rg < kMinRange) rg = 1.0;
      } else {
        mn = 0.0;
        rg = 1.0;
      }
      st.norm_mins[(size_t)k * d + lane] = mn;
      st.norm_ranges[(size_t)k * d + lane] = rg;
    } else {
      mn = st.norm_mins[(size_t)k * d + lane];
      rg = st.norm_ranges[(size_t)k * d + lane];
    }
    tr[lane] = mn;
    tr[CRS_MAX_DIM + lane] = rg;
  }
  if (warp == 1) {
    double ybar, ysd;
    if (st.flags & CRS_LEARN_STANDARDIZE) {
      const double y = lane < n ? s_y[lane] : 0.0;
      ybar = n > 0 ? warp_sum(y) / n : 0.0;
      const double dv = lane < n ? y - ybar : 0.0;
      const double ss = warp_sum(dv * dv);
      ysd = n > 1 ? sqrt(ss / (n - 1)) : 1.0;
      if (!(ysd >= kMinStdv)) ysd = 1.0;
      if (lane == 0) {
        st.y_mean[k] = ybar;
        st.y_std[k] = ysd;
      }
    } else {
      ybar = st.y_mean[k];
      ysd = st.y_std[k];
    }
    if (lane == 0) {
      s_stat[0] = ybar;
      s_stat[1] = ysd;
    }
  }
  __syncthreads();
  CRS_PREP_MARK(1);
  const double ybar = s_stat[0], ysd = s_stat[1];

  // ---- transformed inputs ----------------------------------------------------------------
  T* xn_out = reinterpret_cast<T*>(st.xn) + (size_t)k * ncap * d;
  T* xs_out = reinterpret_cast<T*>(st.xs) + (size_t)k * ncap * dp;
  for (int idx = tid; idx < ncap * dp; idx += nt) {
    const int i = idx / dp, dd = idx - i * dp;
    double v = 0.0;
    if (i < n && dd < d) {
      const double x = s_x[i * d + dd];
      const double xn = (x - tr[dd]) / tr[CRS_MAX_DIM + dd];
      v = xn / s_ls[dd];
      xn_out[(size_t)i * d + dd] = (st.flags & CRS_TRAIN_INPUTS_RAW) ? T(x) : (T(x) - T(tr[dd])) / T(tr[CRS_MAX_DIM + dd]);
    } else if (dd < d) {
      xn_out[(size_t)i * d + dd] = T(0);
    }
    xs[i * CRS_MAX_DIM + dd] = v;
    xs_out[idx] = T(v);
  }
  if (tid < n) rhs[tid] = (s_y[tid] - ybar) / ysd - m;
  for (int idx = tid; idx < kSmallCap * lda; idx += nt) B[idx] = 0.0;
  __syncthreads();
  CRS_PREP_MARK(2);

  // ---- K^ (lower triangle), all warps --------------------------------------------------------
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    if (j <= i) {
      double r2 = 0.0;
      for (int dd = 0; dd < d; ++dd) {
        const double df = xs[i * CRS_MAX_DIM + dd] - xs[j * CRS_MAX_DIM + dd];
        r2 += df * df;
      }
      double v = o * correlation<double>(r2, st.kernel);
      if (i == j) v += noise;
      A[i * lda + j] = v;
    }
  }
  __syncthreads();
  CRS_PREP_MARK(3);

  // ---- factorisation: warp 0 alone -----------------------------------------------------------
  if (warp == 0) {
    const bool row = lane < n;
    int bad = 0;
    double dinv = 0.0;  // lane j keeps 1 / L[j][j]: the substitutions multiply instead of divide
    for (int j = 0; j < n; ++j) {
      double s0 = (row && lane >= j) ? A[lane * lda + j] : 0.0, s1 = 0.0;
      if (row && lane >= j) {
        int c = 0;
        for (; c + 1 < j; c += 2) {  // two independent chains: the step is DFMA-latency bound
          s0 = fma(-A[lane * lda + c], A[j * lda + c], s0);
          s1 = fma(-A[lane * lda + c + 1], A[j * lda + c + 1], s1);
        }
        if (c < j) s0 = fma(-A[lane * lda + c], A[j * lda + c], s0);
      }
      const double s = s0 + s1;
      const double sjj = __shfl_sync(kFull, s, j);
      if (!(sjj > 0.0)) {
        bad = j + 1;
        break;  // uniform: every lane sees the same pivot
      }
      const double rs = rsqrt(sjj);
      if (lane == j) dinv = rs;
      if (row && lane >= j) A[lane * lda + j] = lane == j ? sjj * rs : s * rs;
      __syncwarp();
    }
    if (bad && lane == 0) fail = bad;
    s_dinv[lane] = dinv;
  }
  __syncthreads();
  CRS_PREP_MARK(4);
  if (fail) {
    if (tid == 0) st.status[k] = fail;
    return;
  }
  // ---- alpha (warp 1) and L^-1 (warp 0) side by side: both only read L ---------------------------
  if (warp == 1) {
    const bool row = lane < n;
    const double dinv = s_dinv[lane];
    double r = row ? rhs[lane] : 0.0;
    for (int j = 0; j < n; ++j) {  // L z = rhs
      const double zj = __shfl_sync(kFull, r * dinv, j);
      if (lane == j) r = zj;
      else if (row && lane > j) r = fma(-A[lane * lda + j], zj, r);
    }
    for (int j = n - 1; j >= 0; --j) {  // L^T a = z
      const double aj = __shfl_sync(kFull, r * dinv, j);
      if (lane == j) r = aj;
      else if (lane < j) r = fma(-A[j * lda + lane], aj, r);
    }
    if (row) rhs[lane] = r;
  } else if (warp == 0) {
    // column `lane` of X = L^-1 by forward substitution, the column in registers: x[c] = 0 for c < lane
    // (above the diagonal), so the sums need no lane-dependent bounds; L[i][c] is a broadcast load.
    // (clock64, n = 20: 615 cycles per row with the column re-read from shared memory -> see above.)
    double x[kSmallCap];
#pragma unroll
    for (int i = 0; i < kSmallCap; ++i) {
      x[i] = 0.0;
      if (i < n) {  // uniform
        double a0 = i == lane ? 1.0 : 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int c = 0; c < i; ++c) {
          const double l = A[i * lda + c];
          if ((c & 3) == 0) a0 = fma(-l, x[c], a0);
          else if ((c & 3) == 1) a1 = fma(-l, x[c], a1);
          else if ((c & 3) == 2) a2 = fma(-l, x[c], a2);
          else a3 = fma(-l, x[c], a3);
        }
        x[i] = ((a0 + a1) + (a2 + a3)) * s_dinv[i];
        if (lane < n) B[i * lda + lane] = x[i];
      }
    }
  }
  __syncthreads();
  CRS_PREP_MARK(5);
  if (fail) {
    if (tid == 0) st.status[k] = fail;
    return;
  }

  // ---- write the prepared state, all warps -----------------------------------------------------
  T* alpha = reinterpret_cast<T*>(st.alpha) + (size_t)k * ncap;
  for (int i = tid; i < ncap; i += nt) alpha[i] = i < n ? T(o * rhs[i]) : T(0);
  T* lt = reinterpret_cast<T*>(st.linv_t) + (size_t)k * linv_stride(ncap, sizeof(T));
  constexpr int ld = PanelLd<T>::value;
  for (int idx = tid; idx < ncap * ld; idx += nt) {  // n_cap <= 32: a single panel
    const int j = idx / ld, i = idx - j * ld;
    lt[idx] = (i >= j && i < n) ? T(o * B[i * lda + j]) : T(0);
  }
  T* head = reinterpret_cast<T*>(st.arm_head) + (size_t)k * kHeadElems;
  if (tid < CRS_MAX_DIM) {
    if (tid < d) {
      head[kHeadSub + tid] = T(tr[tid]);
      head[kHeadMul + tid] = T(1.0 / (tr[CRS_MAX_DIM + tid] * s_ls[tid]));
    } else {
      head[kHeadSub + tid] = T(0);
      head[kHeadMul + tid] = T(0);
    }
  }
  if (tid == 0) {
    head[kHeadScal + 0] = T(o);
    head[kHeadScal + 1] = T(m);
    head[kHeadScal + 2] = T(ybar);
    head[kHeadScal + 3] = T(ysd);
    st.status[k] = 0;
#ifdef CRS_PREP_TIMING
    if (blockIdx.x == 0) {
      const long long t6 = clock64();
      printf("prepare phases (cycles): load %lld | bounds %lld | transforms %lld | Khat %lld | cholesky %lld | alpha+Linv %lld | write %lld\n",
             g_prep_clk[0] - g_prep_start, g_prep_clk[1] - g_prep_clk[0], g_prep_clk[2] - g_prep_clk[1], g_prep_clk[3] - g_prep_clk[2],
             g_prep_clk[4] - g_prep_clk[3], g_prep_clk[5] - g_prep_clk[4], t6 - g_prep_clk[5]);
    }
#endif
  }
}

}  // namespace

int launch_gp_prepare(const crs_gp_state* st, int arm_begin, int arm_end, cudaStream_t s) {
  return launch_gp_prepare_copy(st, arm_begin, arm_end, nullptr, nullptr, 0, s);
}

// cp_src != nullptr: also copy cp_bytes (multiple of 16) from device-visible cp_src (mapped pinned host
// memory) to cp_dst, inside the prepare launch when the small-arm kernel is used, else by a plain copy.
int launch_gp_prepare_copy(const crs_gp_state* st, int arm_begin, int arm_end, const void* cp_src, void* cp_dst,
                           size_t cp_bytes, cudaStream_t s) {
  if (!st || arm_begin < 0 || arm_end > st->num_arms || arm_begin > arm_end) return CRS_ERR_INVALID_ARG;
  if (cp_src && (!cp_dst || (cp_bytes & 15) || arm_begin == arm_end)) return CRS_ERR_INVALID_ARG;
  if (st->dim < 1 || st->dim > CRS_MAX_DIM || st->dim_pad != dim_pad_of(st->dim)) return CRS_ERR_UNSUPPORTED;
  if (st->n_cap < 1 || st->n_cap > CRS_MAX_OBS || (st->n_cap & 7)) return CRS_ERR_UNSUPPORTED;
  if (st->dtype != CRS_F64 && st->dtype != CRS_F32) return CRS_ERR_INVALID_ARG;
  if (arm_begin == arm_end) return CRS_OK;
  const int narms = arm_end - arm_begin;
  if (st->n_cap <= kSmallCap) {
    const int n16 = cp_src ? (int)(cp_bytes >> 4) : 0;
    int cp_blocks = (n16 + 2 * kSmallThreads - 1) / (2 * kSmallThreads);
    if (cp_blocks > 48) cp_blocks = 48;
    const auto* src = reinterpret_cast<const uint4*>(cp_src);
    auto* dst = reinterpret_cast<uint4*>(cp_dst);
    if (st->dtype == CRS_F64) gp_prepare_small_kernel<double><<<narms + cp_blocks, kSmallThreads, 0, s>>>(*st, arm_begin, arm_end, src, dst, n16);
    else gp_prepare_small_kernel<float><<<narms + cp_blocks, kSmallThreads, 0, s>>>(*st, arm_begin, arm_end, src, dst, n16);
    CRS_CUDA_TRY(cudaGetLastError(), "gp_prepare launch");
    return CRS_OK;
  }
  if (cp_src) CRS_CUDA_TRY(cudaMemcpyAsync(cp_dst, cp_src, cp_bytes, cudaMemcpyDefault, s), "ctx copy");
  const size_t smem = prep_group_doubles(st->n_cap, st->dim_pad) * sizeof(double);
  auto go = [&](auto kern) -> int {
    if (smem > 40 * 1024)
      CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "prepare attr");
    kern<<<narms, kPrepThreads, smem, s>>>(*st, arm_begin, arm_end);
    CRS_CUDA_TRY(cudaGetLastError(), "gp_prepare launch");
    return CRS_OK;
  };
  if (st->dtype == CRS_F64) return go(gp_prepare_kernel<double>);
  return go(gp_prepare_kernel<float>);
}

}  // namespace crs
