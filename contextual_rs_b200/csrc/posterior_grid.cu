// crs_posterior_grid (K1): the ModelListGP posterior mean / marginal variance over the full
// arm x context grid — replaces `model.posterior(context_set)` at
// contextual_rs/contextual_rs_strategies.py:133-136 (and generalized_pcs.py:443-444,
// continuous_context.py:68,96, experiments/wsc_experiments/main.py:460).
//
// Mapping: CTA = (tile of TB contexts) x (tile of AT arms); thread = one context.  The arm's
// prepared state sits in shared memory and is read by warp-wide broadcast.  Per (context, arm):
//   phase A  c_j = corr(|| x~* - X~_j ||^2), j = 0..n-1   (direct differences; two observations
//            per iteration for ILP), macc += c_j alpha_j, c_j parked in shared memory [j][thread]
//   phase B  v = (o L^-1) c in register chunks of VB rows: v[i] += linv_t[j][i] c_j for i >= j,
//            the matching column panel of linv_t staged in shared memory; q += sum v_i^2
//   mu = ybar + s (m + macc),  var = max(s^2 (o - q), floor)
// The kernel is FP64/FP32-pipe bound (~55 flop per output byte); HBM traffic is the
// context tile in and the two tables out.  Tables are [n_c, ld] arm-fastest, so each thread buffers
// its AT results in shared memory and writes them as one contiguous run.
//
// fp64 transcendental cost dominates at n ~ 20, so exp() is a 16-entry-table + degree-7 polynomial
// (12 FP64 ops, coefficients in constant memory so that no uniform-register moves are issued) and
// sqrt() is rsqrt.approx + two coupled Newton steps + one correction (branch-free).
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kGridThreads = 128;
constexpr int kMaxArmTile = 8;

// ---- fp64 exp(x), x <= 0: x = (16 k + i) ln2/16 + r, exp(x) = 2^k * 2^(i/16) * e^r ------------
__constant__ double kExpTable[16] = {
    1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721,
    1.241857812073484, 1.2968395546510096, 1.3542555469368927, 1.4142135623730951,
    1.4768261459394993, 1.5422108254079407, 1.6104903319492543, 1.681792830507429,
    1.7562521603732995, 1.8340080864093424, 1.9152065613971474};
__constant__ double kExpPoly[6] = {  // 1/7!, 1/6!, 1/5!, 1/4!, 1/3!, 1/2!
    1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};

__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ s_tab) {
  // t = x * 16/ln2 + 1.5*2^52: the integer n = 16 k + i lands in the low mantissa bits of t
  const double t = fma(x, 23.083120654223414, 6755399441055744.0);
  const double nd = t - 6755399441055744.0;
  double r = fma(nd, -0.043321698784996581, x);        // ln2/16 (high part)
  r = fma(nd, -1.4494042586539372e-18, r);             // ln2/16 (low part)
  int n = __double2loint(t);
  n = max(n, -16000);                                   // exp(-693) ~ 1e-301: clamp instead of underflow
  double p = kExpPoly[0];
  p = fma(p, r, kExpPoly[1]);
  p = fma(p, r, kExpPoly[2]);
  p = fma(p, r, kExpPoly[3]);
  p = fma(p, r, kExpPoly[4]);
  p = fma(p, r, kExpPoly[5]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double y = s_tab[n & 15] * p;
  return __hiloint2double(__double2hiint(y) + ((n >> 4) << 20), __double2loint(y));
}

// sqrt(x) for x >= 0 (x = 0 returns ~1e-145, harmless: it only enters s and exp(-s)).
__device__ __forceinline__ double sqrt_pos(double x) {
  x = fmax(x, 1e-290);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double e = fma(-g, g, x);
  return fma(e, h, g);
}

template <typename T> struct Corr;
template <> struct Corr<double> {
  static __device__ __forceinline__ double eval(double r2, int kernel, const double* s_tab) {
    if (kernel == CRS_MATERN52) {
      const double s = sqrt_pos(5.0 * r2);                       // sqrt5 * r
      const double poly = fma(s, fma(s, 1.0 / 3.0, 1.0), 1.0);  // 1 + s + s^2/3
      return poly * exp_neg(-s, s_tab);
    }
    return exp_neg(-0.5 * r2, s_tab);
  }
};
template <> struct Corr<float> {
  static __device__ __forceinline__ float eval(float r2, int kernel, const float*) {
    return correlation<float>(r2, kernel);
  }
};

template <typename T, int D, int VB>
__global__ void __launch_bounds__(kGridThreads)
posterior_grid_kernel(crs_gp_state st, const T* __restrict__ ctx, int64_t n_c, int arm_begin,
                      int arm_end, int arm_tile, int n_rows, T* __restrict__ mean,
                      T* __restrict__ var, int64_t ld, int col_offset) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int ncap = st.n_cap, d = st.dim;
  // shared layout (T): panel [n_rows][VB] | xs [n_rows][D] | alpha [n_rows] | sub [D] | mul [D] |
  //                    tab [16] | c [n_rows][TB] | out [2][arm_tile][TB]
  T* s_panel = reinterpret_cast<T*>(smem_raw);
  T* s_xs = s_panel + n_rows * VB;
  T* s_alpha = s_xs + n_rows * D;
  T* s_sub = s_alpha + n_rows;
  T* s_mul = s_sub + D;
  T* s_tab = s_mul + D;
  T* s_c = s_tab + 16;
  T* s_out = s_c + n_rows * kGridThreads;

  if (tid < 16) s_tab[tid] = (T)kExpTable[tid];

  const int64_t c = (int64_t)blockIdx.x * kGridThreads + tid;
  const bool live = c < n_c;
  const int a0 = arm_begin + blockIdx.y * arm_tile;
  const int a1 = min(a0 + arm_tile, arm_end);

  T xraw[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) xraw[dd] = (live && dd < d) ? ctx[c * d + dd] : T(0);

  for (int k = a0; k < a1; ++k) {
    const int n = min(min(st.n_obs[k], ncap), n_rows);
    const int n2 = (n + 1) & ~1;
    const T* g_lt = reinterpret_cast<const T*>(st.linv_t) + (size_t)k * ncap * ncap;
    __syncthreads();  // previous arm's readers are done with the shared state
    {
      const T* g_xs = reinterpret_cast<const T*>(st.xs) + (size_t)k * ncap * D;
      for (int idx = tid; idx < n2 * D; idx += kGridThreads) s_xs[idx] = idx < n * D ? g_xs[idx] : T(0);
      const T* g_al = reinterpret_cast<const T*>(st.alpha) + (size_t)k * ncap;
      for (int idx = tid; idx < n2; idx += kGridThreads) s_alpha[idx] = idx < n ? g_al[idx] : T(0);
      if (tid < D) {
        s_sub[tid] = reinterpret_cast<const T*>(st.ctx_sub)[(size_t)k * D + tid];
        s_mul[tid] = reinterpret_cast<const T*>(st.ctx_mul)[(size_t)k * D + tid];
      }
      // panel of chunk 0 (the only one when n <= VB)
      const int rows0 = min(n, VB);
      for (int idx = tid; idx < rows0 * VB; idx += kGridThreads) {
        const int j = idx / VB, i = idx - j * VB;
        s_panel[idx] = i < ncap ? g_lt[(size_t)j * ncap + i] : T(0);
      }
    }
    __syncthreads();

    // ---- phase A: correlations, two observations per iteration ------------------------------
    T xq[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) xq[dd] = (xraw[dd] - s_sub[dd]) * s_mul[dd];
    T macc0 = T(0), macc1 = T(0);
    for (int j = 0; j < n2; j += 2) {
      T ra = T(0), rb = T(0);
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        const T da = xq[dd] - s_xs[j * D + dd];
        const T db = xq[dd] - s_xs[(j + 1) * D + dd];
        ra = fma(da, da, ra);
        rb = fma(db, db, rb);
      }
      const T ca = Corr<T>::eval(ra, st.kernel, s_tab);
      const T cb = Corr<T>::eval(rb, st.kernel, s_tab);
      macc0 = fma(ca, s_alpha[j], macc0);
      macc1 = fma(cb, s_alpha[j + 1], macc1);   // alpha is zero in the padded row
      s_c[j * kGridThreads + tid] = ca;
      s_c[(j + 1) * kGridThreads + tid] = cb;
    }
    // (s_c is private per thread column: no barrier needed before phase B reads it back)

    // ---- phase B: v = (o L^-1) c, VB rows at a time -------------------------------------------
    T q = T(0);
    for (int p0 = 0; p0 < n; p0 += VB) {
      const int jend = min(n, p0 + VB);  // rows j <= last row of this chunk contribute
      if (p0 > 0) {
        __syncthreads();
        for (int idx = tid; idx < jend * VB; idx += kGridThreads) {
          const int j = idx / VB, i = p0 + idx - j * VB;
          s_panel[idx] = i < ncap ? g_lt[(size_t)j * ncap + i] : T(0);
        }
        __syncthreads();
      }
      T v[VB];
#pragma unroll
      for (int i = 0; i < VB; ++i) v[i] = T(0);
      for (int j = 0; j < jend; ++j) {
        const T cj = s_c[j * kGridThreads + tid];
        const T* row = s_panel + j * VB;
#pragma unroll
        for (int ib = 0; ib < VB / 8; ++ib) {
          if (p0 + ib * 8 + 7 >= j) {  // warp-uniform: entries above the diagonal are zero
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) v[ib * 8 + ii] = fma(row[ib * 8 + ii], cj, v[ib * 8 + ii]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VB; ++i) q = fma(v[i], v[i], q);
    }

    const T* sc = reinterpret_cast<const T*>(st.arm_scal) + (size_t)k * 4;
    const T o = sc[0], m = sc[1], ybar = sc[2], ysd = sc[3];
    const T mu = fma(ysd, m + (macc0 + macc1), ybar);
    T vr = (ysd * ysd) * (o - q);
    vr = vr > variance_floor<T>() ? vr : variance_floor<T>();  // also maps NaN to the floor
    s_out[(k - a0) * kGridThreads + tid] = mu;
    s_out[(arm_tile + (k - a0)) * kGridThreads + tid] = vr;
  }

  if (live) {
    const int na = a1 - a0;
    T* mrow = mean + c * ld + col_offset + (a0 - arm_begin);
    T* vrow = var + c * ld + col_offset + (a0 - arm_begin);
    for (int a = 0; a < na; ++a) {
      mrow[a] = s_out[a * kGridThreads + tid];
      vrow[a] = s_out[(arm_tile + a) * kGridThreads + tid];
    }
  }
}

template <typename T, int D, int VB>
int launch_variant(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                   int arm_end, int n_rows, void* mean, void* var, int64_t ld, int col_offset,
                   cudaStream_t s) {
  const int narms = arm_end - arm_begin;
  const int64_t ctx_tiles = (n_c + kGridThreads - 1) / kGridThreads;
  // Arm tile: as large as possible (contiguous output runs, context tile reused) while keeping
  // >= ~4 CTAs per SM in flight on 148 SMs for small problems.
  int arm_tile = kMaxArmTile;
  while (arm_tile > 1 && ctx_tiles * ((narms + arm_tile - 1) / arm_tile) < 148 * 4) arm_tile >>= 1;
  auto smem_for = [&](int at) {
    return sizeof(T) * ((size_t)n_rows * VB + (size_t)n_rows * D + n_rows + 2 * D + 16 +
                        (size_t)n_rows * kGridThreads + 2 * (size_t)at * kGridThreads);
  };
  while (arm_tile > 1 && smem_for(arm_tile) > 227 * 1024) arm_tile >>= 1;
  const size_t smem = smem_for(arm_tile);
  if (smem > 227 * 1024) return CRS_ERR_UNSUPPORTED;
  const int arm_tiles = (narms + arm_tile - 1) / arm_tile;
  if (arm_tiles > 65535) return CRS_ERR_UNSUPPORTED;
  auto kern = posterior_grid_kernel<T, D, VB>;
  if (smem > 48 * 1024)
    CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "grid attr");
  dim3 grid((unsigned)ctx_tiles, (unsigned)arm_tiles);
  kern<<<grid, kGridThreads, smem, s>>>(*st, reinterpret_cast<const T*>(ctx), n_c, arm_begin,
                                         arm_end, arm_tile, n_rows, reinterpret_cast<T*>(mean),
                                         reinterpret_cast<T*>(var), ld, col_offset);
  CRS_CUDA_TRY(cudaGetLastError(), "posterior_grid launch");
  return CRS_OK;
}

template <typename T, int VB>
int dispatch_dim(const crs_gp_state* st, const void* ctx, int64_t n_c, int a0, int a1, int n_rows,
                 void* mean, void* var, int64_t ld, int col, cudaStream_t s) {
  switch (st->dim_pad) {
    case 1: return launch_variant<T, 1, VB>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 2: return launch_variant<T, 2, VB>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 4: return launch_variant<T, 4, VB>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 8: return launch_variant<T, 8, VB>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 16: return launch_variant<T, 16, VB>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  }
  return CRS_ERR_UNSUPPORTED;
}

template <typename T>
int dispatch_vb(const crs_gp_state* st, const void* ctx, int64_t n_c, int a0, int a1, void* mean,
                void* var, int64_t ld, int col, cudaStream_t s) {
  // n_max: the caller's bound on n_obs (0 = unknown -> the capacity).  Rows are padded to even.
  int n_max = st->n_max > 0 && st->n_max <= st->n_cap ? st->n_max : st->n_cap;
  const int n_rows = (n_max + 1) & ~1;
  if (n_max <= 8) return dispatch_dim<T, 8>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  if (n_max <= 16) return dispatch_dim<T, 16>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  if (n_max <= 24) return dispatch_dim<T, 24>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  return dispatch_dim<T, 32>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
}

}  // namespace

int launch_posterior_grid(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                          int arm_end, void* mean, void* var, int64_t ld, int col_offset,
                          cudaStream_t s) {
  if (!st || !ctx || !mean || !var || n_c < 0 || arm_begin < 0 || arm_end > st->num_arms ||
      arm_begin > arm_end || ld < col_offset + (arm_end - arm_begin))
    return CRS_ERR_INVALID_ARG;
  if (st->n_cap > CRS_MAX_OBS || st->dim_pad != dim_pad_of(st->dim)) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0 || arm_begin == arm_end) return CRS_OK;
  if (st->dtype == CRS_F64) return dispatch_vb<double>(st, ctx, n_c, arm_begin, arm_end, mean, var, ld, col_offset, s);
  if (st->dtype == CRS_F32) return dispatch_vb<float>(st, ctx, n_c, arm_begin, arm_end, mean, var, ld, col_offset, s);
  return CRS_ERR_INVALID_ARG;
}

}  // namespace crs
