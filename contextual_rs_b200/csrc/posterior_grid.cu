// crs_posterior_grid (K1): the ModelListGP posterior mean / marginal variance over the full
// arm x context grid — replaces `model.posterior(context_set)` at
// contextual_rs/contextual_rs_strategies.py:133-136 (and generalized_pcs.py:443-444,
// continuous_context.py:68,96, experiments/wsc_experiments/main.py:460).
//
// Mapping: CTA = (tile of TB contexts) x (tile of AT arms); thread = one context.  The arm's
// prepared state (scaled train inputs, alpha, o*L^-T) sits in shared memory and is read by
// warp-wide broadcast; per context the thread streams over the n observations j:
//     c_j   = corr(|| x~* - X~_j ||^2)                  (direct differences, no cancellation)
//     macc += c_j * alpha_j
//     v[i] += linv_t[j][i] * c_j      for i >= j        (register accumulators, 8-wide blocks)
// and finishes with  mu = ybar + s (m + macc),  var = max(s^2 (o - sum v_i^2), floor).
// The kernel is FP64/FP32-pipe bound (~55 flop per output byte); the only HBM
// traffic is the context tile in and the two tables out.  Tables are [n_c, ld] arm-fastest, so
// each thread buffers its AT results in shared memory and writes them as one contiguous run.
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kGridThreads = 128;
constexpr int kMaxArmTile = 8;

template <typename T, int D, int NB>
__global__ void __launch_bounds__(kGridThreads)
posterior_grid_kernel(crs_gp_state st, const T* __restrict__ ctx, int64_t n_c, int arm_begin,
                      int arm_end, int arm_tile, T* __restrict__ mean, T* __restrict__ var,
                      int64_t ld, int col_offset) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int ncap = st.n_cap, d = st.dim;
  // shared layout (T): lt [NB][NB] | xs [NB][D] | alpha [NB] | sub [D] | mul [D] | out [2][AT][TB]
  T* s_lt = reinterpret_cast<T*>(smem_raw);
  T* s_xs = s_lt + NB * NB;
  T* s_alpha = s_xs + NB * D;
  T* s_sub = s_alpha + NB;
  T* s_mul = s_sub + D;
  T* s_out = s_mul + D;

  const int64_t c = (int64_t)blockIdx.x * kGridThreads + tid;
  const bool live = c < n_c;
  const int a0 = arm_begin + blockIdx.y * arm_tile;
  const int a1 = min(a0 + arm_tile, arm_end);

  T xraw[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) xraw[dd] = (live && dd < d) ? ctx[c * d + dd] : T(0);

  for (int k = a0; k < a1; ++k) {
    const int n = min(st.n_obs[k], ncap);
    __syncthreads();  // previous arm's readers are done with the shared state
    {
      const T* g_lt = reinterpret_cast<const T*>(st.linv_t) + (size_t)k * ncap * ncap;
      for (int idx = tid; idx < n * NB; idx += kGridThreads) {
        const int j = idx / NB, i = idx - j * NB;
        s_lt[idx] = i < ncap ? g_lt[(size_t)j * ncap + i] : T(0);
      }
      const T* g_xs = reinterpret_cast<const T*>(st.xs) + (size_t)k * ncap * D;
      for (int idx = tid; idx < n * D; idx += kGridThreads) s_xs[idx] = g_xs[idx];
      const T* g_al = reinterpret_cast<const T*>(st.alpha) + (size_t)k * ncap;
      for (int idx = tid; idx < n; idx += kGridThreads) s_alpha[idx] = g_al[idx];
      if (tid < D) {
        s_sub[tid] = reinterpret_cast<const T*>(st.ctx_sub)[(size_t)k * D + tid];
        s_mul[tid] = reinterpret_cast<const T*>(st.ctx_mul)[(size_t)k * D + tid];
      }
    }
    __syncthreads();

    T xq[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) xq[dd] = (xraw[dd] - s_sub[dd]) * s_mul[dd];

    T v[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) v[i] = T(0);
    T macc = T(0);

    for (int j = 0; j < n; ++j) {
      T r2 = T(0);
#pragma unroll
      for (int dd = 0; dd < D; ++dd) {
        const T df = xq[dd] - s_xs[j * D + dd];
        r2 = fma(df, df, r2);
      }
      const T cj = correlation<T>(r2, st.kernel);
      macc = fma(cj, s_alpha[j], macc);
      const T* row = s_lt + j * NB;
#pragma unroll
      for (int ib = 0; ib < NB / 8; ++ib) {
        if (ib * 8 + 7 >= j) {  // warp-uniform: rows above the diagonal block are zero
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) v[ib * 8 + ii] = fma(row[ib * 8 + ii], cj, v[ib * 8 + ii]);
        }
      }
    }
    T q = T(0);
#pragma unroll
    for (int i = 0; i < NB; ++i) q = fma(v[i], v[i], q);

    const T* sc = reinterpret_cast<const T*>(st.arm_scal) + (size_t)k * 4;
    const T o = sc[0], m = sc[1], ybar = sc[2], ysd = sc[3];
    const T mu = fma(ysd, m + macc, ybar);
    T vr = (ysd * ysd) * (o - q);
    vr = vr > variance_floor<T>() ? vr : variance_floor<T>();  // also maps NaN to the floor
    s_out[(k - a0) * kGridThreads + tid] = mu;
    s_out[(kMaxArmTile + (k - a0)) * kGridThreads + tid] = vr;
  }

  if (live) {
    const int na = a1 - a0;
    T* mrow = mean + c * ld + col_offset + (a0 - arm_begin);
    T* vrow = var + c * ld + col_offset + (a0 - arm_begin);
    for (int a = 0; a < na; ++a) {
      mrow[a] = s_out[a * kGridThreads + tid];
      vrow[a] = s_out[(kMaxArmTile + a) * kGridThreads + tid];
    }
  }
}

template <typename T, int D, int NB>
int launch_variant(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                   int arm_end, void* mean, void* var, int64_t ld, int col_offset,
                   cudaStream_t s) {
  const int narms = arm_end - arm_begin;
  const int64_t ctx_tiles = (n_c + kGridThreads - 1) / kGridThreads;
  // Arm tile: as large as possible (contiguous output runs, context tile reused) while keeping
  // >= ~4 CTAs per SM in flight on 148 SMs for small problems.
  int arm_tile = kMaxArmTile;
  while (arm_tile > 1 && ctx_tiles * ((narms + arm_tile - 1) / arm_tile) < 148 * 4) arm_tile >>= 1;
  const int arm_tiles = (narms + arm_tile - 1) / arm_tile;
  if (arm_tiles > 65535) return CRS_ERR_UNSUPPORTED;
  const size_t smem = sizeof(T) * ((size_t)NB * NB + NB * D + NB + 2 * D + 2 * kMaxArmTile * kGridThreads);
  auto kern = posterior_grid_kernel<T, D, NB>;
  if (smem > 48 * 1024)
    CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "grid attr");
  dim3 grid((unsigned)ctx_tiles, (unsigned)arm_tiles);
  kern<<<grid, kGridThreads, smem, s>>>(*st, reinterpret_cast<const T*>(ctx), n_c, arm_begin,
                                         arm_end, arm_tile, reinterpret_cast<T*>(mean),
                                         reinterpret_cast<T*>(var), ld, col_offset);
  CRS_CUDA_TRY(cudaGetLastError(), "posterior_grid launch");
  return CRS_OK;
}

template <typename T, int NB>
int dispatch_dim(const crs_gp_state* st, const void* ctx, int64_t n_c, int a0, int a1, void* mean,
                 void* var, int64_t ld, int col, cudaStream_t s) {
  switch (st->dim_pad) {
    case 1: return launch_variant<T, 1, NB>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
    case 2: return launch_variant<T, 2, NB>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
    case 4: return launch_variant<T, 4, NB>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
    case 8: return launch_variant<T, 8, NB>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
    case 16: return launch_variant<T, 16, NB>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
  }
  return CRS_ERR_UNSUPPORTED;
}

template <typename T>
int dispatch_nb(const crs_gp_state* st, const void* ctx, int64_t n_c, int a0, int a1, void* mean,
                void* var, int64_t ld, int col, cudaStream_t s) {
  if (st->n_cap <= 32) return dispatch_dim<T, 32>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
  if (st->n_cap <= 64) return dispatch_dim<T, 64>(st, ctx, n_c, a0, a1, mean, var, ld, col, s);
  return CRS_ERR_UNSUPPORTED;  // TODO(next): shared-memory-resident k* path for n_cap > 64
}

}  // namespace

int launch_posterior_grid(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                          int arm_end, void* mean, void* var, int64_t ld, int col_offset,
                          cudaStream_t s) {
  if (!st || !ctx || !mean || !var || n_c < 0 || arm_begin < 0 || arm_end > st->num_arms ||
      arm_begin > arm_end || ld < col_offset + (arm_end - arm_begin))
    return CRS_ERR_INVALID_ARG;
  if (n_c == 0 || arm_begin == arm_end) return CRS_OK;
  if (st->dtype == CRS_F64) return dispatch_nb<double>(st, ctx, n_c, arm_begin, arm_end, mean, var, ld, col_offset, s);
  if (st->dtype == CRS_F32) return dispatch_nb<float>(st, ctx, n_c, arm_begin, arm_end, mean, var, ld, col_offset, s);
  return CRS_ERR_INVALID_ARG;
}

}  // namespace crs
