// crs_ocba_scan / crs_ocba_resolve_tie / crs_ocba_select (K2): the GP-C-OCBA allocation on a
// materialised posterior grid — replaces contextual_rs/contextual_rs_strategies.py:137-202 and
// contextual_rs/continuous_context.py:21-36, 57-74, 96-132.
//
// The reference sorts every row of the n_c x K table, gathers variances, builds five n_c x (K-1)
// temporaries and arg-mins the flattened zeta.  Only the arg-max arm and the zeta minimum of each
// row are needed, so one warp owns one context row: pass 1 finds the predicted best arm by warp
// shuffle, pass 2 re-reads the row (L1-resident, the row was just streamed) for
//     zeta_j = (mu_b - mu_j)^2 / (var_b + var_j)          [same IEEE ops as the reference]
// and a last-arriving CTA folds the per-CTA minima into the global decision.  HBM-bound: one read
// of mean and var (2*K*n_c*w bytes) plus n_c*(w+12) bytes of per-context outputs.
#include <math_constants.h>
#include <cfloat>
#include <climits>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanMaxBlocks = 148 * 8;

struct __align__(16) ScanPartial {   // 32 bytes
  double z;            // min zeta
  long long cnt;       // entries equal to z
  int ctx;             // lowest context attaining z
  int zarm;            // zeta-minimising arm id there
  int best;            // predicted best arm id there
  int pad;
};

__device__ __forceinline__ void combine(ScanPartial& a, const ScanPartial& b) {
  if (b.z < a.z) {
    a = b;
  } else if (b.z == a.z) {
    a.cnt += b.cnt;
    if (b.ctx < a.ctx) {
      a.ctx = b.ctx;
      a.zarm = b.zarm;
      a.best = b.best;
    }
  }
}

__device__ __forceinline__ ScanPartial shfl_partial(const ScanPartial& p, int o) {
  ScanPartial r;
  r.z = __shfl_xor_sync(kFull, p.z, o);
  r.cnt = __shfl_xor_sync(kFull, p.cnt, o);
  r.ctx = __shfl_xor_sync(kFull, p.ctx, o);
  r.zarm = __shfl_xor_sync(kFull, p.zarm, o);
  r.best = __shfl_xor_sync(kFull, p.best, o);
  r.pad = 0;
  return r;
}

template <typename T> __device__ __forceinline__ T pos_inf();
template <> __device__ __forceinline__ double pos_inf<double>() { return CUDART_INF; }
template <> __device__ __forceinline__ float pos_inf<float>() { return CUDART_INF_F; }

// zeta exactly as the reference computes it: squared_diffs / added_vars (strategies.py:140-143).
template <typename T> __device__ __forceinline__ T zeta_of(T mu_b, T var_b, T mu, T vr) {
  const T gap = mu_b - mu;
  const T sq = gap * gap;
  const T den = var_b + vr;
  return sq / den;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
ocba_scan_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K,
                 int64_t ld, int arm_offset, const int32_t* __restrict__ ext_arm,
                 const T* __restrict__ ext_mean, const T* __restrict__ ext_var,
                 int32_t* __restrict__ best_arm, T* __restrict__ min_zeta,
                 int32_t* __restrict__ min_arm, int32_t* __restrict__ tie_cnt,
                 ScanPartial* partials, unsigned int* ticket, crs_decision* decision) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * kScanWarps;
  ScanPartial acc;
  acc.z = CUDART_INF;
  acc.cnt = 0;
  acc.ctx = INT_MAX;
  acc.zarm = -1;
  acc.best = -1;
  acc.pad = 0;

  for (int64_t c = (int64_t)blockIdx.x * kScanWarps + warp; c < n_c; c += warps_total) {
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    T bm, vb;
    int bi;  // local column of the best arm (may fall outside [0, K) when it lives on a peer rank)
    if (ext_arm == nullptr) {
      bm = -pos_inf<T>();
      bi = INT_MAX;
      for (int a = lane; a < K; a += 32) {
        const T m = mrow[a];
        if (m > bm) {
          bm = m;
          bi = a;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T om = __shfl_xor_sync(kFull, bm, o);
        const int oi = __shfl_xor_sync(kFull, bi, o);
        if (om > bm || (om == bm && oi < bi)) {
          bm = om;
          bi = oi;
        }
      }
      if (bi == INT_MAX) bi = 0;  // all-NaN row
      vb = vrow[bi];
    } else {
      bi = ext_arm[c] - arm_offset;
      bm = ext_mean[c];
      vb = ext_var[c];
    }

    T zmin = pos_inf<T>(), zmu = -pos_inf<T>();
    int zi = INT_MAX, cnt = 0;
    for (int a = lane; a < K; a += 32) {
      if (a == bi) continue;
      const T m = mrow[a];
      const T z = zeta_of<T>(bm, vb, m, vrow[a]);
      if (z < zmin) {
        zmin = z;
        zmu = m;
        zi = a;
        cnt = 1;
      } else if (z == zmin) {
        ++cnt;
        if (m > zmu) {
          zmu = m;
          zi = a;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T oz = __shfl_xor_sync(kFull, zmin, o);
      const T omu = __shfl_xor_sync(kFull, zmu, o);
      const int oi = __shfl_xor_sync(kFull, zi, o);
      const int oc = __shfl_xor_sync(kFull, cnt, o);
      if (oz < zmin) {
        zmin = oz;
        zmu = omu;
        zi = oi;
        cnt = oc;
      } else if (oz == zmin) {
        cnt += oc;
        if (omu > zmu || (omu == zmu && oi < zi)) {
          zmu = omu;
          zi = oi;
        }
      }
    }
    const int gbest = bi + arm_offset;
    const int gz = zi == INT_MAX ? -1 : zi + arm_offset;
    if (lane == 0) {
      if (best_arm) best_arm[c] = gbest;
      if (min_zeta) min_zeta[c] = zmin;
      if (min_arm) min_arm[c] = gz;
      if (tie_cnt) tie_cnt[c] = cnt;
    }
    ScanPartial p;
    p.z = (double)zmin;
    p.cnt = cnt;
    p.ctx = (int)c;
    p.zarm = gz;
    p.best = gbest;
    p.pad = 0;
    if (cnt > 0) combine(acc, p);
  }

  if (decision == nullptr) return;

  __shared__ ScanPartial s_part[kScanWarps];
  __shared__ bool s_last;
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    ScanPartial p = lane < kScanWarps ? s_part[lane] : acc;
    if (lane >= kScanWarps) {
      p.z = CUDART_INF;
      p.cnt = 0;
      p.ctx = INT_MAX;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const ScanPartial q = shfl_partial(p, o);
      combine(p, q);
    }
    if (lane == 0) {
      partials[blockIdx.x] = p;
      __threadfence();
      const unsigned int t = atomicAdd(ticket, 1u);
      s_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: fold all per-CTA partials
  ScanPartial p;
  p.z = CUDART_INF;
  p.cnt = 0;
  p.ctx = INT_MAX;
  p.zarm = -1;
  p.best = -1;
  p.pad = 0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += kScanThreads) {
    // written by other CTAs: read through L2 (ld.global.cg), never a stale L1 line
    const int4 w0 = __ldcg(reinterpret_cast<const int4*>(partials + b));
    const int4 w1 = __ldcg(reinterpret_cast<const int4*>(partials + b) + 1);
    ScanPartial q;
    q.z = __hiloint2double(w0.y, w0.x);
    q.cnt = ((long long)(unsigned int)w0.w << 32) | (unsigned int)w0.z;
    q.ctx = w1.x;
    q.zarm = w1.y;
    q.best = w1.z;
    q.pad = 0;
    combine(p, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const ScanPartial q = shfl_partial(p, o);
    combine(p, q);
  }
  __syncthreads();
  if (lane == 0) s_part[warp] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kScanWarps; ++w) combine(p, s_part[w]);
    decision->min_zeta = p.z;
    decision->tie_count = p.cnt;
    decision->context = p.ctx == INT_MAX ? -1 : p.ctx;
    decision->zeta_arm = p.zarm;
    decision->best_arm = p.best;
    decision->next_arm = -1;
    *ticket = 0;  // leave the workspace ready for the next launch
  }
}

// ---- tie randomisation (rare path; strategies.py:147-154) -----------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
resolve_tie_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K,
                   int64_t ld, const int32_t* __restrict__ best_arm, const T* __restrict__ min_zeta,
                   const int32_t* __restrict__ tie_cnt, long long r, crs_decision* decision) {
  __shared__ long long s_sum[1024];
  const int tid = threadIdx.x;
  const double gmin = decision->min_zeta;
  const int64_t chunk = (n_c + 1023) / 1024;
  const int64_t c0 = (int64_t)tid * chunk, c1 = min(c0 + chunk, n_c);
  long long s = 0;
  for (int64_t c = c0; c < c1; ++c)
    if ((double)min_zeta[c] == gmin) s += tie_cnt[c];
  s_sum[tid] = s;
  __syncthreads();
  __shared__ int s_owner;
  __shared__ long long s_rem;
  if (tid == 0) {
    long long run = 0;
    s_owner = -1;
    for (int t = 0; t < 1024; ++t) {
      if (r < run + s_sum[t]) {
        s_owner = t;
        s_rem = r - run;
        break;
      }
      run += s_sum[t];
    }
  }
  __syncthreads();
  if (tid != s_owner) return;
  long long rem = s_rem;
  for (int64_t c = c0; c < c1; ++c) {
    if ((double)min_zeta[c] != gmin) continue;
    if (rem >= tie_cnt[c]) {
      rem -= tie_cnt[c];
      continue;
    }
    // the rem-th tied arm of row c in descending-mean (then ascending id) order
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    const int b = best_arm[c];
    const T bm = mrow[b], vb = vrow[b], zt = min_zeta[c];
    int chosen = -1;
    for (int a = 0; a < K && chosen < 0; ++a) {
      if (a == b || !(zeta_of<T>(bm, vb, mrow[a], vrow[a]) == zt)) continue;
      long long rank = 0;
      for (int e = 0; e < K; ++e) {
        if (e == b || e == a || !(zeta_of<T>(bm, vb, mrow[e], vrow[e]) == zt)) continue;
        if (mrow[e] > mrow[a] || (mrow[e] == mrow[a] && e < a)) ++rank;
      }
      if (rank == rem) chosen = a;
    }
    decision->context = (int)c;
    decision->zeta_arm = chosen;
    decision->best_arm = b;
    return;
  }
}

// ---- OCBA step 2(b) (strategies.py:158-202; continuous_context.py:110-132) --------------------
template <typename T>
__global__ void __launch_bounds__(256)
ocba_select_kernel(crs_gp_state st, const T* __restrict__ ctx, const T* __restrict__ var,
                   int64_t ld, int arm_offset, int p_mode, T kernel_scale,
                   crs_decision* decision) {
  __shared__ T s_x[CRS_MAX_DIM];
  __shared__ T s_red[8];
  __shared__ T s_best;
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int c = decision->context;
  const int best = decision->best_arm - arm_offset;
  const int d = st.dim, ncap = st.n_cap, K = st.num_arms;
  if (c < 0) return;
  if (tid < d) s_x[tid] = ctx[(int64_t)c * d + tid];
  if (tid == 0) {
    s_best = T(0);
    s_bad = 0;
  }
  __syncthreads();
  T rest = T(0);
  for (int i = tid; i < K; i += blockDim.x) {
    const T vr = var[(int64_t)c * ld + i];
    const int n = min(st.n_obs[i], ncap);
    const T* xn = reinterpret_cast<const T*>(st.xn) + (size_t)i * ncap * d;
    T p;
    if (p_mode == CRS_P_COUNT) {
      int cnt = 0;
      for (int row = 0; row < n; ++row) {
        bool eq = true;
        for (int dd = 0; dd < d; ++dd) eq = eq && (xn[row * d + dd] == s_x[dd]);
        cnt += eq ? 1 : 0;
      }
      p = T(cnt);
    } else if (p_mode == CRS_P_KDE) {
      T s = T(0);
      for (int row = 0; row < n; ++row) {
        T r2 = T(0);
        for (int dd = 0; dd < d; ++dd) {
          const T df = xn[row * d + dd] - s_x[dd];
          r2 += df * df;
        }
        s += exp(-kernel_scale * sqrt(r2));
      }
      p = s;
    } else {
      p = reinterpret_cast<const T*>(st.arm_head)[(size_t)i * kHeadElems + kHeadScal] / vr;  // prior var / posterior var
    }
    const T y = p / vr;
    if (i == best) s_best = y; else rest += y;
    if (st.status[i] != 0) atomicMax(&s_bad, i + 1);
  }
  rest = warp_sum(rest);
  if ((tid & 31) == 0) s_red[tid >> 5] = rest;
  __syncthreads();
  if (tid == 0) {
    T tot = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
    const T yb = s_best;
    decision->psi_best = (double)yb;
    decision->psi_rest = (double)tot;
    decision->next_arm = (yb < tot) ? decision->best_arm : decision->zeta_arm;
    decision->reserved[0] = s_bad;
  }
}

}  // namespace

int64_t scan_workspace_bytes() { return (int64_t)kScanMaxBlocks * sizeof(ScanPartial) + 64; }

int launch_ocba_scan(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                     int arm_offset, const int32_t* ext_best_arm, const void* ext_best_mean,
                     const void* ext_best_var, int32_t* best_arm, void* min_zeta, int32_t* min_arm,
                     int32_t* tie_cnt, crs_decision* decision, void* workspace, cudaStream_t s) {
  if (!mean || !var || n_c < 0 || n_c > INT_MAX || K < 1 || ld < K) return CRS_ERR_INVALID_ARG;
  if (decision && !workspace) return CRS_ERR_INVALID_ARG;
  if (ext_best_arm && (!ext_best_mean || !ext_best_var)) return CRS_ERR_INVALID_ARG;
  if (n_c == 0) return CRS_OK;
  int blocks = (int)((n_c + kScanWarps - 1) / kScanWarps);
  if (blocks > kScanMaxBlocks) blocks = kScanMaxBlocks;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  ScanPartial* partials = workspace ? reinterpret_cast<ScanPartial*>(reinterpret_cast<char*>(workspace) + 64) : nullptr;
  if (dtype == CRS_F64) {
    ocba_scan_kernel<double><<<blocks, kScanThreads, 0, s>>>(
        (const double*)mean, (const double*)var, n_c, K, ld, arm_offset, ext_best_arm,
        (const double*)ext_best_mean, (const double*)ext_best_var, best_arm, (double*)min_zeta,
        min_arm, tie_cnt, partials, ticket, decision);
  } else if (dtype == CRS_F32) {
    ocba_scan_kernel<float><<<blocks, kScanThreads, 0, s>>>(
        (const float*)mean, (const float*)var, n_c, K, ld, arm_offset, ext_best_arm,
        (const float*)ext_best_mean, (const float*)ext_best_var, best_arm, (float*)min_zeta,
        min_arm, tie_cnt, partials, ticket, decision);
  } else {
    return CRS_ERR_INVALID_ARG;
  }
  CRS_CUDA_TRY(cudaGetLastError(), "ocba_scan launch");
  return CRS_OK;
}

int launch_ocba_resolve_tie(const void* mean, const void* var, int64_t n_c, int K, int64_t ld,
                            int dtype, const int32_t* best_arm, const void* min_zeta,
                            const int32_t* tie_cnt, int64_t r, crs_decision* decision,
                            cudaStream_t s) {
  if (!mean || !var || !best_arm || !min_zeta || !tie_cnt || !decision || r < 0) return CRS_ERR_INVALID_ARG;
  if (dtype == CRS_F64)
    resolve_tie_kernel<double><<<1, 1024, 0, s>>>((const double*)mean, (const double*)var, n_c, K, ld, best_arm, (const double*)min_zeta, tie_cnt, (long long)r, decision);
  else if (dtype == CRS_F32)
    resolve_tie_kernel<float><<<1, 1024, 0, s>>>((const float*)mean, (const float*)var, n_c, K, ld, best_arm, (const float*)min_zeta, tie_cnt, (long long)r, decision);
  else
    return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "resolve_tie launch");
  return CRS_OK;
}

int launch_ocba_select(const crs_gp_state* st, const void* ctx, const void* mean, const void* var,
                       int64_t ld, int arm_offset, int p_mode, double kernel_scale,
                       crs_decision* decision, cudaStream_t s) {
  (void)mean;
  if (!st || !ctx || !var || !decision || p_mode < 0 || p_mode > 2) return CRS_ERR_INVALID_ARG;
  if (st->dtype == CRS_F64)
    ocba_select_kernel<double><<<1, 256, 0, s>>>(*st, (const double*)ctx, (const double*)var, ld, arm_offset, p_mode, kernel_scale, decision);
  else if (st->dtype == CRS_F32)
    ocba_select_kernel<float><<<1, 256, 0, s>>>(*st, (const float*)ctx, (const float*)var, ld, arm_offset, p_mode, (float)kernel_scale, decision);
  else
    return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "ocba_select launch");
  return CRS_OK;
}

}  // namespace crs
