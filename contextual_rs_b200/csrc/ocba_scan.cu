// crs_ocba_scan / crs_ocba_resolve_tie / crs_ocba_select (K2): the GP-C-OCBA allocation on a
// materialised posterior grid — replaces contextual_rs/contextual_rs_strategies.py:137-202 and
// contextual_rs/continuous_context.py:21-36, 57-74, 96-132.
//
// The reference sorts every row of the n_c x K table, gathers variances, builds five n_c x (K-1)
// temporaries and arg-mins the flattened zeta.  Only the arg-max arm and the zeta minimum of each
// row are needed, so one warp owns one context row: pass 1 finds the predicted best arm by warp
// shuffle, pass 2 evaluates
//     zeta_j = (mu_b - mu_j)^2 / (var_b + var_j)          [same IEEE ops as the reference]
// on the row held in registers (K <= 256) or re-read from L1 (larger K), and the last-arriving CTA
// folds the per-CTA minima into the global decision and — fused — runs OCBA step 2(b) for the
// selected context, so that a whole decision needs no host round-trip between its kernels.
// HBM-bound: one read of mean and var (2*K*n_c*w bytes) plus n_c*(w+12) bytes of outputs.
#include <math_constants.h>
#include <cfloat>
#include <climits>
#include <cstring>
#include "crs_common.cuh"
#include "peer_mailbox.cuh"

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

// step 2(b) arguments for the fused tail of the scan kernel (do_select = 0: scan only)
struct SelectArgs {
  crs_gp_state st;
  const void* ctx;
  double kernel_scale;
  int p_mode;
  int do_select;
  crs_decision* host_mirror;  // mapped pinned host copy of the decision (may be null)
  int seq;                    // written to reserved[1] of the mirror LAST: completion tag
  crs_peer_exchange px;       // world > 1: combine the ranks' local decisions through peer mailboxes
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

__device__ __forceinline__ ScanPartial empty_partial() {
  ScanPartial p;
  p.z = CUDART_INF;
  p.cnt = 0;
  p.ctx = INT_MAX;
  p.zarm = -1;
  p.best = -1;
  p.pad = 0;
  return p;
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

// fp64 division is ~30 instructions, so the row minimum is located with an approximate quotient
// (rcp.approx + one Newton step: relative error < 1e-11) and only the entries within 2e-9 of the
// approximate minimum — a superset of the true minimisers — pay for the exact IEEE quotient the
// reference computes.  The minimum, its position and its multiplicity are therefore bit-identical
// to the unscreened scan.
__device__ __forceinline__ double approx_quot(double sq, double den) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(fma(-den, r, 1.0), r, r);
  return sq * r;
}
constexpr double kScreenSlack = 1.0 + 2e-9;
__device__ __forceinline__ float approx_quot(float sq, float den) { return sq / den; }  // fp32: exact already

// cheap stand-in for zeta used to find the candidates of the row minimum
template <typename T> __device__ __forceinline__ T zeta_approx(T mu_b, T var_b, T m, T vr) {
  const T gap = mu_b - m;
  return approx_quot(gap * gap, var_b + vr);
}
template <typename T> __device__ __forceinline__ T screen_threshold(T zt_min) {
  return sizeof(T) == 8 ? (T)(zt_min * (T)kScreenSlack) : zt_min;
}
template <typename T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// running (min zeta, its mean, its column, multiplicity); ties: larger mean, then lower column
template <typename T> struct ZMin {
  T z, mu;
  int col, cnt;
  __device__ __forceinline__ void init() { z = pos_inf<T>(); mu = -pos_inf<T>(); col = INT_MAX; cnt = 0; }
  __device__ __forceinline__ void add(T zz, T m, int a) {
    if (zz < z) { z = zz; mu = m; col = a; cnt = 1; }
    else if (zz == z) { ++cnt; if (m > mu) { mu = m; col = a; } }
  }
  // zeta of (mu_b, var_b) against (m, vr), exactly as the reference: squared_diffs / added_vars
  __device__ __forceinline__ void add_exact(T mu_b, T var_b, T m, T vr, int a) {
    const T gap = mu_b - m;
    add((gap * gap) / (var_b + vr), m, a);
  }
  __device__ __forceinline__ void warp_reduce() {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T oz = __shfl_xor_sync(kFull, z, o);
      const T omu = __shfl_xor_sync(kFull, mu, o);
      const int oi = __shfl_xor_sync(kFull, col, o);
      const int oc = __shfl_xor_sync(kFull, cnt, o);
      if (oz < z) { z = oz; mu = omu; col = oi; cnt = oc; }
      else if (oz == z) { cnt += oc; if (omu > mu || (omu == mu && oi < col)) { mu = omu; col = oi; } }
    }
  }
};

template <typename T> __device__ __forceinline__ void warp_argmax(T& bm, int& bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T om = __shfl_xor_sync(kFull, bm, o);
    const int oi = __shfl_xor_sync(kFull, bi, o);
    if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
  }
}

// ---- OCBA step 2(b) for one CTA (strategies.py:158-202; continuous_context.py:110-132) ---------
// p_i = exact-match count | KDE weight | prior/posterior variance ratio; y_i = p_i / var_i; psi of
// the best arm vs the sum over the others.  The normalised training inputs of a chunk of arms are
// staged in shared memory by TMA bulk copies issued back to back (ONE memory round trip per chunk
// instead of one per arm — the step is pure latency), then one warp reduces one arm, lanes over
// its training rows.
// The stage buffer is the kernel's dynamic shared memory (stage_bytes of it).
constexpr int kSelStageFused = 32 * 1024;    // fused tail of the scan kernel (small problems only)
constexpr int kSelStageAlone = 160 * 1024;   // stand-alone select kernel

template <typename T>
__device__ void select_block(const crs_gp_state& st, const T* __restrict__ ctx,
                             const T* __restrict__ var, int64_t ld, int arm_offset, int p_mode,
                             T kernel_scale, crs_decision* decision, unsigned char* s_stage_raw,
                             int stage_bytes) {
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ T s_x[CRS_MAX_DIM];
  __shared__ T s_v[256];
  __shared__ int s_n[256];
  __shared__ double s_red[32][2];
  __shared__ T s_best;
  __shared__ int s_bad;
  T* s_stage = reinterpret_cast<T*>(s_stage_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int c = decision->context;
  if (c < 0) return;  // uniform
  const int best = decision->best_arm - arm_offset;
  const int d = st.dim, ncap = st.n_cap, K = st.num_arms;
  const int blk = ncap * d;  // elements of one arm's xn block
  const bool rows = p_mode != CRS_P_INFER;
  int chunk = min(min((int)(stage_bytes / (blk * sizeof(T))), 256), K);
  if (chunk < 1) chunk = 1;  // (blk <= 160*16 elements always fits at least once)
  if (tid < d) s_x[tid] = ctx[(int64_t)c * d + tid];
  if (tid == 0) {
    s_best = T(0);
    s_bad = 0;
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const T* g_xn = reinterpret_cast<const T*>(st.xn);
  // y_s_2 (strategies.py:196) is a sum of K - 1 terms whose order the reference leaves to torch's
  // vectorised CPU reduction (machine-dependent in the last ulp).  Here it is accumulated in
  // double-double (error ~2^-104) and rounded ONCE to T: the result does not depend on the order,
  // on the warp layout or — arm-sharded runs — on how the arms are split over ranks.
  double rest_hi = 0.0, rest_lo = 0.0;
  uint32_t phase = 0;
  for (int c0 = 0; c0 < K; c0 += chunk) {
    const int na = min(chunk, K - c0);
    if (rows && tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&s_bar, (uint32_t)(na * blk * sizeof(T)));
      for (int a = 0; a < na; ++a)
        tma_bulk_g2s(s_stage + (size_t)a * blk, g_xn + (size_t)(c0 + a) * blk, (uint32_t)(blk * sizeof(T)), &s_bar);
    }
    for (int a = tid; a < na; a += blockDim.x) {
      s_v[a] = var[(int64_t)c * ld + c0 + a];
      s_n[a] = min(st.n_obs[c0 + a], ncap);
      if (st.status[c0 + a] != 0) atomicMax(&s_bad, c0 + a + 1);
    }
    if (rows) {
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
    }
    __syncthreads();
    for (int a = warp; a < na; a += nw) {
      const T vr = s_v[a];
      T p;
      if (!rows) {
        p = reinterpret_cast<const T*>(st.arm_head)[(size_t)(c0 + a) * kHeadElems + kHeadScal] / vr;  // prior / posterior
      } else {
        const int n = s_n[a];
        const T* xn = s_stage + (size_t)a * blk;
        T acc = T(0);
        for (int row = lane; row < n; row += 32) {
          if (p_mode == CRS_P_COUNT) {
            bool eq = true;
            for (int dd = 0; dd < d; ++dd) eq = eq && (xn[row * d + dd] == s_x[dd]);
            acc += eq ? T(1) : T(0);
          } else {
            T r2 = T(0);
            for (int dd = 0; dd < d; ++dd) {
              const T df = xn[row * d + dd] - s_x[dd];
              r2 += df * df;
            }
            acc += exp(-kernel_scale * sqrt(r2));
          }
        }
        p = warp_sum(acc);
      }
      if (lane == 0) {
        const T y = p / vr;
        if (c0 + a == best) {
          s_best = y;
        } else {  // TwoSum
          const double yd = (double)y;
          const double sh = rest_hi + yd;
          const double bb = sh - rest_hi;
          rest_lo += (rest_hi - (sh - bb)) + (yd - bb);
          rest_hi = sh;
        }
      }
    }
    __syncthreads();  // the stage buffer is free for the next chunk
  }
  if (lane == 0) { s_red[warp][0] = rest_hi; s_red[warp][1] = rest_lo; }
  __syncthreads();
  if (tid == 0) {
    double th = 0.0, tl = 0.0;
    for (int w = 0; w < nw; ++w) {
      const double yd = s_red[w][0];
      const double sh = th + yd;
      const double bb = sh - th;
      tl += (th - (sh - bb)) + (yd - bb) + s_red[w][1];
      th = sh;
    }
    const T tot = (T)(th + tl);
    const T yb = s_best;
    decision->psi_best = (double)yb;
    decision->psi_rest = (double)tot;
    decision->next_arm = (yb < tot) ? decision->best_arm : decision->zeta_arm;
    decision->reserved[0] = s_bad;
  }
}

// ---- multi-GPU combine of a context-sharded decision, device side ------------------------------
// The rank's finishing CTA posts its 64-byte record (context made global) to every peer's mailbox
// (peer_mailbox.cuh), waits for the world records of this exchange in its own and folds them with the
// single-GPU rule: smallest zeta, then lowest global context; tie counts of the ranks that attain the
// minimum are added.  A peer that never shows up ends the wait after ~2 s with status -2.
__device__ void peer_exchange_fold(crs_decision* dec, const crs_peer_exchange& px) {
  __shared__ int s_timeout;
  const int t = threadIdx.x;
  if (t == 0) s_timeout = 0;
  __syncthreads();  // the local decision is complete and visible to the CTA
  if (t < px.world) {
    alignas(16) crs_decision rec = *dec;
    if (rec.context >= 0) rec.context += (int32_t)px.ctx_begin;
    else rec.min_zeta = CUDART_INF;
    if (!mailbox::post_and_wait(px, t, reinterpret_cast<const int4*>(&rec))) atomicExch(&s_timeout, 1);
  }
  __syncthreads();
  if (t == 0) {
    crs_decision out;
    out.min_zeta = CUDART_INF;
    out.tie_count = 0;
    out.context = -1;
    out.zeta_arm = out.best_arm = out.next_arm = -1;
    out.psi_best = out.psi_rest = 0.0;
    int bad = 0;
    for (int r = 0; r < px.world; ++r) {
      alignas(16) crs_decision q;
      mailbox::read_slot(px, r, reinterpret_cast<int4*>(&q));
      if (q.reserved[0] != 0 && bad == 0) bad = q.reserved[0];
      if (q.context < 0) continue;  // an empty shard
      if (q.min_zeta < out.min_zeta || (q.min_zeta == out.min_zeta && out.context < 0)) {
        const long long ties = q.tie_count;
        out = q;
        out.tie_count = ties;
      } else if (q.min_zeta == out.min_zeta) {
        const long long ties = out.tie_count + q.tie_count;
        if (q.context < out.context) out = q;
        out.tie_count = ties;
      }
    }
    out.reserved[0] = s_timeout ? -2 : bad;
    out.reserved[1] = out.reserved[2] = out.reserved[3] = 0;
    *dec = out;
  }
  __syncthreads();
}

__device__ __forceinline__ void publish(const crs_decision* dec, crs_decision* mirror, int seq) {
  // thread 0 only: copy the finished decision to mapped host memory; the tag goes last
  mirror->min_zeta = dec->min_zeta;
  mirror->psi_best = dec->psi_best;
  mirror->psi_rest = dec->psi_rest;
  mirror->tie_count = dec->tie_count;
  mirror->context = dec->context;
  mirror->zeta_arm = dec->zeta_arm;
  mirror->best_arm = dec->best_arm;
  mirror->next_arm = dec->next_arm;
  mirror->reserved[0] = dec->reserved[0];
  __threadfence_system();
  *reinterpret_cast<volatile int*>(&mirror->reserved[1]) = seq;
  __threadfence_system();
}

// Fold the per-CTA minima into the global decision (ticket atomic: the last-arriving CTA reduces
// all partials) and optionally run the fused step 2(b).  `sel_smem`: stage buffer for the fused
// select (nullptr = the kernel's dynamic shared memory).
template <typename T>
__device__ void scan_tail(ScanPartial acc, ScanPartial* partials, unsigned int* ticket,
                          crs_decision* decision, const SelectArgs& sel, const T* __restrict__ var,
                          int64_t ld, int arm_offset, unsigned char* sel_smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  if (decision == nullptr) return;

  __shared__ ScanPartial s_part[kScanWarps];  // blockDim <= kScanThreads
  __shared__ bool s_last;
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    ScanPartial p = lane < nwarps ? s_part[lane] : empty_partial();
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
  ScanPartial p = empty_partial();
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
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
    for (int w = 1; w < nwarps; ++w) combine(p, s_part[w]);
    decision->min_zeta = p.z;
    decision->tie_count = p.cnt;
    decision->context = p.ctx == INT_MAX ? -1 : p.ctx;
    decision->zeta_arm = p.zarm;
    decision->best_arm = p.best;
    decision->next_arm = -1;
    decision->reserved[0] = 0;
    *ticket = 0;  // leave the workspace ready for the next launch
  }
  if (sel.do_select) {
    __syncthreads();  // decision fields are visible to the whole CTA
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    select_block<T>(sel.st, reinterpret_cast<const T*>(sel.ctx), var, ld, arm_offset, sel.p_mode,
                    (T)sel.kernel_scale, decision, sel_smem ? sel_smem : dyn_smem, kSelStageFused);
  }
  if (sel.px.world > 1) peer_exchange_fold(decision, sel.px);  // set only when this launch finishes the decision
  if (sel.host_mirror && threadIdx.x == 0) publish(decision, sel.host_mirror, sel.seq);
}

// One row of the scan.  WPR warps share a row (WPR = 1: a warp per row; WPR = 4: the 4 warps of a
// 128-thread CTA split a row of up to 1024 columns, 8 per lane, so that the whole row is in flight
// at once with ~100 registers per thread instead of 250).  R = columns per lane held in registers
// (K <= 32 R WPR); R == 0: generic path, pass 2 re-reads the row (L1 / L2).
// Column owned by (warp w of the row, lane, slot r): a = (r WPR + w) 32 + lane.
template <typename T, int R, int WPR>
__global__ void __launch_bounds__(WPR == 4 ? 128 : kScanThreads, (R == 0 ? 4 : (WPR == 4 ? 4 : 1)))
ocba_scan_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K,
                 int64_t ld, int arm_offset, const int32_t* __restrict__ ext_arm,
                 const T* __restrict__ ext_mean, const T* __restrict__ ext_var,
                 int32_t* __restrict__ best_arm, T* __restrict__ min_zeta,
                 int32_t* __restrict__ min_arm, int32_t* __restrict__ tie_cnt,
                 ScanPartial* partials, unsigned int* ticket, crs_decision* decision,
                 SelectArgs sel) {
  griddep_wait();  // the posterior grid of the same decision may still be running (PDL)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  ScanPartial acc = empty_partial();
  // cross-warp exchange for WPR = 4 (double-buffered by row parity: one barrier per stage)
  __shared__ T s_xm[2][4], s_xv[2][4], s_xz[2][4], s_zz[2][4], s_zmu[2][4];
  __shared__ int s_xi[2][4], s_zc[2][4], s_zn[2][4];

  const int64_t row0 = WPR == 4 ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * nwarps + warp;
  const int64_t row_step = WPR == 4 ? (int64_t)gridDim.x : (int64_t)gridDim.x * nwarps;
  const int w = WPR == 4 ? warp : 0;
  int par = 0;
  for (int64_t c = row0; c < n_c; c += row_step, par ^= 1) {
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    T bm, vb;
    int bi;  // local column of the best arm (may fall outside [0, K) when it lives on a peer rank)
    ZMin<T> zm;
    zm.init();
    if constexpr (R > 0) {
      T m[R], v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int a = (r * WPR + w) * 32 + lane;
        if ((r * WPR + w + 1) * 32 <= K) {  // warp-uniform: a full 32-column segment
          m[r] = mrow[a];
          v[r] = vrow[a];
        } else {
          m[r] = a < K ? mrow[a] : -pos_inf<T>();  // sentinels: never the best, zeta = +inf
          v[r] = a < K ? vrow[a] : T(1);
        }
      }
      if (ext_arm == nullptr) {
        bm = m[0];
        int br = 0;
#pragma unroll
        for (int r = 1; r < R; ++r)
          if (m[r] > bm) { bm = m[r]; br = r; }
        T vcand = v[0];
#pragma unroll
        for (int r = 1; r < R; ++r) vcand = br == r ? v[r] : vcand;
        bi = (br * WPR + w) * 32 + lane;
        if (!(bm > -pos_inf<T>())) bi = INT_MAX;  // sentinel / NaN: not a candidate
        const int mine = bi;
        warp_argmax(bm, bi);
        const unsigned own = __ballot_sync(kFull, mine == bi && bi != INT_MAX);
        vb = __shfl_sync(kFull, vcand, own ? __ffs(own) - 1 : 0);  // the winner's variance sits in its owner lane
        if constexpr (WPR == 4) {
          if (lane == 0) { s_xm[par][w] = bm; s_xv[par][w] = vb; s_xi[par][w] = bi; }
          __syncthreads();
          bm = s_xm[par][0]; vb = s_xv[par][0]; bi = s_xi[par][0];
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            const T om = s_xm[par][q];
            const int oi = s_xi[par][q];
            if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; vb = s_xv[par][q]; }
          }
        }
        if (bi == INT_MAX) { bi = 0; vb = vrow[0]; bm = mrow[0]; }  // all-NaN row
      } else {
        bi = ext_arm[c] - arm_offset;
        bm = ext_mean[c];
        vb = ext_var[c];
      }
      // slot of the best arm in this lane (or -1): excluded from the zeta minimum
      const int rb = ((bi & 31) == lane && ((bi >> 5) % WPR) == w && bi >= 0) ? (bi >> 5) / WPR : -1;
      // pass 2a: approximate zeta (no IEEE division), row minimum of the approximation
      T zt_min = pos_inf<T>();
#ifdef CRS_SCAN_EXPERIMENT
      T vsum = T(0);
#pragma unroll
      for (int r = 0; r < R; ++r) vsum += v[r];
      zt_min = vsum;
#else
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T zt = zeta_approx<T>(bm, vb, m[r], v[r]);
        if (r != rb) zt_min = fmin(zt_min, zt);
      }
#endif
      zt_min = warp_min<T>(zt_min);
      if constexpr (WPR == 4) {
        if (lane == 0) s_xz[par][w] = zt_min;
        __syncthreads();
        zt_min = fmin(fmin(s_xz[par][0], s_xz[par][1]), fmin(s_xz[par][2], s_xz[par][3]));
      }
      const T thr = screen_threshold<T>(zt_min);
      // pass 2b: mark the (typically one per row) candidates within the screening slack; only
      // they divide exactly.  Branch-free marking keeps the unrolled loop free of phi-moves.
      unsigned cand = 0;
#ifndef CRS_SCAN_EXPERIMENT
#pragma unroll
      for (int r = 0; r < R; ++r)
        cand |= (r != rb && zeta_approx<T>(bm, vb, m[r], v[r]) <= thr) ? (1u << r) : 0u;
#endif
      while (__any_sync(kFull, cand != 0)) {
        if (cand) {
          const int r = __ffs(cand) - 1;
          cand &= cand - 1;
          T mm = m[0], vv = v[0];
#pragma unroll
          for (int q = 1; q < R; ++q) {
            mm = q == r ? m[q] : mm;
            vv = q == r ? v[q] : vv;
          }
          zm.add_exact(bm, vb, mm, vv, (r * WPR + w) * 32 + lane);
        }
      }
    } else {
      if (ext_arm == nullptr) {
        bm = -pos_inf<T>();
        bi = INT_MAX;
        for (int a = lane; a < K; a += 32) {
          const T m = mrow[a];
          if (m > bm) { bm = m; bi = a; }
        }
        warp_argmax(bm, bi);
        if (bi == INT_MAX) bi = 0;
        vb = vrow[bi];
      } else {
        bi = ext_arm[c] - arm_offset;
        bm = ext_mean[c];
        vb = ext_var[c];
      }
      T zt_min = pos_inf<T>();
      for (int a = lane; a < K; a += 32)
        if (a != bi) zt_min = fmin(zt_min, zeta_approx<T>(bm, vb, mrow[a], vrow[a]));
      const T thr = screen_threshold<T>(warp_min<T>(zt_min));
      for (int a = lane; a < K; a += 32) {
        if (a == bi) continue;
        const T m = mrow[a], vr = vrow[a];
        if (zeta_approx<T>(bm, vb, m, vr) <= thr) zm.add_exact(bm, vb, m, vr, a);
      }
    }
    zm.warp_reduce();
    if constexpr (WPR == 4) {
      if (lane == 0) { s_zz[par][w] = zm.z; s_zmu[par][w] = zm.mu; s_zc[par][w] = zm.col; s_zn[par][w] = zm.cnt; }
      __syncthreads();
      if (warp != 0) continue;  // warp 0 finishes the row
      zm.z = s_zz[par][0]; zm.mu = s_zmu[par][0]; zm.col = s_zc[par][0]; zm.cnt = s_zn[par][0];
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const T oz = s_zz[par][q], omu = s_zmu[par][q];
        const int oi = s_zc[par][q], oc = s_zn[par][q];
        if (oz < zm.z) { zm.z = oz; zm.mu = omu; zm.col = oi; zm.cnt = oc; }
        else if (oz == zm.z) { zm.cnt += oc; if (omu > zm.mu || (omu == zm.mu && oi < zm.col)) { zm.mu = omu; zm.col = oi; } }
      }
    }
    const int gbest = bi + arm_offset;
    const int gz = zm.col == INT_MAX ? -1 : zm.col + arm_offset;
    if (lane == 0) {
      if (best_arm) best_arm[c] = gbest;
      if (min_zeta) min_zeta[c] = zm.z;
      if (min_arm) min_arm[c] = gz;
      if (tie_cnt) tie_cnt[c] = zm.cnt;
    }
    ScanPartial p;
    p.z = (double)zm.z;
    p.cnt = zm.cnt;
    p.ctx = (int)c;
    p.zarm = gz;
    p.best = gbest;
    p.pad = 0;
    if (zm.cnt > 0) combine(acc, p);
  }

  scan_tail<T>(acc, partials, ticket, decision, sel, var, ld, arm_offset, nullptr);
}

// ---- screening keys -----------------------------------------------------------------------------
// zeta is a non-negative float, so the bit pattern of an approximate quotient orders like the
// quotient itself: the row minimum is located on 32-bit integer keys (one reciprocal approximation,
// one multiply, one integer min per entry) and only the entries whose key lies within kKeySlack of
// the smallest key — a superset of the true minimisers, normally a single entry — pay for the
// exact IEEE quotient the reference computes.  fp64 key: high word of sq * rcp.approx(den)
// (MUFU.RCP64H, ~2^-19 relative; truncation to the high word 2^-20); fp32 key: bits of
// sq * rcp.approx(den) (2 ulp).  kKeySlack = 64 key-ulps >= 3e-5 relative covers both with a wide
// margin.  NaNs compare as the largest unsigned keys and never become candidates.
constexpr unsigned int kKeyNone = 0xffffffffu;   // sentinel columns and the best arm itself
constexpr unsigned int kKeySlack = 64;
__device__ __forceinline__ unsigned int zeta_key(double mu_b, double var_b, double m, double vr) {
  const double gap = mu_b - m;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(var_b + vr));
  return (unsigned int)__double2hiint((gap * gap) * r);
}
__device__ __forceinline__ unsigned int zeta_key(float mu_b, float var_b, float m, float vr) {
  const float gap = mu_b - m;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(var_b + vr));
  return __float_as_uint((gap * gap) * r);
}

// Wide rows (256 < K <= 1024): a warp per row, rows staged in shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier), NS stages per warp.  The row stays in shared memory while it is
// scanned and nothing row-sized lives in registers (~50 registers per thread, so 12-28 warps per SM
// cover the TMA latency and the long per-row dependency chains):
//   pass 1  arg-max of the means (lane-strided reads, shuffle reduction); the best arm's entry is
//           then overwritten with -inf in the stage, which turns its zeta key into +inf — no
//           per-entry "is this the best arm" test in the hot loop;
//   pass 2  32-bit screening keys, written over the means in the stage, running minimum,
//           REDUX.MIN across the warp;
//   rare    lanes whose minimum is within the slack re-read their keys and take the exact IEEE
//           quotient of the (normally single) candidate, its mean re-read from global memory (L2).
// Warps never synchronise with each other; the CTA only meets in scan_tail.
template <typename T> struct InfKey;
template <> struct InfKey<double> { static constexpr unsigned int value = 0x7ff00000u; };
template <> struct InfKey<float> { static constexpr unsigned int value = 0x7f800000u; };

constexpr int kRowWarps = 4;  // warps (rows in flight) per CTA

template <typename T, int R, int NS>
__global__ void __launch_bounds__(32 * kRowWarps, 6)
ocba_scan_row_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K,
                     int64_t ld, int arm_offset, const int32_t* __restrict__ ext_arm,
                     const T* __restrict__ ext_mean, const T* __restrict__ ext_var,
                     int32_t* __restrict__ best_arm, T* __restrict__ min_zeta,
                     int32_t* __restrict__ min_arm, int32_t* __restrict__ tie_cnt,
                     ScanPartial* partials, unsigned int* ticket, crs_decision* decision,
                     SelectArgs sel) {
  constexpr int kpad = R * 32;
  griddep_wait();  // the posterior grid of the same decision may still be running (PDL)
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  // dynamic smem: [select stage (kSelStageFused bytes, only if fused)] | per warp: NS x {m[kpad], v[kpad]}
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* wbase = reinterpret_cast<T*>(dyn_smem + (sel.do_select ? kSelStageFused : 0)) + (size_t)warp * NS * 2 * kpad;
  __shared__ __align__(8) uint64_t bars[kRowWarps][2];
  const uint32_t row_bytes = (uint32_t)(K * sizeof(T));
  ScanPartial acc = empty_partial();
  for (int s = 0; s < NS; ++s)
    for (int i = K + lane; i < kpad; i += 32) {  // sentinels behind the row: never best, key = +inf
      wbase[s * 2 * kpad + i] = -pos_inf<T>();
      wbase[s * 2 * kpad + kpad + i] = T(1);
    }
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int64_t c, int st) {  // lane 0 only
    T* sm = wbase + st * 2 * kpad;
    mbar_expect_tx(&bars[warp][st], 2 * row_bytes);
    tma_bulk_g2s(sm, mean + c * ld, row_bytes, &bars[warp][st]);
    tma_bulk_g2s(sm + kpad, var + c * ld, row_bytes, &bars[warp][st]);
  };
  const int64_t row0 = (int64_t)blockIdx.x * kRowWarps + warp, step = (int64_t)gridDim.x * kRowWarps;
  if (lane == 0)
    for (int s = 0; s < NS; ++s)
      if (row0 + s * step < n_c) issue(row0 + s * step, s);
  int it = 0;
  for (int64_t c = row0; c < n_c; c += step, ++it) {
    const int st = NS == 2 ? (it & 1) : 0;
    mbar_wait(&bars[warp][st], (uint32_t)(NS == 2 ? (it >> 1) : it) & 1u);
    T* rm = wbase + st * 2 * kpad;
    const T* rv = rm + kpad;
    T bm, vb;
    int bi;
    if (ext_arm == nullptr) {
      bm = -pos_inf<T>();
      bi = INT_MAX;
#pragma unroll 8
      for (int j = 0; j < R; ++j) {
        const T m = rm[j * 32 + lane];
        if (m > bm) { bm = m; bi = j * 32 + lane; }
      }
      warp_argmax(bm, bi);
      if (bi == INT_MAX) bi = 0;  // all-NaN row: arm 0 by convention
      bm = rm[bi];
      vb = rv[bi];
    } else {
      bi = ext_arm[c] - arm_offset;
      bm = ext_mean[c];
      vb = ext_var[c];
    }
    __syncwarp();
    if (lane == 0 && bi >= 0 && bi < K) rm[bi] = -pos_inf<T>();  // the best arm never competes with itself
    __syncwarp();
    unsigned int* rk = reinterpret_cast<unsigned int*>(rm);  // key of entry a at word a * (sizeof(T) / 4)
    constexpr int kw = sizeof(T) / 4;
    unsigned int kmin = kKeyNone;
#pragma unroll 8
    for (int j = 0; j < R; ++j) {
      const int a = j * 32 + lane;
      const unsigned int key = zeta_key(bm, vb, rm[a], rv[a]);
      rk[a * kw] = key;
      kmin = min(kmin, key);
    }
    const unsigned int wmin = __reduce_min_sync(kFull, kmin);
    // candidates: finite keys within the slack of the warp minimum (inf / NaN keys never qualify)
    unsigned int thr = wmin + kKeySlack;
    if (thr >= InfKey<T>::value || thr < wmin) thr = InfKey<T>::value - 1u;
    ZMin<T> zm;
    zm.init();
    if (kmin <= thr) {
      const T* grow = mean + c * ld;
      for (int j = 0; j < R; ++j) {
        const int a = j * 32 + lane;
        if (rk[a * kw] <= thr) zm.add_exact(bm, vb, grow[a], rv[a], a);
      }
    }
    const unsigned has = __ballot_sync(kFull, zm.cnt > 0);
    if (__popc(has) == 1) {  // the common case: broadcast instead of a full reduction
      const int src = __ffs(has) - 1;
      zm.z = __shfl_sync(kFull, zm.z, src);
      zm.mu = __shfl_sync(kFull, zm.mu, src);
      zm.col = __shfl_sync(kFull, zm.col, src);
      zm.cnt = __shfl_sync(kFull, zm.cnt, src);
    } else if (has) {
      zm.warp_reduce();
    }
    for (int i = K + lane; i < kpad; i += 32) rm[i] = -pos_inf<T>();  // the keys overwrote the sentinels too
    __syncwarp();  // every lane is done with the stage
    if (lane == 0 && c + NS * step < n_c) {
      fence_proxy_async();
      issue(c + NS * step, st);
    }
    const int gbest = bi + arm_offset;
    const int gz = zm.col == INT_MAX ? -1 : zm.col + arm_offset;
    if (lane == 0) {
      if (best_arm) best_arm[c] = gbest;
      if (min_zeta) min_zeta[c] = zm.z;
      if (min_arm) min_arm[c] = gz;
      if (tie_cnt) tie_cnt[c] = zm.cnt;
    }
    ScanPartial p;
    p.z = (double)zm.z;
    p.cnt = zm.cnt;
    p.ctx = (int)c;
    p.zarm = gz;
    p.best = gbest;
    p.pad = 0;
    if (zm.cnt > 0) combine(acc, p);
  }
  scan_tail<T>(acc, partials, ticket, decision, sel, var, ld, arm_offset, dyn_smem);
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

template <typename T>
__global__ void __launch_bounds__(1024)
ocba_select_kernel(crs_gp_state st, const T* __restrict__ ctx, const T* __restrict__ var,
                   int64_t ld, int arm_offset, int p_mode, T kernel_scale, crs_decision* decision,
                   crs_decision* host_mirror, int seq, int stage_bytes, crs_peer_exchange px) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  select_block<T>(st, ctx, var, ld, arm_offset, p_mode, kernel_scale, decision, dyn_smem, stage_bytes);
  if (px.world > 1) peer_exchange_fold(decision, px);
  if (host_mirror && threadIdx.x == 0) publish(decision, host_mirror, seq);
}

// An empty shard (a rank without contexts) still takes part in the exchange.
__global__ void __launch_bounds__(64)
peer_exchange_only_kernel(crs_decision* decision, crs_decision* host_mirror, int seq, crs_peer_exchange px) {
  if (threadIdx.x == 0) {
    decision->min_zeta = CUDART_INF;
    decision->psi_best = decision->psi_rest = 0.0;
    decision->tie_count = 0;
    decision->context = decision->zeta_arm = decision->best_arm = decision->next_arm = -1;
    decision->reserved[0] = decision->reserved[1] = decision->reserved[2] = decision->reserved[3] = 0;
  }
  if (px.world > 1) peer_exchange_fold(decision, px);
  if (host_mirror && threadIdx.x == 0) publish(decision, host_mirror, seq);
}

template <typename T>
int scan_dispatch(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int arm_offset,
                  const int32_t* ext_best_arm, const void* ext_best_mean, const void* ext_best_var,
                  int32_t* best_arm, void* min_zeta, int32_t* min_arm, int32_t* tie_cnt,
                  crs_decision* decision, void* workspace, const SelectArgs& sel, cudaStream_t s) {
  // K <= 1024: the whole row sits in registers (all loads of a row in flight at once — the kernel
  // is HBM-latency bound otherwise, ncu r01a: 66 % long-scoreboard stalls, 8 % L1 hit on the
  // re-read).  K <= 256: one warp per row; wider rows: the 4 warps of a 128-thread CTA share a row.
  const bool wide = K > 256 && K <= 1024;
  const int threads = wide ? 128 : kScanThreads;
  const int wpb = threads / 32;
  int64_t want = wide ? n_c : (n_c + wpb - 1) / wpb;
  const int max_blocks = wide ? 148 * 16 : kScanMaxBlocks;  // partials[] holds kScanMaxBlocks * 2
  int blocks = (int)(want < max_blocks ? want : max_blocks);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  ScanPartial* partials = workspace ? reinterpret_cast<ScanPartial*>(reinterpret_cast<char*>(workspace) + 64) : nullptr;
  const size_t dyn = sel.do_select ? kSelStageFused : 0;
  auto go = [&](auto kern) {
    (void)launch_pdl(kern, dim3(blocks), dim3(threads), dyn, s, (const T*)mean, (const T*)var, n_c, K, ld, arm_offset,
                     ext_best_arm, (const T*)ext_best_mean, (const T*)ext_best_var, best_arm, (T*)min_zeta,
                     min_arm, tie_cnt, partials, ticket, decision, sel);
  };
  const size_t esz = sizeof(T);
  const int seg = K <= 512 ? 4 : (K <= 768 ? 6 : 8);       // kpad = 128 seg >= K
  const size_t stage = 2 * (size_t)seg * 128 * esz;         // one {mean row, var row} pair
  const int ns = stage <= 4 * 1024 ? 2 : 1;                 // double-buffer a warp's rows only while >= 24 warps/SM still fit
  const size_t tma_smem = dyn + (size_t)kRowWarps * ns * stage;
  const bool tma_ok = K > 256 && K <= 1024 && (K * esz) % 16 == 0 && (ld * esz) % 16 == 0 &&
                      ((uintptr_t)mean % 16 == 0) && ((uintptr_t)var % 16 == 0);
  if (tma_ok) {
    auto launch = [&](auto kern) -> int {
      if (tma_smem + 8 * 1024 > 48 * 1024)  // static shared memory (~4 KB) counts against the 48 KB default
        CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem), "scan attr");
      int per_sm = (int)((224 * 1024) / (tma_smem + 1024 + 4608));
      if (per_sm > 8) per_sm = 8;
      if (per_sm < 1) per_sm = 1;
      const int64_t want_rows = (n_c + kRowWarps - 1) / kRowWarps;
      const int tblocks = (int)(want_rows < 148 * per_sm ? want_rows : 148 * per_sm);
      CRS_CUDA_TRY(launch_pdl(kern, dim3(tblocks), dim3(32 * kRowWarps), tma_smem, s, (const T*)mean, (const T*)var, n_c, K, ld,
                              arm_offset, ext_best_arm, (const T*)ext_best_mean, (const T*)ext_best_var, best_arm,
                              (T*)min_zeta, min_arm, tie_cnt, partials, ticket, decision, sel), "ocba_scan launch");
      return CRS_OK;
    };
    int rc;
    if (ns == 2) rc = seg == 4 ? launch(ocba_scan_row_kernel<T, 16, 2>) : (seg == 6 ? launch(ocba_scan_row_kernel<T, 24, 2>) : launch(ocba_scan_row_kernel<T, 32, 2>));
    else rc = seg == 4 ? launch(ocba_scan_row_kernel<T, 16, 1>) : (seg == 6 ? launch(ocba_scan_row_kernel<T, 24, 1>) : launch(ocba_scan_row_kernel<T, 32, 1>));
    if (rc != CRS_OK) return rc;
  } else if (K <= 32) go(ocba_scan_kernel<T, 1, 1>);
  else if (K <= 64) go(ocba_scan_kernel<T, 2, 1>);
  else if (K <= 128) go(ocba_scan_kernel<T, 4, 1>);
  else if (K <= 256) go(ocba_scan_kernel<T, 8, 1>);
  else if (K <= 512) go(ocba_scan_kernel<T, 4, 4>);
  else if (K <= 1024) go(ocba_scan_kernel<T, 8, 4>);
  else go(ocba_scan_kernel<T, 0, 1>);
  CRS_CUDA_TRY(cudaGetLastError(), "ocba_scan launch");
  return CRS_OK;
}

}  // namespace

int64_t scan_workspace_bytes() { return (int64_t)kScanMaxBlocks * 2 * sizeof(ScanPartial) + 64; }

int launch_ocba_scan_select(const void* mean, const void* var, int64_t n_c, int K, int64_t ld,
                            int dtype, int arm_offset, const int32_t* ext_best_arm,
                            const void* ext_best_mean, const void* ext_best_var, int32_t* best_arm,
                            void* min_zeta, int32_t* min_arm, int32_t* tie_cnt,
                            crs_decision* decision, void* workspace, const crs_gp_state* st,
                            const void* ctx, int p_mode, double kernel_scale,
                            crs_decision* host_mirror, int seq, const crs_peer_exchange* px, cudaStream_t s) {
  if (!mean || !var || n_c < 0 || n_c > INT_MAX || K < 1 || ld < K) return CRS_ERR_INVALID_ARG;
  if (px && (!mailbox::valid(px) || (px->world > 1 && !st))) return CRS_ERR_INVALID_ARG;
  if (decision && !workspace) return CRS_ERR_INVALID_ARG;
  if (ext_best_arm && (!ext_best_mean || !ext_best_var)) return CRS_ERR_INVALID_ARG;
  if (st && (!ctx || !decision || p_mode < 0 || p_mode > 2 || st->num_arms != K)) return CRS_ERR_INVALID_ARG;
  if (host_mirror && !decision) return CRS_ERR_INVALID_ARG;
  if (n_c == 0) return CRS_OK;
  SelectArgs sel;
  memset(&sel, 0, sizeof(sel));
  // Fusing step 2(b) into the scan's last CTA saves a launch, but its 32 KB stage buffer shrinks the
  // L1 the K > 256 row re-read relies on: fuse only where the decision is latency-bound anyway.
  const bool fuse = st && (n_c * (int64_t)K <= (1 << 20));
  if (fuse) {
    sel.st = *st;
    sel.ctx = ctx;
    sel.kernel_scale = kernel_scale;
    sel.p_mode = p_mode;
    sel.do_select = 1;
    sel.host_mirror = host_mirror;
    sel.seq = seq;
    if (px && px->world > 1) sel.px = *px;
  } else if (!st) {
    sel.host_mirror = host_mirror;
    sel.seq = seq;
  }
  int rc;
  if (dtype == CRS_F64)
    rc = scan_dispatch<double>(mean, var, n_c, K, ld, arm_offset, ext_best_arm, ext_best_mean, ext_best_var,
                               best_arm, min_zeta, min_arm, tie_cnt, decision, workspace, sel, s);
  else if (dtype == CRS_F32)
    rc = scan_dispatch<float>(mean, var, n_c, K, ld, arm_offset, ext_best_arm, ext_best_mean, ext_best_var,
                              best_arm, min_zeta, min_arm, tie_cnt, decision, workspace, sel, s);
  else
    return CRS_ERR_INVALID_ARG;
  if (rc != CRS_OK || !st || fuse) return rc;
  return launch_ocba_select(st, ctx, mean, var, ld, arm_offset, p_mode, kernel_scale, decision, host_mirror, seq, px, s);
}

int launch_peer_exchange_only(crs_decision* decision, crs_decision* host_mirror, int seq, const crs_peer_exchange* px,
                              cudaStream_t s) {
  if (!decision || !mailbox::valid(px)) return CRS_ERR_INVALID_ARG;
  peer_exchange_only_kernel<<<1, 64, 0, s>>>(decision, host_mirror, seq, *px);
  CRS_CUDA_TRY(cudaGetLastError(), "peer_exchange launch");
  return CRS_OK;
}

int launch_ocba_scan(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                     int arm_offset, const int32_t* ext_best_arm, const void* ext_best_mean,
                     const void* ext_best_var, int32_t* best_arm, void* min_zeta, int32_t* min_arm,
                     int32_t* tie_cnt, crs_decision* decision, void* workspace, cudaStream_t s) {
  return launch_ocba_scan_select(mean, var, n_c, K, ld, dtype, arm_offset, ext_best_arm, ext_best_mean,
                                 ext_best_var, best_arm, min_zeta, min_arm, tie_cnt, decision, workspace,
                                 nullptr, nullptr, 0, 0.0, nullptr, 0, nullptr, s);
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
                       crs_decision* decision, crs_decision* host_mirror, int seq, const crs_peer_exchange* px,
                       cudaStream_t s) {
  (void)mean;
  if (!st || !ctx || !var || !decision || p_mode < 0 || p_mode > 2) return CRS_ERR_INVALID_ARG;
  crs_peer_exchange pxv;
  memset(&pxv, 0, sizeof(pxv));
  if (px && px->world > 1) pxv = *px;
  const int threads = st->num_arms >= 32 ? 1024 : (st->num_arms >= 8 ? 256 : 64);
  const size_t blk = (size_t)st->n_cap * st->dim * (st->dtype == CRS_F64 ? 8 : 4);
  size_t stage = blk * (size_t)(st->num_arms < 256 ? st->num_arms : 256);
  if (stage > (size_t)kSelStageAlone) stage = kSelStageAlone;
  if (stage < blk) stage = blk;
  if (st->dtype == CRS_F64) {
    auto kern = ocba_select_kernel<double>;
    if (stage > 40 * 1024) CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage), "select attr");
    kern<<<1, threads, stage, s>>>(*st, (const double*)ctx, (const double*)var, ld, arm_offset, p_mode, kernel_scale, decision, host_mirror, seq, (int)stage, pxv);
  } else if (st->dtype == CRS_F32) {
    auto kern = ocba_select_kernel<float>;
    if (stage > 40 * 1024) CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage), "select attr");
    kern<<<1, threads, stage, s>>>(*st, (const float*)ctx, (const float*)var, ld, arm_offset, p_mode, (float)kernel_scale, decision, host_mirror, seq, (int)stage, pxv);
  } else {
    return CRS_ERR_INVALID_ARG;
  }
  CRS_CUDA_TRY(cudaGetLastError(), "ocba_select launch");
  return CRS_OK;
}

}  // namespace crs
