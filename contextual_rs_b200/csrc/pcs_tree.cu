// crs_pcs_counts_tree (K3, max-tree stream): Monte-Carlo correct-selection counts per context —
// the sampling half of estimate_current_generalized_pcs for a ModelListGP
// (contextual_rs/generalized_pcs.py:441-475): does the predicted-best arm's posterior draw beat the
// draws of all other arms (func_I = (delta > 0), experiments/wsc_experiments/main.py:453)?
//
// A sample fails as soon as ONE competitor beats y_b and succeeds only if NONE does, and with
// K = 1,000 diffuse posteriors almost every sample fails.  The flat stream (pcs_philox.cu) has to
// draw arms until it meets a winner (3-4 Philox calls) and all K of them to certify a success
// (250 calls).  This kernel samples the same law top-down instead: the competitors are sorted by
// threat, the five most threatening ones are drawn directly and the others are the leaves of a
// 4-ary tree.  A node covering m leaves carries t = -ln P(max of its m iid N(0,1) draws <= its
// value) (Exp(1)/m at the root, one uniform) and the leaf J that attains the maximum (uniform).
// Children of a node: the one containing J inherits (t, J), the others draw the maximum of m/4
// normals truncated at the parent's value, t_c = t + Exp(1)/(m/4), and their own J_c.
// Stream definition and brute-force restatement: oracle/pcs_tree_ref.py.
//   root     one Philox call per sample: y_b, the draw of the most threatening competitor (the
//            second Box-Muller variate, free), the largest of the K-6 standard normals of the
//            tree and the leaf that owns it.  If either arm reaches y_b the sample has failed
//            (60-90 % of the samples at 1,000 arms end here).
//   direct   samples that survive draw the next four threats directly (one more Philox call, two
//            Box-Muller pairs): most failures the root missed end here, before any tree work.
//   expand   otherwise the node is expanded (one Philox call, three inverse-normal evaluations),
//            each new child's arg-max leaf is tested, and only children whose bound reaches y_b
//            are kept (per-lane stack in shared memory).  The bound of a node is the convex
//            envelope g(z) = max_j (mu_j + sd_j z) of its leaves, tabulated at z = 0, 1.8, .. 7.2
//            and interpolated by chords (a chord of a convex function lies above it; the
//            interpolant is the max of the four chord lines: 4 FMA + 3 max) — exact pruning,
//            because zfun is non-increasing in t.  A success costs ~6-10 expansions at
//            K = 1,000 (30-60 with the cruder bound mu_max + sd_max z).
//   pool     per warp: root tests run warp-wide (no divergence) and park their survivors in a
//            shared-memory queue; lanes walk one survivor each and take the next one when done; when
//            the queue is empty and few lanes are still walking, the warp goes back to root tests.
#include <atomic>
#include <mutex>
#include <math_constants.h>
#include "crs_common.cuh"
#include "pcs_rng.cuh"

namespace crs {
namespace {

using namespace rng;

constexpr int kTreeThreads = 256;
constexpr int kTreeWarps = kTreeThreads / 32;
#ifndef CRS_TREE_QUEUE
#define CRS_TREE_QUEUE 96
#endif
#ifndef CRS_TREE_MIN_ACTIVE
#define CRS_TREE_MIN_ACTIVE 28
#endif
constexpr int kTreeQueue = CRS_TREE_QUEUE;            // root survivors parked per warp
constexpr int kTreeMinActive = CRS_TREE_MIN_ACTIVE;   // fewer walking lanes than this (and fresh samples left): refill first
constexpr int kTreeMaxLeaves = 4096;    // 10-bit arg-max fields per child: depth <= 6
constexpr int kTreeDirect = 5;          // most threatening competitors, drawn outside the tree
constexpr uint32_t kTreeDirectNode = 0xffffffffu;  // Philox counter word of the call that draws direct arms 1..4
constexpr float kTreeGrid = 1.8f;       // spacing of the envelope table (zfun < 6.9 < 4 * 1.8)
constexpr float kTreeEmpty = -1e30f;    // envelope of a node without leaves
constexpr int kTreeSplitSamples = 8192; // smallest sample range worth a plan build of its own
constexpr float kNegLn2 = -0.6931471805599453f;

__device__ __forceinline__ float approx_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// z with Phi(z) = exp(-t), t >= 0 (|error| < 3e-6, non-increasing in t; fp32 operation sequence
// mirrored by oracle/pcs_tree_ref.py::zfun32).  p = smaller tail probability (1 - e^-t by series for
// t < 1/8), w = -ln(4 p (1 - p)), z = +-(1 - 2p) * poly, poly ~ sqrt(2) erfinv(x) / x in (w - 3.125)
// for w < 6.25 and in (sqrt(w) - 3.848) up to w = 27 (|z| < 6.9).
// kSmallT: the caller guarantees t < 0.125 (roots of trees with >= 256 leaves) — same values, fewer
// instructions.
template <bool kSmallT = false>
__device__ __forceinline__ float zfun(float t) {
  const bool upper = kSmallT || t < 0.6931471805599453f;
  const float P = kSmallT ? 0.f : approx_exp2(t * -1.4426950408889634f);
  float qp = 0.00833333377f;
  qp = fmaf(qp, t, -0.0416666679f);
  qp = fmaf(qp, t, 0.166666672f);
  qp = fmaf(qp, t, -0.5f);
  qp = fmaf(qp, t, 1.f);
  const float q = (kSmallT || t < 0.125f) ? __fmul_rn(qp, t) : 1.f - P;
  const float p = upper ? q : P;
  const float a = fmaf(-p, p, p);
  float w = fmaf(approx_log2(a), kNegLn2, -1.3862943611198906f);
  w = fminf(w, 27.f);  // w >= -1e-6 (rounding); a NaN square root below is never selected
  const float u = w - 3.125f;
  float ec = 4.58798354e-07f;
  ec = fmaf(ec, u, -2.39523933e-06f);
  ec = fmaf(ec, u, -1.85905155e-05f);
  ec = fmaf(ec, u, 0.000266286603f);
  ec = fmaf(ec, u, -0.00105035095f);
  ec = fmaf(ec, u, -0.00853662658f);
  ec = fmaf(ec, u, 0.339637041f);
  ec = fmaf(ec, u, 2.33862185f);
  const float v = approx_sqrt(w) - 3.84807611f;
  float et = 4.68631297e-05f;
  et = fmaf(et, v, -0.000144537815f);
  et = fmaf(et, v, 0.000162138473f);
  et = fmaf(et, v, -0.000207302306f);
  et = fmaf(et, v, 0.000602424378f);
  et = fmaf(et, v, -0.0015758829f);
  et = fmaf(et, v, 0.00241446355f);
  et = fmaf(et, v, 1.42700899f);
  et = fmaf(et, v, 5.21342754f);
  const float e = w < 6.25f ? ec : et;
  const float z = __fmul_rn(e, fmaf(-2.f, p, 1.f));
  return upper ? z : -z;
}

// A node's bound: the four chord lines a_k + s_k z of its envelope table g(0), g(1.8), .. g(7.2).
// The interpolant of a convex function is the maximum of its (extended) chords; z < 0: g(0).
struct __align__(16) NodeLines { float4 a, s; };
__device__ __forceinline__ float node_bound(const NodeLines& nd, float z) {
  const float zp = fmaxf(z, 0.f);
  return fmaxf(fmaxf(fmaf(nd.s.x, zp, nd.a.x), fmaf(nd.s.y, zp, nd.a.y)),
               fmaxf(fmaf(nd.s.z, zp, nd.a.z), fmaf(nd.s.w, zp, nd.a.w)));
}
// lines from the table g[0..4] (grid spacing kTreeGrid)
__device__ __forceinline__ NodeLines make_lines(const float g[5]) {
  NodeLines nd;
  const float inv = 1.f / kTreeGrid;
  nd.s = make_float4((g[1] - g[0]) * inv, (g[2] - g[1]) * inv, (g[3] - g[2]) * inv, (g[4] - g[3]) * inv);
  nd.a = make_float4(g[0], fmaf(-nd.s.y, kTreeGrid, g[1]), fmaf(-nd.s.z, 2.f * kTreeGrid, g[2]),
                     fmaf(-nd.s.w, 3.f * kTreeGrid, g[3]));
  return nd;
}

// index of the first node of level l: (4^l - 1) / 3 = 0b0101...01 (l ones)
__device__ __forceinline__ uint32_t level_base(int l) { return 0x55555555u & ((1u << (2 * l)) - 1u); }
// -ln 2 * 2^-sh, exact (exponent arithmetic)
__device__ __forceinline__ float neg_ln2_scaled(int sh) {
  return __uint_as_float(__float_as_uint(kNegLn2) - ((uint32_t)sh << 23));
}

// Work-item tickets: one 16-byte slot {next item, CTAs done} per launch, taken round-robin from a
// ring.  The last CTA to leave a launch zeroes its slot, so a slot is clean whenever the previous
// launch that used it has finished; two launches could only share a slot if kTreeTicketRing launches
// of this kernel were in flight on one device at once.
constexpr int kTreeTicketRing = 1024;
__device__ unsigned long long g_tree_tickets[2 * kTreeTicketRing];

#ifndef CRS_TREE_MIN_BLOCKS
#define CRS_TREE_MIN_BLOCKS 3
#endif

template <typename T>
__global__ void __launch_bounds__(kTreeThreads, CRS_TREE_MIN_BLOCKS)
pcs_tree_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K, int64_t ld,
                int64_t num_samples, const __grid_constant__ RoundKeys rk, uint32_t offset,
                int64_t ctx_offset, int splits, int D, int n2, int slots,
                unsigned long long* __restrict__ ticket, uint32_t* counts, unsigned long long* stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NL = 1 << (2 * D);
  const int NI = (NL - 1) / 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // [ leaves (mu, sd) | nodes (envelope table) | sort scratch  ==alias==  stacks + queues ]
  float2* leaf = reinterpret_cast<float2*>(smem_raw);
  NodeLines* node = reinterpret_cast<NodeLines*>(leaf + NL);
  unsigned char* region = reinterpret_cast<unsigned char*>(node + NI);
  float* s_key = reinterpret_cast<float*>(region);
  int* s_ord = reinterpret_cast<int*>(s_key + n2);
  float* s_mu = reinterpret_cast<float*>(s_ord + n2);                // centred mean / std-dev of competitor e
  float* s_sd = s_mu + n2;
  float* s_g = reinterpret_cast<float*>(region);                    // [NI][5] envelope tables while the nodes are built
  uint32_t* stk_meta = reinterpret_cast<uint32_t*>(region);          // [slots][threads]: level<<24 | index<<12 | J
  float* stk_t = reinterpret_cast<float*>(stk_meta + slots * kTreeThreads);
  float* stk_z = stk_t + slots * kTreeThreads;
  uint32_t* q_s = reinterpret_cast<uint32_t*>(stk_z + slots * kTreeThreads) + warp * (5 * kTreeQueue);
  float* q_yb = reinterpret_cast<float*>(q_s + kTreeQueue);
  float* q_t = q_yb + kTreeQueue;
  uint32_t* q_J = reinterpret_cast<uint32_t*>(q_t + kTreeQueue);
  float* q_z = reinterpret_cast<float*>(q_J + kTreeQueue);
  __shared__ T s_bm[kTreeWarps];
  __shared__ int s_bi[kTreeWarps];
  __shared__ int s_best;
  __shared__ float s_sdb;
  __shared__ float2 s_top[kTreeDirect];
  __shared__ unsigned int s_hits;
  __shared__ long long s_item;
  const int64_t items = n_c * splits;
  const int nsort = K - 1;                                        // competitors
  const int nleaf = nsort > kTreeDirect ? nsort - kTreeDirect : 0;  // all but the five most threatening ones
  const uint32_t jmask = (uint32_t)NL - 1u;
  const float c_root = neg_ln2_scaled(2 * D);
  const bool small_root = D >= 4;  // t_root <= 16.7 / 256
  const unsigned int lt_mask = (1u << lane) - 1u;
  unsigned long long n_roots = 0, n_exp = 0, n_direct = 0;

  for (;;) {
    if (tid == 0) s_item = (long long)atomicAdd(ticket, 1ull);
    __syncthreads();
    const int64_t item = s_item;
    if (item >= items) break;
    const int64_t c = item / splits;
    const int sp_i = (int)(item - c * splits);
    const int64_t s_lo = num_samples * sp_i / splits, s_hi = num_samples * (sp_i + 1) / splits;
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    // ---- predicted best arm: first arg-max of the mean row (generalized_pcs.py:464) ----------
    T bm = -CUDART_INF;
    int bi = 0x7fffffff;
    for (int a = tid; a < K; a += kTreeThreads) {
      const T m = mrow[a];
      if (m > bm) { bm = m; bi = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T om = __shfl_xor_sync(kFull, bm, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
    }
    if (lane == 0) { s_bm[warp] = bm; s_bi[warp] = bi; }
    if (tid == 0) s_hits = 0;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kTreeWarps; ++w)
        if (s_bm[w] > bm || (s_bm[w] == bm && s_bi[w] < bi)) { bm = s_bm[w]; bi = s_bi[w]; }
      if (bi == 0x7fffffff) bi = 0;
      s_best = bi;
      s_bm[0] = mrow[bi];
      s_sdb = (float)sqrt((double)vrow[bi]);
    }
    __syncthreads();
    const int best = s_best;
    const T mu_b = s_bm[0];
    const float sd_b = s_sdb;
    // ---- leaves in threat order: IEEE fp32 key, ties by arm id (reproducible on the host) ------
    const float sdb2 = __fmul_rn(sd_b, sd_b);
    for (int e = tid; e < n2; e += kTreeThreads) {
      float key = -CUDART_INF_F;
      int arm = 0x7fffffff;
      if (e < nsort) {
        arm = e + (e >= best ? 1 : 0);
        const float mu = (float)(mrow[arm] - mu_b);  // centred in the grid dtype, then fp32
        float sd;
        if constexpr (sizeof(T) == 4) sd = __fsqrt_rn((float)vrow[arm]);  // == (float)sqrt((double)v)
        else sd = (float)sqrt((double)vrow[arm]);
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(sd, sd), sdb2), 1e-30f);
        key = __fdiv_rn(mu, __fsqrt_rn(s));
        s_mu[e] = mu;
        s_sd[e] = sd;
      }
      s_key[e] = key;
      s_ord[e] = arm;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n2; i += kTreeThreads) {
          const int p = i ^ j;
          if (p > i) {
            const float ka = s_key[i], kb = s_key[p];
            const int oa = s_ord[i], ob = s_ord[p];
            const bool a_first = ka > kb || (ka == kb && oa < ob);
            const bool desc = (i & k) == 0;
            if (desc != a_first) {
              s_key[i] = kb; s_key[p] = ka;
              s_ord[i] = ob; s_ord[p] = oa;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int p = tid; p < NL + kTreeDirect; p += kTreeThreads) {  // sorted positions 0..4 = direct arms, then the leaves
      float2 lf = make_float2(-CUDART_INF_F, 0.f);
      if (p < nsort) {
        const int arm = s_ord[p];
        const int e = arm - (arm > best ? 1 : 0);
        lf.x = s_mu[e];
        lf.y = s_sd[e];
      }
      if (p < kTreeDirect) s_top[p] = lf;
      else leaf[p - kTreeDirect] = lf;
    }
    __syncthreads();  // the sort scratch is dead from here on (aliased by stacks and queues)
    for (int l = D - 1; l >= 0; --l) {
      const int nn = 1 << (2 * l);
      const uint32_t b0 = level_base(l), b1 = level_base(l + 1);
      for (int i = tid; i < nn; i += kTreeThreads) {
        float g[5] = {kTreeEmpty, kTreeEmpty, kTreeEmpty, kTreeEmpty, kTreeEmpty};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (l == D - 1) {
            const int p = 4 * i + k;
            if (p < nleaf) {
              const float2 lf = leaf[p];
#pragma unroll
              for (int q = 0; q < 5; ++q) g[q] = fmaxf(g[q], fmaf(lf.y, (float)q * kTreeGrid, lf.x));
            }
          } else {
#pragma unroll
            for (int q = 0; q < 5; ++q) g[q] = fmaxf(g[q], s_g[5 * (b1 + 4 * i + k) + q]);
          }
        }
        if (l == D - 1) {  // head-room for the rounding of the lines (g is non-decreasing: |g| <= max(|g0|, |g4|))
          const float pad = fmaxf(fabsf(g[0]), fabsf(g[4])) * 3.814697265625e-06f;  // 2^-18
#pragma unroll
          for (int q = 0; q < 5; ++q) g[q] += pad;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) s_g[5 * (b0 + i) + q] = g[q];
        node[b0 + i] = make_lines(g);
      }
      __syncthreads();
    }
    const uint32_t cw = (uint32_t)(c + ctx_offset);
    const NodeLines root = node[0];
    __syncthreads();  // s_g (aliased by stacks and queues) is dead from here on
    const float2 top = s_top[0];

    // ---- sampling -------------------------------------------------------------------------------
    const unsigned int n_item = (unsigned int)(s_hi - s_lo);
    const unsigned int w_hi = (unsigned int)((uint64_t)n_item * (warp + 1) / kTreeWarps);
    unsigned int wnext = (unsigned int)((uint64_t)n_item * warp / kTreeWarps);
    unsigned int hits = 0, steps = 0, directs = 0;
    int sp = 0, qn = 0, qdone = 0;  // queue entries [0, qdone) have passed the direct-arm test
    uint32_t smp = 0;
    float yb = 0.f;
    for (;;) {
      // Root tests, warp-wide: y_b, the top threat's draw, the largest tree normal and its owner.
      // Samples that survive are parked; the parked ones then draw direct arms 1..4 (second Philox
      // call, again warp-wide) and only those that still stand — and whose root bound reaches y_b —
      // stay in the queue for the tree walk.
      for (;;) {
        __syncwarp();  // queue slots read by the walk below are rewritten here
        while (qn <= kTreeQueue - 32 && wnext < w_hi) {
          const unsigned int my = wnext + lane;
          const bool valid = my < w_hi;
          const uint32_t s = (uint32_t)(s_lo + (valid ? my : w_hi - 1));
          const Words w = philox4x32_10(s, cw, 0u, offset, rk);
          float zb, ztop;
          box_muller(w.x, w.y, zb, ztop);
          const float ybs = __fmul_rn(sd_b, zb);
          const float t = __fmul_rn(log2_unit_open(w.z), c_root);
          const uint32_t J = w.w & jmask;
          const float z = small_root ? zfun<true>(t) : zfun<false>(t);
          const float2 lf = leaf[J];
          const bool fail = fmaf(top.y, ztop, top.x) >= ybs || fmaf(lf.y, z, lf.x) >= ybs;
          const bool open = node_bound(root, z) >= ybs;  // some leaf may still reach y_b
          const bool surv = valid && !fail;
          const unsigned bal = __ballot_sync(kFull, surv);
          if (surv) {
            const int pos = qn + __popc(bal & lt_mask);
            q_s[pos] = s; q_yb[pos] = ybs; q_t[pos] = t; q_J[pos] = J | (open ? 0u : 0x80000000u); q_z[pos] = z;
          }
          qn += __popc(bal);
          n_roots += lane == 0 ? (unsigned long long)min(32u, w_hi - wnext) : 0ull;
          wnext += 32;
        }
        // direct arms 1..4 of the newly parked samples; the queue is compacted in place
        int wq = qdone;
        for (int base = qdone; base < qn; base += 32) {
          __syncwarp();
          const int idx = base + lane;
          const bool v = idx < qn;
          const int ri = v ? idx : qn - 1;
          const uint32_t s = q_s[ri], Jf = q_J[ri];
          const float ybs = q_yb[ri], t = q_t[ri], z = q_z[ri];
          __syncwarp();  // every lane holds its entry: slots may be overwritten now
          const Words w = philox4x32_10(s, cw, kTreeDirectNode, offset, rk);
          float z1, z2, z3, z4;
          box_muller(w.x, w.y, z1, z2);
          box_muller(w.z, w.w, z3, z4);
          const float2 d1 = s_top[1], d2 = s_top[2], d3 = s_top[3], d4 = s_top[4];
          const float ymax = fmaxf(fmaxf(fmaf(d1.y, z1, d1.x), fmaf(d2.y, z2, d2.x)),
                                   fmaxf(fmaf(d3.y, z3, d3.x), fmaf(d4.y, z4, d4.x)));
          const bool ok = v && !(ymax >= ybs);
          hits += (ok && (Jf >> 31)) ? 1u : 0u;  // no leaf can reach y_b either: success
          const bool keep = ok && !(Jf >> 31);
          const unsigned bal = __ballot_sync(kFull, keep);
          if (keep) {
            const int pos = wq + __popc(bal & lt_mask);
            q_s[pos] = s; q_yb[pos] = ybs; q_t[pos] = t; q_J[pos] = Jf; q_z[pos] = z;
          }
          wq += __popc(bal);
          directs += v ? 1u : 0u;
        }
        qn = qdone = wq;
        if (qn > kTreeQueue - 64 || wnext >= w_hi) break;  // enough parked work, or no fresh samples left
      }
      __syncwarp();
      if (qn == 0 && !__any_sync(kFull, sp > 0)) break;  // no fresh samples left either
      // tree walk: one node expansion per lane and iteration
      for (;;) {
        unsigned act = __ballot_sync(kFull, sp > 0);
        if (qn > 0 && act != kFull) {  // idle lanes take the parked survivors
          const unsigned idle = ~act;
          const int rank = __popc(idle & lt_mask);
          if (sp == 0 && rank < qn) {
            const int idx = qn - 1 - rank;
            smp = q_s[idx];
            yb = q_yb[idx];
            stk_meta[tid] = q_J[idx];  // level 0, index 0
            stk_t[tid] = q_t[idx];
            stk_z[tid] = q_z[idx];
            sp = 1;
          }
          qn -= min(__popc(idle), qn);
          act = __ballot_sync(kFull, sp > 0);
        }
        if (act == 0) break;
        if (qn == 0 && wnext < w_hi && __popc(act) < kTreeMinActive) break;  // refill the queue first

        if (sp > 0) {
          --sp;
          const uint32_t meta = stk_meta[sp * kTreeThreads + tid];
          const float t = stk_t[sp * kTreeThreads + tid];
          const float z = stk_z[sp * kTreeThreads + tid];
          const int l = (int)(meta >> 24);
          const uint32_t i = (meta >> 12) & 0xfffu, J = meta & 0xfffu;
          const int sh = 2 * (D - 1 - l);
          const uint32_t mmask = (1u << sh) - 1u;
          const float c_m = neg_ln2_scaled(sh);
          const uint32_t kp = (J >> sh) & 3u;  // the child that contains the arg-max leaf
          const Words w = philox4x32_10(smp, cw, 1u + level_base(l) + i, offset, rk);
          const bool inner = l + 1 < D;
          const uint32_t cbase = level_base(l + 1) + 4u * i;
          const uint32_t lvl = (uint32_t)(l + 1) << 24;
          bool fail = false;
          ++steps;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const uint32_t k = (uint32_t)r + ((uint32_t)r >= kp ? 1u : 0u);
            const uint32_t word = r == 0 ? w.x : (r == 1 ? w.y : w.z);
            const float tc = fmaf(log2_unit_open(word), c_m, t);
            const uint32_t ci = 4u * i + k;
            const uint32_t Jc = (ci << sh) + ((w.w >> (10 * r)) & mmask);
            const float zc = zfun(tc);
            const float2 lf = leaf[Jc];
            fail = fail || fmaf(lf.y, zc, lf.x) >= yb;
            if (inner) {
              if (node_bound(node[cbase + k], zc) >= yb) {
                stk_meta[sp * kTreeThreads + tid] = lvl | (ci << 12) | Jc;
                stk_t[sp * kTreeThreads + tid] = tc;
                stk_z[sp * kTreeThreads + tid] = zc;
                ++sp;
              }
            }
          }
          if (inner) {  // the child that inherits (t, J): its leaf J is already known to be below y_b
            if (node_bound(node[cbase + kp], z) >= yb) {
              stk_meta[sp * kTreeThreads + tid] = lvl | ((4u * i + kp) << 12) | J;
              stk_t[sp * kTreeThreads + tid] = t;
              stk_z[sp * kTreeThreads + tid] = z;
              ++sp;
            }
          }
          if (fail) sp = 0;
          else if (sp == 0) ++hits;
        }
      }
      qdone = qn;  // what is left in the queue has passed the direct-arm test
    }
    n_exp += steps;
    n_direct += directs;
    hits = warp_sum(hits);
    if (lane == 0 && hits) atomicAdd(&s_hits, hits);
    __syncthreads();
    if (tid == 0 && s_hits) atomicAdd(counts + c, s_hits);
  }
  // leave the ticket slot clean for the next launch that draws it (no per-launch memset)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ticket + 1, 1ull) == (unsigned long long)gridDim.x - 1ull) {
      ticket[0] = 0ull;
      ticket[1] = 0ull;
      __threadfence();
    }
  }
  if (stats) {
    n_roots = warp_sum(n_roots);
    n_exp = warp_sum(n_exp);
    n_direct = warp_sum(n_direct);
    if (lane == 0) {
      atomicAdd(stats, n_roots);
      atomicAdd(stats + 1, n_exp);
      atomicAdd(stats + 2, n_direct);
    }
  }
}

// Peak probes for the roofline of the tree kernel (crs_peak_probe kinds 6 and 7): the bare
// arithmetic of one root test / one node expansion on register operands — no memory, no queueing,
// no control flow besides the loop.  One step per thread and iteration.
__global__ void __launch_bounds__(kTreeThreads)
tree_probe_kernel(int kind, int iters, const __grid_constant__ RoundKeys rk, float* out) {
  const uint32_t smp = blockIdx.x * kTreeThreads + threadIdx.x;
  unsigned int hits = 0;
  const float mu = -0.3f, sd = 1.1f, c_m = neg_ln2_scaled(6);
  if (kind == 6) {
    for (int it = 0; it < iters; ++it) {
      const Words w = philox4x32_10(smp, 7u, (uint32_t)it, 0u, rk);
      float zb, ztop;
      box_muller(w.x, w.y, zb, ztop);
      const float yb = 0.9f * zb;
      const float z = zfun<true>(log2_unit_open(w.z) * c_m);
      const float m2 = __uint_as_float((w.w & 0xffu) | 0x3f000000u);
      hits += (fmaf(sd, ztop, mu) >= yb || fmaf(m2, z, mu) >= yb) ? 1u : 0u;
      hits += (node_bound(NodeLines{make_float4(mu, m2, -1.f, -3.f), make_float4(0.2f, 0.5f, 0.9f, 1.4f)}, z) >= yb) ? 2u : 0u;
    }
  } else {
    float t = 1e-3f;
    for (int it = 0; it < iters; ++it) {
      const Words w = philox4x32_10(smp, 7u, (uint32_t)it, 0u, rk);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const uint32_t word = r == 0 ? w.x : (r == 1 ? w.y : w.z);
        const float zc = zfun(fmaf(log2_unit_open(word), c_m, t));
        const float m2 = __uint_as_float(((w.w >> (10 * r)) & 0xffu) | 0x3f000000u);
        hits += (fmaf(sd, zc, mu) >= 2.5f) ? 1u : 0u;
        hits += (node_bound(NodeLines{make_float4(mu, m2, -1.f, -3.f), make_float4(0.2f, 0.5f, 0.9f, 1.4f)}, zc) >= 2.5f) ? 2u : 0u;
      }
      t = __uint_as_float((hits & 0xffu) | 0x3a000000u);
    }
  }
  if (hits == 0xffffffffu) out[0] = 1.f;  // keep the loop alive
}

__global__ void tree_zfun_kernel(const float* __restrict__ t, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = t[i] < 0.125f && (i & 1) ? zfun<true>(t[i]) : zfun<false>(t[i]);
}

}  // namespace

int launch_tree_zfun(const float* t, int64_t n, float* out, cudaStream_t s) {
  if (!t || !out || n < 0) return CRS_ERR_INVALID_ARG;
  if (n == 0) return CRS_OK;
  tree_zfun_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t, n, out);
  CRS_CUDA_TRY(cudaGetLastError(), "tree_zfun launch");
  return CRS_OK;
}

int launch_pcs_counts_tree(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                           int64_t num_samples, uint64_t seed, uint32_t offset, int64_t ctx_offset,
                           uint32_t* counts, uint64_t* stats, cudaStream_t s) {
  if (!mean || !var || !counts || n_c < 0 || K < 1 || ld < K || num_samples < 0) return CRS_ERR_INVALID_ARG;
  if (num_samples > 0xffffffffLL || K < 2 || K - 1 - kTreeDirect > kTreeMaxLeaves) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0) return CRS_OK;
  CRS_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)n_c, s), "pcs_tree memset");
  if (stats) CRS_CUDA_TRY(cudaMemsetAsync(stats, 0, 3 * sizeof(uint64_t), s), "pcs_tree memset stats");
  if (num_samples == 0) return CRS_OK;
  int D = 1;
  while ((1 << (2 * D)) < K - 1 - kTreeDirect) ++D;
  const int NL = 1 << (2 * D), NI = (NL - 1) / 3;
  int n2 = 1;
  while (n2 < K - 1) n2 <<= 1;
  const int slots = D >= 2 ? 3 * (D - 1) + 1 : 1;
  const size_t walk = (size_t)slots * kTreeThreads * 12 + (size_t)kTreeWarps * kTreeQueue * 20;
  size_t build = (size_t)n2 * 16;  // sort scratch + competitor table, then the [NI][5] envelope tables — dead before the walk
  if (build < (size_t)NI * 20) build = (size_t)NI * 20;
  const size_t smem = (size_t)NL * 8 + (size_t)NI * 32 + (walk > build ? walk : build);
  // per-device facts, looked up once under a lock (the library may be entered from several host threads)
  struct DevInfo { unsigned long long* ring = nullptr; int sms = 0; int resident[2][8] = {}; };
  static DevInfo info[64];
  static std::mutex info_mutex;
  static std::atomic<unsigned int> next_slot{0};  // launches from several host threads take distinct tickets
  int dev = 0;
  CRS_CUDA_TRY(cudaGetDevice(&dev), "pcs_tree get device");
  if (dev < 0 || dev >= 64) return CRS_ERR_UNSUPPORTED;
  unsigned long long* ring;
  {
    std::lock_guard<std::mutex> lock(info_mutex);
    if (!info[dev].ring) {
      CRS_CUDA_TRY(cudaGetSymbolAddress((void**)&info[dev].ring, g_tree_tickets), "pcs_tree ticket symbol");
      CRS_CUDA_TRY(cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev), "pcs_tree sm count");
    }
    ring = info[dev].ring;
  }
  unsigned long long* ticket = ring + 2 * (next_slot.fetch_add(1u, std::memory_order_relaxed) % kTreeTicketRing);
  const RoundKeys rk = round_keys((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  auto go = [&](auto kern, auto* m, auto* v) -> int {
    // resident CTAs of this (device, dtype, depth): queried once, the answers do not change
    int resident_c;
    {
      std::lock_guard<std::mutex> lock(info_mutex);
      int& slot = info[dev].resident[dtype == CRS_F64 ? 0 : 1][D];
      if (slot == 0) {
        if (smem > 40 * 1024)
          CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "pcs_tree attr");
        int per_sm = 0;
        CRS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTreeThreads, smem), "pcs_tree occupancy");
        if (per_sm < 1) return CRS_ERR_UNSUPPORTED;
        slot = info[dev].sms * per_sm;
      }
      resident_c = slot;
    }
    const int64_t resident = resident_c;
    // ~12 work items per resident CTA so that the dynamic hand-out evens out their different costs;
    // a work item re-builds the plan of its context, so sample ranges are not split too finely
    const int64_t max_splits = (num_samples + kTreeSplitSamples - 1) / kTreeSplitSamples;
    int64_t splits = (12 * resident + n_c - 1) / n_c;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int64_t items = n_c * splits;
    const int blocks = (int)(items < resident ? items : resident);
    kern<<<blocks, kTreeThreads, smem, s>>>(m, v, n_c, K, ld, num_samples, rk, offset, ctx_offset,
                                            (int)splits, D, n2, slots, ticket, counts,
                                            (unsigned long long*)stats);
    CRS_CUDA_TRY(cudaGetLastError(), "pcs_tree launch");
    return CRS_OK;
  };
  if (dtype == CRS_F64) return go(pcs_tree_kernel<double>, (const double*)mean, (const double*)var);
  if (dtype == CRS_F32) return go(pcs_tree_kernel<float>, (const float*)mean, (const float*)var);
  return CRS_ERR_INVALID_ARG;
}

int launch_tree_probe(int kind, int iters, int blocks, void* out, cudaStream_t s) {
  tree_probe_kernel<<<blocks, kTreeThreads, 0, s>>>(kind, iters, round_keys(1u, 2u), (float*)out);
  CRS_CUDA_TRY(cudaGetLastError(), "tree_probe launch");
  return CRS_OK;
}

}  // namespace crs
