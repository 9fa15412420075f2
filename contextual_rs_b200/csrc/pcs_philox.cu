// crs_pcs_counts (K3): Monte-Carlo correct-selection counts per context — replaces the sampling
// half of estimate_current_generalized_pcs for a ModelListGP
// (contextual_rs/generalized_pcs.py:441-475): draw posterior samples at every (arm, context),
// order arms by posterior mean, delta = y_(1) - max_{j>=2} y_(j), func_I = (delta > 0), mean over
// samples.  The reference materialises an S x K x n_c sample tensor (550 GB at config 3); here
// nothing S-sized ever leaves the SM.
//
// Stream: Philox4x32-10 keyed (seed; sample, context, arm >> 2, offset) -> the draws of one sample
// at 4 consecutive arms (Box-Muller on 23-bit uniforms, |z| <= 5.7683).  Because a (sample,
// context) outcome does not depend on the order in which arms are visited, the kernel may stop a
// sample as soon as its outcome is decided — all of the following leave every count IDENTICAL to
// the brute-force estimator (tests: count-for-count against the CPU restatement):
//   screen   arm j with mu_j + ZB sd_j < mu_b - ZB sd_b (ZB = 5.77) can never win: dropped up front;
//   order    arm blocks are visited by decreasing threat  (mu_j - mu_b) / sqrt(sd_j^2 + sd_b^2);
//   fail     the first block in which some y_j >= y_b ends the sample (failure);
//   bound    if no remaining arm can reach y_b (suffix max of mu_j + ZB sd_j < y_b) the sample ends
//            (success);
//   pool     every lane owns one live sample and its own position in the visiting order; a lane
//            whose sample is decided takes the next unassigned sample in the same iteration, so
//            warps stay full although samples live for very different numbers of blocks (the
//            per-lane plan reads are shared-memory gathers of 40 bytes per block).
// Executed and skipped draws are reported in `stats`.
#include <atomic>
#include <mutex>
#include <math_constants.h>
#include "crs_common.cuh"
#include "pcs_rng.cuh"

namespace crs {
namespace {

using namespace rng;

constexpr int kPcsThreads = 256;
constexpr int kPcsTail = 8;                       // survivors per warp handed to the flattened tail
constexpr int kPcsSplitSamples = 4096;           // smallest sample range worth a work item of its own
constexpr float kZBound = 5.77f;                 // > sqrt(-2 ln 2^-24) = 5.7683 + MUFU slack
// one candidate arm block in visiting order (48 bytes: two LDS.128 + one LDS.64 per step)
struct __align__(16) Entry {
  float4 mu, sd;     // centred means / std-devs of the 4 arms (-inf / 0: not a competitor)
  float suf_next;    // max of mu + ZB sd over all LATER entries (-inf behind the last one)
  uint32_t id;       // global arm-block id (Philox counter word)
  float pad[2];
};

// shared-memory plan of one context: the candidate arm blocks in visiting order
struct Plan {
  float4* mu;      // [nb] centred means of the 4 arms of a block (-inf: not a competitor)
  float4* sd;      // [nb] std-devs (0 for non-competitors)
  float* ub;       // [nb] max over the block of mu + ZB sd
  uint32_t* id;    // [nb] global arm-block id (Philox counter word)
  float* key;      // [np2] sort key (threat), descending
  int* ord;        // [np2] visiting order -> entry of the arrays above
  float* suf;      // [nb] suffix max of ub in visiting order
  Entry* ent;      // [nb] the plan in visiting order: what the sampling loop gathers
};

#ifndef CRS_PCS_MIN_BLOCKS
#define CRS_PCS_MIN_BLOCKS 4
#endif
constexpr int kPcsTicketRing = 1024;
__device__ unsigned long long g_pcs_tickets[kPcsTicketRing];
template <typename T>
__global__ void __launch_bounds__(kPcsThreads, CRS_PCS_MIN_BLOCKS)
pcs_counts_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K,
                  int64_t ld, int64_t num_samples, const __grid_constant__ RoundKeys rk, uint32_t offset,
                  int blk_offset, int64_t ctx_offset, int splits, int nbmax, int np2,
                  unsigned long long* __restrict__ ticket,
                  uint32_t* counts, unsigned long long* stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ plan in visiting order | scratch of the plan build ]
  Plan pl;
  pl.ent = reinterpret_cast<Entry*>(smem_raw);
  pl.mu = reinterpret_cast<float4*>(pl.ent + nbmax);
  pl.sd = pl.mu + nbmax;
  pl.ub = reinterpret_cast<float*>(pl.sd + nbmax);
  pl.suf = pl.ub + nbmax;
  pl.id = reinterpret_cast<uint32_t*>(pl.suf + nbmax);
  pl.key = reinterpret_cast<float*>(pl.id + nbmax);
  pl.ord = reinterpret_cast<int*>(pl.key + np2);
  __shared__ T s_bm[kPcsThreads / 32];
  __shared__ int s_bi[kPcsThreads / 32];
  __shared__ int s_nb, s_best;
  __shared__ float s_sdb;
  __shared__ unsigned int s_hits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t items = n_c * splits;
  const int nblk_all = (K + 3) >> 2;
  unsigned long long drawn = 0;
  long long skipped = 0;

  // Work items (context x sample range) cost anything between 4 and 60 plan entries per sample, so
  // they are handed out dynamically: a CTA takes the next item from a global ticket when it is done.
  __shared__ long long s_item;
  for (;;) {
    if (tid == 0) s_item = (long long)atomicAdd(ticket, 1ull);
    __syncthreads();
    const int64_t item = s_item;
    if (item >= items) break;
    const int64_t c = item / splits;
    const int sp = (int)(item - c * splits);
    const int64_t s_lo = num_samples * sp / splits, s_hi = num_samples * (sp + 1) / splits;
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    __syncthreads();
    // ---- predicted best arm: first arg-max of the mean row (generalized_pcs.py:464) -------
    T bm = -CUDART_INF;
    int bi = 0x7fffffff;
    for (int a = tid; a < K; a += kPcsThreads) {
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
    if (tid == 0) { s_nb = 0; s_hits = 0; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kPcsThreads / 32; ++w)
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
    const float low_b = -kZBound * sd_b;  // lowest value the best arm's draw can take
    // ---- plan: screen, threat key, compaction of the candidate arm blocks -------------------
    for (int b0 = 0; b0 < nblk_all; b0 += kPcsThreads) {
      const int blk = b0 + tid;
      float mu4[4], sd4[4], key = -CUDART_INF_F, ub = -CUDART_INF_F;
      bool keep = false;
      if (blk < nblk_all) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int a = 4 * blk + u;
          mu4[u] = -CUDART_INF_F;
          sd4[u] = 0.f;
          if (a < K && a != best) {
            const float mu = (float)(mrow[a] - mu_b);  // centred in the grid dtype, then fp32
            const float sd = (float)sqrt((double)vrow[a]);
            const float top = fmaf(sd, kZBound, mu);
            if (!(top < low_b)) {  // exact screen: can it ever reach the best arm's lowest draw?
              mu4[u] = mu;
              sd4[u] = sd;
              ub = fmaxf(ub, top);
              key = fmaxf(key, mu * rsqrtf(fmaf(sd, sd, sd_b * sd_b) + 1e-30f));
              keep = true;
            }
          }
        }
        if (blk == (best >> 2)) { keep = true; key = CUDART_INF_F; }  // the best arm's block goes first
      }
      const unsigned ball = __ballot_sync(kFull, keep);
      int base = 0;
      if (lane == 0 && ball) base = atomicAdd(&s_nb, __popc(ball));
      base = __shfl_sync(kFull, base, 0);
      if (keep) {
        const int pos = base + __popc(ball & ((1u << lane) - 1u));
        pl.mu[pos] = make_float4(mu4[0], mu4[1], mu4[2], mu4[3]);
        pl.sd[pos] = make_float4(sd4[0], sd4[1], sd4[2], sd4[3]);
        pl.ub[pos] = ub;
        pl.id[pos] = (uint32_t)(blk + blk_offset);
        pl.key[pos] = key;
        pl.ord[pos] = pos;
      }
    }
    __syncthreads();
    const int nb = s_nb;
    int n2 = 1;
    while (n2 < nb) n2 <<= 1;
    for (int i = nb + tid; i < n2; i += kPcsThreads) { pl.key[i] = -CUDART_INF_F; pl.ord[i] = -1; }
    __syncthreads();
    // bitonic sort of (key, ord), descending key (ties: lower entry first keeps it deterministic)
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n2; i += kPcsThreads) {
          const int p = i ^ j;
          if (p > i) {
            const float ka = pl.key[i], kb = pl.key[p];
            const int oa = pl.ord[i], ob = pl.ord[p];
            const bool a_first = ka > kb || (ka == kb && (unsigned)oa < (unsigned)ob);
            const bool desc = (i & k) == 0;
            if (desc != a_first) {
              pl.key[i] = kb; pl.key[p] = ka;
              pl.ord[i] = ob; pl.ord[p] = oa;
            }
          }
        }
        __syncthreads();
      }
    }
    // suffix max of the upper bounds in visiting order
    for (int e = tid; e < nb; e += kPcsThreads) {
      const int idx = pl.ord[e];
      pl.suf[e] = pl.ub[idx];
      pl.ent[e].mu = pl.mu[idx];
      pl.ent[e].sd = pl.sd[idx];
      pl.ent[e].id = pl.id[idx];
    }
    __syncthreads();
    for (int off = 1; off < nb; off <<= 1) {
      float vals[(8192 / 4 + kPcsThreads - 1) / kPcsThreads];
      int cnt = 0;
      for (int e = tid; e < nb; e += kPcsThreads, ++cnt)
        vals[cnt] = e + off < nb ? fmaxf(pl.suf[e], pl.suf[e + off]) : pl.suf[e];
      __syncthreads();
      cnt = 0;
      for (int e = tid; e < nb; e += kPcsThreads, ++cnt) pl.suf[e] = vals[cnt];
      __syncthreads();
    }
    for (int e = tid; e < nb; e += kPcsThreads) pl.ent[e].suf_next = e + 1 < nb ? pl.suf[e + 1] : -CUDART_INF_F;
    __syncthreads();
    const uint32_t cw = (uint32_t)(c + ctx_offset);
    const int bslot = best & 3;

    // ---- sampling: a pool of 32 live samples per warp ---------------------------------------
    // The warp owns a contiguous range of the item's samples; `wnext` is warp-uniform.  The step
    // is branch-free: lanes without a sample (only at the very end) compute on entry 0 and are
    // masked out of the tallies.
    const unsigned int n_item = (unsigned int)(s_hi - s_lo);
    const unsigned int w_hi = (unsigned int)((uint64_t)n_item * (warp + 1) / (kPcsThreads / 32));
    unsigned int wnext = (unsigned int)((uint64_t)n_item * warp / (kPcsThreads / 32));
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int hits = 0, steps = 0, wsteps = 0;  // wsteps: warp-uniform count of pooled lane-steps
    uint32_t smp = 0;
    float yb = 0.f;
    int e = 0;
    bool active = false;
    // one visit of plan entry `en` by sample `sm`: the 4 draws, max over the competitors of the block
    auto visit = [&](uint32_t sm, const Entry* en, float& suf_next, float& zb) -> float {
      const float4 mu = en->mu, sd = en->sd;
      const float2 sv = *reinterpret_cast<const float2*>(&en->suf_next);
      const Words w = philox4x32_10(sm, cw, __float_as_uint(sv.y), offset, rk);
      float z0, z1, z2, z3;
      box_muller(w.x, w.y, z0, z1);
      box_muller(w.z, w.w, z2, z3);
      zb = (bslot & 2) ? ((bslot & 1) ? z3 : z2) : ((bslot & 1) ? z1 : z0);
      suf_next = sv.x;
      return fmaxf(fmaxf(fmaf(sd.x, z0, mu.x), fmaf(sd.y, z1, mu.y)),
                   fmaxf(fmaf(sd.z, z2, mu.z), fmaf(sd.w, z3, mu.w)));
    };
    unsigned act;
    for (;;) {
      act = __ballot_sync(kFull, active);
      if (act != kFull && wnext < w_hi) {  // hand the next unassigned samples to the idle lanes
        const unsigned int my = wnext + __popc(~act & lt_mask);
        if (!active && my < w_hi) {
          smp = (uint32_t)(s_lo + my);
          active = true;
        }
        wsteps += min((unsigned int)__popc(~act), w_hi - wnext);
        wnext += __popc(~act);
      } else if (__popc(act) <= kPcsTail) {
        // out of fresh samples and few survivors left: finish them in the flattened tail below —
        // unless one of them has not drawn its y_b yet (entry 0): then one more pooled step
        if (!__any_sync(kFull, active && e == 0)) break;
      }
      wsteps += __popc(act);  // lanes stepping in this iteration (+ the refilled ones counted above)
      float suf_next, zb;
      const float top = visit(smp, pl.ent + e, suf_next, zb);
      yb = e == 0 ? sd_b * zb : yb;
      const bool fail = top >= yb;       // delta = y_b - max_j y_j <= 0: func_I = 0
      const bool done = suf_next < yb;   // nothing left can reach y_b (or nothing left at all)
      hits += (active && !fail && done) ? 1u : 0u;
      active = active && !(fail || done);
      e = active ? e + 1 : 0;
    }
    // flattened tail: the few long-lived samples left each spread their remaining entries over the
    // 32 lanes.  Entries behind the point where the sequential visit would have stopped cannot
    // fail (their mu + ZB sd is below y_b), so the outcome — and every count — is unchanged.
    while (act) {
      const int src = __ffs(act) - 1;
      act &= act - 1;
      const uint32_t t_smp = __shfl_sync(kFull, smp, src);
      const float t_yb = __shfl_sync(kFull, yb, src);
      for (int base = __shfl_sync(kFull, e, src); base < nb; base += 32) {
        const int ee = base + lane;
        const bool valid = ee < nb;
        float suf_next, zb;
        const float top = visit(t_smp, pl.ent + (valid ? ee : nb - 1), suf_next, zb);
        steps += valid ? 1u : 0u;
        const unsigned fm = __ballot_sync(kFull, valid && top >= t_yb);
        const unsigned dm = __ballot_sync(kFull, valid && suf_next < t_yb);
        if (fm) break;
        if (dm) { hits += lane == 0 ? 1u : 0u; break; }
      }
    }
    drawn += 4ull * steps + (lane == 0 ? 4ull * wsteps : 0ull);
    hits = warp_sum(hits);
    if (lane == 0 && hits) atomicAdd(&s_hits, hits);
    __syncthreads();
    if (tid == 0 && s_hits) atomicAdd(counts + c, s_hits);
    if (tid == 0) skipped += (long long)(s_hi - s_lo) * K;  // nominal draws of this item
  }
  if (stats) {
    drawn = warp_sum(drawn);
    skipped = warp_sum(skipped);
    if (lane == 0) {
      atomicAdd(stats, drawn);
      atomicAdd(stats + 1, (unsigned long long)(skipped - (long long)drawn));  // nominal - executed
    }
  }
}

// Peak probe for the PCS roofline (crs_peak_probe kind 5): the bare draw sequence of one plan
// step — Philox4x32-10, two Box-Muller pairs, 4 FMA, max, compare — on register operands, no
// memory, no control flow besides the loop.  draws = blocks * 256 * iters * 4.
__global__ void __launch_bounds__(kPcsThreads)
pcs_probe_kernel(int iters, const __grid_constant__ RoundKeys rk, float* out) {
  const uint32_t smp = blockIdx.x * kPcsThreads + threadIdx.x;
  const float4 mu = make_float4(-0.1f, -0.2f, -0.3f, -0.4f), sd = make_float4(1.f, 1.1f, 1.2f, 1.3f);
  unsigned int hits = 0;
  for (int it = 0; it < iters; ++it) {
    const Words w = philox4x32_10(smp, 7u, (uint32_t)it, 0u, rk);
    float z0, z1, z2, z3;
    box_muller(w.x, w.y, z0, z1);
    box_muller(w.z, w.w, z2, z3);
    const float top = fmaxf(fmaxf(fmaf(sd.x, z0, mu.x), fmaf(sd.y, z1, mu.y)),
                            fmaxf(fmaf(sd.z, z2, mu.z), fmaf(sd.w, z3, mu.w)));
    hits += top >= 3.0f ? 1u : 0u;
  }
  if (hits == 0xffffffffu) out[0] = 1.f;  // keep the loop alive
}

__global__ void philox_words_kernel(const uint32_t* __restrict__ ctr, int64_t n, uint32_t k0,
                                    uint32_t k1, uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Words w = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], round_keys(k0, k1));
  out[4 * i] = w.x;
  out[4 * i + 1] = w.y;
  out[4 * i + 2] = w.z;
  out[4 * i + 3] = w.w;
}

}  // namespace

int launch_pcs_counts(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                      int64_t num_samples, uint64_t seed, uint32_t offset, int arm_offset,
                      int64_t ctx_offset, uint32_t* counts, uint64_t* stats, cudaStream_t s) {
  if (!mean || !var || !counts || n_c < 0 || K < 1 || ld < K || num_samples < 0) return CRS_ERR_INVALID_ARG;
  if (arm_offset & 3) return CRS_ERR_INVALID_ARG;  // the stream is keyed by blocks of 4 arms
  if (num_samples > 0xffffffffLL || K > 8192) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0) return CRS_OK;
  CRS_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)n_c, s), "pcs memset");
  if (stats) CRS_CUDA_TRY(cudaMemsetAsync(stats, 0, 2 * sizeof(uint64_t), s), "pcs memset stats");
  if (num_samples == 0) return CRS_OK;
  const int nbmax = ((K + 3) / 4 + 3) & ~3;  // array stride: keeps every float4 array 16-byte aligned
  int np2 = 1;
  while (np2 < nbmax) np2 <<= 1;
  const size_t smem = (size_t)nbmax * (sizeof(Entry) + 2 * 16 + 3 * 4) + (size_t)np2 * 8;
  // enough work items to fill the machine: split the samples of a context when contexts are few
  const int64_t max_splits = (num_samples + kPcsSplitSamples - 1) / kPcsSplitSamples;
  // ~16 items per resident CTA, so that the dynamic hand-out can even out their very different costs
  int64_t splits = (16 * 148 * CRS_PCS_MIN_BLOCKS + n_c - 1) / n_c;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int64_t items = n_c * splits;
  const int blocks = (int)(items < 148 * CRS_PCS_MIN_BLOCKS * 2 ? items : 148 * CRS_PCS_MIN_BLOCKS * 2);
  // ticket of this launch: a small ring of device counters, so that launches in flight on different
  // streams do not share one
  static unsigned long long* rings[64] = {nullptr};  // per device: the symbol lives in each context
  static std::mutex rings_mutex;                     // the library may be entered from several host threads
  static std::atomic<unsigned int> next_slot{0};  // launches from several host threads take distinct tickets
  int dev = 0;
  CRS_CUDA_TRY(cudaGetDevice(&dev), "pcs get device");
  if (dev < 0 || dev >= 64) return CRS_ERR_UNSUPPORTED;
  unsigned long long* ring;
  {
    std::lock_guard<std::mutex> lock(rings_mutex);
    if (!rings[dev]) CRS_CUDA_TRY(cudaGetSymbolAddress((void**)&rings[dev], g_pcs_tickets), "pcs ticket symbol");
    ring = rings[dev];
  }
  unsigned long long* ticket = ring + (next_slot.fetch_add(1u, std::memory_order_relaxed) % kPcsTicketRing);
  CRS_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), s), "pcs memset ticket");
  const RoundKeys rk = round_keys((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  auto go = [&](auto kern, auto* m, auto* v) -> int {
    if (smem > 40 * 1024)
      CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pcs attr");
    kern<<<blocks, kPcsThreads, smem, s>>>(m, v, n_c, K, ld, num_samples, rk, offset, arm_offset >> 2,
                                           ctx_offset, (int)splits, nbmax, np2, ticket, counts,
                                           (unsigned long long*)stats);
    CRS_CUDA_TRY(cudaGetLastError(), "pcs_counts launch");
    return CRS_OK;
  };
  if (dtype == CRS_F64) return go(pcs_counts_kernel<double>, (const double*)mean, (const double*)var);
  if (dtype == CRS_F32) return go(pcs_counts_kernel<float>, (const float*)mean, (const float*)var);
  return CRS_ERR_INVALID_ARG;
}

int launch_pcs_probe(int iters, int blocks, void* out, cudaStream_t s) {
  pcs_probe_kernel<<<blocks, kPcsThreads, 0, s>>>(iters, round_keys(1u, 2u), (float*)out);
  CRS_CUDA_TRY(cudaGetLastError(), "pcs_probe launch");
  return CRS_OK;
}

int launch_philox_words(const uint32_t* ctr, int64_t n, uint32_t k0, uint32_t k1, uint32_t* out,
                        cudaStream_t s) {
  if (!ctr || !out || n < 0) return CRS_ERR_INVALID_ARG;
  if (n == 0) return CRS_OK;
  philox_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ctr, n, k0, k1, out);
  CRS_CUDA_TRY(cudaGetLastError(), "philox_words launch");
  return CRS_OK;
}

}  // namespace crs
