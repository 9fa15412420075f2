// crs_pcs_counts_cdf (K3, "cdf" stream, default): Monte-Carlo correct-selection counts per context —
// the sampling half of estimate_current_generalized_pcs for a ModelListGP
// (contextual_rs/generalized_pcs.py:441-475): in how many of the S posterior draws does the
// predicted-best arm b beat every other arm (func_I = (delta > 0), experiments/wsc_experiments/main.py:453)?
//
// Arms are independent, so given the best arm's draw y_b = mu_b + sd_b z the other K - 1 draws are
// still independent normals and
//     P(success | z) = F_c(z) = prod_{j != b} Phi(a_j z + b_j),   a_j = sd_b / sd_j,  b_j = (mu_b - mu_j) / sd_j.
// One sample of the indicator is therefore: z ~ N(0,1), u ~ U(0,1), success iff u < F_c(z) — the SAME
// Bernoulli law as drawing all K normals and comparing (counts[c] ~ Binomial(S, PCS(c)) either way), at
// the cost of one normal and one uniform instead of K normals.  Two kernels:
//   build   one warp per context: arg-max, the range [lo, hi] of z outside which F is 0 / 1 to 1e-9
//           (lo = max_j (-6 - b_j) / a_j, hi = max_j (7 - b_j) / a_j, clipped to the Box-Muller range), narrowed
//           to the LIVE range (two 8-point evaluations of ln F over all competitors bracket ln F = -23 and
//           ln F = -2e-8: with many competitors the product is dead on most of [lo, hi]); ln F at 32 R grid
//           points, R = 1..4 (ln Phi from a 1,024-cell cubic table kept in EIGHT bank-interleaved copies in
//           shared memory — 128-bit look-ups without bank conflicts for any access pattern; 14 instructions per
//           term; competitors grouped by the number of leading quarters in which their factor is not yet 1),
//           then the 4-point Lagrange cubic of F on each of the G = 32 R - 3 cells as one float4 -> workspace,
//           with a guard slot below (F = 0) and above (F = 1).  Competitors whose factor is too sharp for the
//           grid (a_j h > 1: the best arm's posterior is much wider than theirs) are kept OUT of the table,
//           listed in the context's header and evaluated exactly per sample (up to 8 per context; beyond that
//           the context is flagged in stats).
//   sample  one warp per (context, sample range): the 2 KB table goes to shared memory; one
//           Philox4x32-10 call per TWO samples — Box-Muller pair (z1, z2) from words 0, 1, the uniforms
//           are words 2, 3 compared as integers with F * 2^32 — ~24 instructions per sample,
//           independent of K, with no divergence and no queues; four calls in flight per lane.
// Philox counters (call, global context, 0xCDF00001, offset): counts do not depend on how contexts are
// sharded.  Stream definition and CPU restatement: oracle/pcs_cdf_ref.py.  Interpolation bias of F,
// integrated against the normal density: < 3e-7 on all BASELINE configurations (tests/tools/cdf_design.py),
// three orders of magnitude below the Monte-Carlo standard error of a 2^16-sample estimate.
#include <math_constants.h>
#include <type_traits>
#include <atomic>
#include <climits>
#include <mutex>
#include "crs_common.cuh"
#include "pcs_rng.cuh"

namespace crs {
namespace {

using namespace rng;

constexpr int kCdfThreads = 256;
constexpr int kCdfWarps = kCdfThreads / 32;
constexpr int kCdfCells = 125;          // most cells an F table has (grid points -1 .. 126: 4 per lane)
constexpr int kCdfMinCells = 29;        // fewest: one quarter of the grid
constexpr float kCdfH0 = 11.6f / 125.f; // design spacing: a full-range table, z in [-5.8, 5.8]
constexpr float kCdfLnDead = -23.f;     // ln F below this (F < 1e-10): F = 0 below the live range
constexpr float kCdfLnFull = -2e-8f;    // ln F above this: F = 1 above the live range
constexpr int kCdfStride = 128;         // float4 per context in the workspace
constexpr int kCdfHdr = 24;             // floats per context header: lo, inv_h, n_sharp, flags, pad[4], sharp[8] (a, b)
constexpr int kCdfMaxSharp = 8;
constexpr float kCdfZMax = 5.8f;        // Box-Muller on 24-bit uniforms: |z| <= 5.77
constexpr float kCdfXLo = 6.0f;         // Phi(-6) = 1e-9: F = 0 below lo
constexpr float kCdfXHi = 7.0f;         // K * (1 - Phi(7)) < 1e-8 for K < 8,000: F = 1 above hi
constexpr float kCdfSharp = 1.0f;       // a_j h above this: the factor is evaluated exactly per sample
constexpr uint32_t kCdfTag = 0xCDF00001u;

// ---- ln Phi(x) on [-9, 7]: 1,024 cells of width 1/64, 4-point Lagrange cubic per cell ----------
constexpr int kLnPhiCells = 1024;
constexpr float kLnPhiX0 = -9.0f, kLnPhiScale = 64.0f;
__device__ float4 g_lnphi[kLnPhiCells];

__global__ void lnphi_init_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kLnPhiCells) return;
  double p[4];
  for (int k = 0; k < 4; ++k) {
    const double x = (double)kLnPhiX0 + (double)(i - 1 + k) / (double)kLnPhiScale;
    // ln Phi(x): log of the lower tail for x < 0, log1p of minus the upper tail for x >= 0
    p[k] = x < 0.0 ? log(0.5 * erfc(-x * 0.70710678118654752440)) : log1p(-0.5 * erfc(x * 0.70710678118654752440));
  }
  const double c1 = -p[0] / 3.0 - p[1] / 2.0 + p[2] - p[3] / 6.0;
  const double c2 = p[0] / 2.0 - p[1] + p[2] / 2.0;
  const double c3 = -p[0] / 6.0 + p[1] / 2.0 - p[2] / 2.0 + p[3] / 6.0;
  g_lnphi[i] = make_float4((float)p[1], (float)c1, (float)c2, (float)c3);
}

__device__ __forceinline__ void stage_lnphi(float4* __restrict__ s_tab, int tid, int nthreads) {
  for (int i = tid; i < kLnPhiCells; i += nthreads) s_tab[i] = g_lnphi[i];
}

// floor of 0 <= t < 2^22 without the conversion unit: t + 2^23 rounded toward zero keeps floor(t) in the
// low mantissa bits (two FADDs and one LOP3 on the FMA / ALU pipes instead of F2I + I2F on the XU pipe)
__device__ __forceinline__ int floor_split(float t, float& frac) {
  const float u = __fadd_rz(t, 8388608.f);
  frac = t - (u - 8388608.f);
  return __float_as_int(u) & 0x3fffff;
}

__device__ __forceinline__ float lnphi(float x, const float4* __restrict__ tab) {
  float t = fmaf(x, kLnPhiScale, -kLnPhiX0 * kLnPhiScale);
  t = fminf(fmaxf(t, 0.f), (float)kLnPhiCells - 0.001f);  // x <= -9: ln Phi = -43.6 (F = 0 anyway); x >= 7: -1.3e-12
  float f;
  const int i = floor_split(t, f);
  const float4 c = tab[i];
  return fmaf(fmaf(fmaf(c.w, f, c.z), f, c.y), f, c.x);
}

__device__ __forceinline__ float exp_approx(float x) {  // e^x, x <= 0
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

template <typename T> __device__ __forceinline__ float sd_of(T v);
template <> __device__ __forceinline__ float sd_of<float>(float v) { return __fsqrt_rn(v); }
template <> __device__ __forceinline__ float sd_of<double>(double v) { return (float)sqrt(v); }

// (a_j, b_j) of competitor `arm` against the best arm (mu_b in the grid dtype, sd_b fp32); the means
// are centred in the grid dtype and only then rounded to fp32, like the other PCS kernels do.
template <typename T>
__device__ __forceinline__ void competitor(const T* __restrict__ mrow, const T* __restrict__ vrow, int arm,
                                           T mu_b, float sd_b, float& a, float& b, float& r) {
  const float mu = (float)(mrow[arm] - mu_b);
  const float sd = sd_of<T>(vrow[arm]);
  r = __fdiv_rn(sd, sd_b);          // 1 / a_j
  a = __fdiv_rn(sd_b, sd);
  b = __fdiv_rn(-mu, sd);
}

// ---- build ---------------------------------------------------------------------------------------
// The build is a stream of table terms ln Phi(a_j z_g + b_j): per term one 16-byte look-up.  A 128-bit
// shared-memory load is served one quarter-warp (8 lanes) per wavefront, and 8 lanes reading cells an
// irregular stride apart collide in the eight 16-byte bank groups (measured: 9 wavefronts per request,
// the LSU pipe 96 % busy — the kernel's bound in round 2a).  So the CTA keeps EIGHT copies of the table,
// copy k occupying bank group k only (cell i of copy k at float4 index 8 i + k), and lane l reads copy
// l & 7: the 8 lanes of a quarter-warp never share a bank group, whatever cells they want — 4 wavefronts
// per request, the minimum for 16 bytes per lane.  128 KB of shared memory: one 1,024-thread CTA per SM.
constexpr int kBuildThreads = 1024;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kBuildSmem = kLnPhiCells * 8 * (int)sizeof(float4);

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// tab8_k16 = shared-window address of the interleaved table + 16 (lane & 7).  t + 2^16 rounded toward zero
// keeps floor(128 t) in the mantissa: bits 7.. are the cell, i.e. exactly the byte offset 128 floor(t).
__device__ __forceinline__ float lnphi8(float t, uint32_t tab8_k16) {
  // t = (x + 9) * 64; x >= -7 for every tabulated factor on [lo - h, ...) except the sharp competitors beyond
  // the eighth of a flagged context, x >= 7 inside a quarter that started below it: clamp both ends
  t = fminf(fmaxf(t, 0.f), (float)kLnPhiCells - 0.001f);
  const int u = __float_as_int(__fadd_rz(t, 65536.f));
  const float f = t - (__int_as_float(u & ~0x7f) - 65536.f);
  const float4 c = lds128(tab8_k16 + (uint32_t)(u & 0x1ff80));
  return fmaf(fmaf(fmaf(c.w, f, c.z), f, c.y), f, c.x);
}

template <typename T>
__global__ void __launch_bounds__(kBuildThreads, 1)
pcs_cdf_build_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K, int64_t ld,
                     float4* __restrict__ tables, float* __restrict__ headers, uint32_t* __restrict__ counts,
                     unsigned long long* __restrict__ ticket, unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  float4* s_tab8 = reinterpret_cast<float4*>(dyn_smem);
  __shared__ float s_F[kBuildWarps][kCdfStride];
  __shared__ float2 s_sharp[kBuildWarps][kCdfMaxSharp];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kLnPhiCells * 8; i += kBuildThreads) s_tab8[i] = g_lnphi[i >> 3];
  __syncthreads();
  const uint32_t copy = smem_u32(s_tab8) + 16u * (uint32_t)(lane & 7);
  const unsigned lt_mask = (1u << lane) - 1u;
  constexpr float kTScale = kLnPhiScale, kTShift = -kLnPhiX0 * kLnPhiScale, kTHi = (kCdfXHi - kLnPhiX0) * kLnPhiScale;
  unsigned long long n_terms = 0, n_sharp_ctx = 0, n_over_ctx = 0;
  for (;;) {
    // contexts are handed out one at a time (their cost depends on how many terms can be skipped)
    long long cc = 0;
    if (lane == 0) cc = (long long)atomicAdd(ticket, 1ull);
    const int64_t c = __shfl_sync(kFull, cc, 0);
    if (c >= n_c) break;
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    // predicted best arm: first arg-max of the mean row (generalized_pcs.py:464)
    T bm = -(T)CUDART_INF;
    int bi = INT_MAX;
    for (int a = lane; a < K; a += 32) {
      const T m = mrow[a];
      if (m > bm) { bm = m; bi = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T om = __shfl_xor_sync(kFull, bm, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
    }
    if (bi == INT_MAX) bi = 0;
    const T mu_b = mrow[bi];
    const float sd_b = sd_of<T>(vrow[bi]);
    // range of z outside which F is 0 / 1
    float zlo = -CUDART_INF_F, zhi = -CUDART_INF_F;
    for (int a = lane; a < K; a += 32) {
      if (a == bi) continue;
      float aj, bj, rj;
      competitor<T>(mrow, vrow, a, mu_b, sd_b, aj, bj, rj);
      zlo = fmaxf(zlo, __fmul_rn(-kCdfXLo - bj, rj));
      zhi = fmaxf(zhi, __fmul_rn(kCdfXHi - bj, rj));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      zlo = fmaxf(zlo, __shfl_xor_sync(kFull, zlo, o));
      zhi = fmaxf(zhi, __shfl_xor_sync(kFull, zhi, o));
    }
    float lo = fmaxf(zlo, -kCdfZMax);
    float hi = fminf(zhi, kCdfZMax);
    if (!(hi > lo + 1e-3f)) hi = lo + 1e-3f;  // F is 0 (or 1) over the whole sampled range
    unsigned int n_search = 0;
    // ---- live range.  [lo, hi] is where ONE factor leaves 0 / the last one reaches 1; with many competitors
    // the product F is below 1e-10 on most of it.  ln F is non-decreasing in z, so the live part [lo', hi']
    // (ln F > -23 ... ln F < -2e-8) is bracketed by two 8-point evaluations of ln F over ALL competitors
    // (lanes over competitors here, approximate (a, b): only the bracket depends on them) and only that
    // part is tabulated: K x 16 look-ups buy back up to three quarters of the K x 128 of the table.
    if (hi - lo > (float)kCdfMinCells * kCdfH0) {
      float zp[8], lf[8];
      auto eval8 = [&]() {
#pragma unroll
        for (int p = 0; p < 8; ++p) lf[p] = 0.f;
        for (int a = lane; a < K; a += 32) {
          if (a == bi) continue;
          const float rs = rsqrtf((float)vrow[a]);
          const float at = sd_b * rs * kTScale;
          const float bt = fmaf((float)(mu_b - mrow[a]) * rs, kTScale, kTShift);
#pragma unroll
          for (int p = 0; p < 8; ++p) lf[p] += lnphi8(fmaf(at, zp[p], bt), copy);
          n_search += 8;
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) lf[p] = warp_sum(lf[p]);
      };
      const float d1 = (hi - lo) * (1.f / 9.f);
#pragma unroll
      for (int p = 0; p < 8; ++p) zp[p] = fmaf((float)(p + 1), d1, lo);
      eval8();
      int nlo = 0, nhi = 0;  // leading points still dead / leading points not yet saturated
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        nlo += (lf[p] < kCdfLnDead && nlo == p) ? 1 : 0;
        nhi += (lf[p] < kCdfLnFull && nhi == p) ? 1 : 0;
      }
      const float l0 = fmaf((float)nlo, d1, lo);                     // dead at l0 (or l0 = lo), alive at l0 + d1
      const float u1 = nhi == 8 ? hi : fmaf((float)(nhi + 1), d1, lo);  // saturated at u1 (or u1 = hi)
      const float u0 = u1 - d1 < lo ? lo : (nhi == 8 ? fmaf(8.f, d1, lo) : u1 - d1);
      const float d2l = d1 * 0.2f, d2u = (u1 - u0) * 0.2f;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        zp[p] = fmaf((float)(p + 1), d2l, l0);
        zp[4 + p] = fmaf((float)(p + 1), d2u, u0);
      }
      eval8();
      int mlo = 0, mhi = 0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        mlo += (lf[p] < kCdfLnDead && mlo == p) ? 1 : 0;
        mhi += (lf[4 + p] < kCdfLnFull && mhi == p) ? 1 : 0;
      }
      const float lo2 = fmaf((float)mlo, d2l, l0);
      const float hi2 = mhi == 4 ? u1 : fmaf((float)(mhi + 1), d2u, u0);
      lo = fmaxf(lo, lo2);
      hi = fminf(hi, fmaxf(hi2, lo + 1e-3f));
    }
    n_terms += n_search;
    // cells: as many as the live range needs at the design spacing h0 (the spacing of a full-range table),
    // in whole quarters of 32 grid points: 29 / 61 / 93 / 125
    int R = ((int)ceilf((hi - lo) * (1.f / kCdfH0)) + 3 + 31) >> 5;
    R = R < 1 ? 1 : (R > 4 ? 4 : R);
    const int cells = 32 * R - 3;
    const float h = __fdiv_rn(hi - lo, (float)cells);
    const float inv_h = __fdiv_rn((float)cells, hi - lo);
    // grid points g = 32 r + lane <-> z = lo + (g - 1) h
    float z[4], acc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      z[r] = fmaf((float)(32 * r + lane - 1), h, lo);
      acc[r] = 0.f;
    }
    float zq[4];  // lowest z of each quarter: terms that are 0 there (x >= 7) are 0 in the whole quarter and above
#pragma unroll
    for (int r = 0; r < 4; ++r) zq[r] = r < R ? fmaf((float)(32 * r - 1), h, lo) : CUDART_INF_F;  // quarters beyond R: never live
    int n_sharp = 0;
    bool overflow = false;
    unsigned int n_quarters = 0;  // warp-uniform count of (competitor, quarter) blocks evaluated
    for (int base = 0; base < K; base += 32) {
      const int a = base + lane;
      float aj = 0.f, bj = 1e30f, rj = 0.f;
      if (a < K && a != bi) competitor<T>(mrow, vrow, a, mu_b, sd_b, aj, bj, rj);
      // too sharp for the grid and alive somewhere in [lo, hi]: exact per sample, not tabulated
      const bool sharp = aj * h > kCdfSharp && fmaf(aj, lo, bj) < kCdfXHi;
      const unsigned sb = __ballot_sync(kFull, sharp);
      if (sb) {
        const int pos = n_sharp + __popc(sb & lt_mask);
        if (sharp && pos < kCdfMaxSharp) {
          s_sharp[warp][pos] = make_float2(aj, bj);
          bj = 1e30f;  // out of the table
          aj = 0.f;
        }
        overflow = overflow || n_sharp + __popc(sb) > kCdfMaxSharp;
        n_sharp = min(n_sharp + __popc(sb), kCdfMaxSharp);
      }
      // table coordinates t = (a_j z + b_j + 9) * 64 = at z + bt, and the number of leading quarters in
      // which the factor is not yet 1 (a_j > 0: monotone in the quarter index), computed once per lane
      const float at = aj * kTScale;
      const float bt = fminf(fmaf(bj, kTScale, kTShift), 1e30f);
      int nq = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) nq += fmaf(at, zq[r], bt) < kTHi ? 1 : 0;
      // one loop per class "live in the first q quarters": the body is straight-line code with exactly q
      // look-ups (a per-competitor switch on q costs ~15 instructions of branch bookkeeping: the compiler
      // cannot see that q is warp-uniform)
      auto run_class = [&](auto q_tag, unsigned mask) {
        constexpr int Q = decltype(q_tag)::value;
        for (unsigned m = mask; m; m &= m - 1) {
          const int i = __ffs(m) - 1;
          const float ai = __shfl_sync(kFull, at, i);
          const float bb = __shfl_sync(kFull, bt, i);
          if constexpr (Q >= 4) acc[3] += lnphi8(fmaf(ai, z[3], bb), copy);
          if constexpr (Q >= 3) acc[2] += lnphi8(fmaf(ai, z[2], bb), copy);
          if constexpr (Q >= 2) acc[1] += lnphi8(fmaf(ai, z[1], bb), copy);
          acc[0] += lnphi8(fmaf(ai, z[0], bb), copy);
        }
      };
      const unsigned m4 = __ballot_sync(kFull, nq == 4), m3 = __ballot_sync(kFull, nq == 3);
      const unsigned m2 = __ballot_sync(kFull, nq == 2), m1 = __ballot_sync(kFull, nq == 1);
      n_quarters += 4 * __popc(m4) + 3 * __popc(m3) + 2 * __popc(m2) + __popc(m1);
      if (m4) run_class(std::integral_constant<int, 4>{}, m4);
      if (m3) run_class(std::integral_constant<int, 3>{}, m3);
      if (m2) run_class(std::integral_constant<int, 2>{}, m2);
      if (m1) run_class(std::integral_constant<int, 1>{}, m1);
    }
    n_terms += n_quarters;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r) s_F[warp][32 * r + lane] = exp_approx(fmaxf(acc[r], -80.f));
    __syncwarp();
    // slot 0 = "below lo" (F = 0), slots 1 .. 125 = the cells, slot 126 = "above hi" (F = 1): the sampler
    // clamps its table coordinate t' = (z - lo) / h + 1 to [0, 127) and needs no range selects
    float4* tab = tables + c * kCdfStride;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int slot = 32 * r + lane, i = slot - 1;
      float4 cf = make_float4(slot == cells + 1 ? 1.f : 0.f, 0.f, 0.f, 0.f);
      if (i >= 0 && i < cells) {
        const float p0 = s_F[warp][i], p1 = s_F[warp][i + 1], p2 = s_F[warp][i + 2], p3 = s_F[warp][i + 3];
        cf.x = p1;
        cf.y = fmaf(-1.f / 3.f, p0, fmaf(-0.5f, p1, fmaf(-1.f / 6.f, p3, p2)));
        cf.z = fmaf(0.5f, p0 + p2, -p1);
        cf.w = fmaf(1.f / 6.f, p3 - p0, 0.5f * (p1 - p2));
      }
      tab[slot] = cf;
    }
    float* hdr = headers + c * kCdfHdr;
    if (lane == 0) {
      hdr[0] = lo;
      hdr[1] = inv_h;
      hdr[4] = fmaf(-lo, inv_h, 1.f);  // t' = z inv_h + hdr[4]
      hdr[5] = (float)(cells + 2) - 0.001f;  // clamp of t': slot cells + 1 is the guard above the range
      hdr[2] = __int_as_float(n_sharp);
      hdr[3] = __int_as_float(overflow ? 1 : 0);
      counts[c] = 0u;  // the sampling kernel adds to it
    }
    if (lane < kCdfMaxSharp) {
      const float2 sh = lane < n_sharp ? s_sharp[warp][lane] : make_float2(0.f, 1e30f);
      hdr[8 + 2 * lane] = sh.x;
      hdr[9 + 2 * lane] = sh.y;
    }
    n_sharp_ctx += (lane == 0 && n_sharp > 0) ? 1 : 0;
    n_over_ctx += (lane == 0 && overflow) ? 1 : 0;
    __syncwarp();  // s_F / s_sharp are rewritten by the next context
  }
  if (stats) {
    n_terms = warp_sum(n_terms);
    n_sharp_ctx = warp_sum(n_sharp_ctx);
    n_over_ctx = warp_sum(n_over_ctx);
    if (lane == 0) {
      if (n_terms) atomicAdd(stats + 1, n_terms);
      if (n_sharp_ctx) atomicAdd(stats + 2, n_sharp_ctx);
      if (n_over_ctx) atomicAdd(stats + 3, n_over_ctx);
    }
  }
}

// ---- sample --------------------------------------------------------------------------------------
struct CdfCtx {
  float inv_h, t0;   // table coordinate t' = z inv_h + t0 (slot 0: below the range, slot cells + 1: above)
  float tmax;        // cells + 2 - 0.001
  uint32_t tab;      // shared-window address of this warp's staged table (2 KB aligned)
};

// F(z) from the staged table: t' + 2^19 rounded toward zero keeps floor(16 t') in the mantissa, i.e. bits
// 4.. are the slot = exactly its byte offset 16 floor(t'); no conversion unit, no range selects.
__device__ __forceinline__ float cdf_eval_t(float tp, const CdfCtx& cx) {
  const float t = fminf(fmaxf(tp, 0.f), cx.tmax);
  const int u = __float_as_int(__fadd_rz(t, 524288.f));
  const float f = t - (__int_as_float(u & ~0xf) - 524288.f);
  const float4 c = lds128(cx.tab | (uint32_t)(u & 0x7f0));
  return fmaf(fmaf(fmaf(c.w, f, c.z), f, c.y), f, c.x);
}

__device__ __forceinline__ float cdf_eval(float z, const CdfCtx& cx) {
  const float t = fminf(fmaxf(fmaf(z, cx.inv_h, cx.t0), 0.f), cx.tmax);
  const int u = __float_as_int(__fadd_rz(t, 524288.f));
  const float f = t - (__int_as_float(u & ~0xf) - 524288.f);
  const float4 c = lds128(cx.tab | (uint32_t)(u & 0x7f0));
  return fmaf(fmaf(fmaf(c.w, f, c.z), f, c.y), f, c.x);
}

__device__ __forceinline__ float sharp_factor(float z, int n_sharp, const float2* __restrict__ sharp,
                                              const float4* __restrict__ s_lnphi) {
  float F = 1.f;
  for (int k = 0; k < n_sharp; ++k) {  // warp-uniform trip count
    const float2 ab = sharp[k];
    F *= exp_approx(lnphi(fmaf(ab.x, z, ab.y), s_lnphi));
  }
  return F;
}

__device__ __forceinline__ uint32_t prob_to_u32(float F) {
  uint32_t r;  // floor(F 2^32), saturating: F >= 1 -> 0xffffffff, F <= 0 (or NaN) -> 0
  asm("cvt.rzi.sat.u32.f32 %0, %1;" : "=r"(r) : "f"(F * 4294967296.f));
  return r;
}

__global__ void __launch_bounds__(kCdfThreads)
pcs_cdf_sample_kernel(const float4* __restrict__ tables, const float* __restrict__ headers, int64_t n_c,
                      int64_t num_samples, const __grid_constant__ RoundKeys rk, uint32_t offset,
                      int64_t ctx_offset, int splits, uint32_t* __restrict__ counts,
                      unsigned long long* __restrict__ stats) {
  __shared__ float4 s_lnphi[kLnPhiCells];
  // the warps' 2 KB tables start on a 2 KB boundary of the shared window (the slot's byte offset is OR-ed
  // into the base): one spare table of slack, aligned by hand
  __shared__ __align__(16) float4 s_tab_raw[(kCdfWarps + 1) * kCdfStride];
  __shared__ float2 s_sharp[kCdfWarps][kCdfMaxSharp];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_lnphi(s_lnphi, tid, kCdfThreads);
  __syncthreads();
  const uint32_t tab_base = ((smem_u32(s_tab_raw) + 2047u) & ~2047u) + 2048u * (uint32_t)warp;
  float4* const my_tab = s_tab_raw + (tab_base - smem_u32(s_tab_raw)) / 16u;
  const int64_t ncalls = (num_samples + 1) >> 1;
  const uint32_t odd_call = (num_samples & 1) ? (uint32_t)(ncalls - 1) : 0xffffffffu;  // its second sample does not exist
  const int64_t items = n_c * splits;
  const int64_t gw = (int64_t)blockIdx.x * kCdfWarps + warp, nw = (int64_t)gridDim.x * kCdfWarps;
  unsigned long long total = 0, calls = 0;
  CdfCtx cx;
  cx.tab = tab_base;
  const float2* sharp = s_sharp[warp];
  for (int64_t item = gw; item < items; item += nw) {
    const int64_t c = item / splits;
    const int sp = (int)(item - c * splits);
    const uint32_t i_lo = (uint32_t)(ncalls * sp / splits), i_hi = (uint32_t)(ncalls * (sp + 1) / splits);
    const float* hdr = headers + c * kCdfHdr;
    cx.inv_h = hdr[1];
    cx.t0 = hdr[4];
    cx.tmax = hdr[5];
    const int n_sharp = __float_as_int(hdr[2]);
    const float4* gt = tables + c * kCdfStride;
    __syncwarp();  // the previous item's reads of this warp's stage are done
#pragma unroll
    for (int r = 0; r < 4; ++r) my_tab[32 * r + lane] = gt[32 * r + lane];
    if (lane < kCdfMaxSharp) s_sharp[warp][lane] = make_float2(hdr[8 + 2 * lane], hdr[9 + 2 * lane]);
    __syncwarp();
    const uint32_t cw = (uint32_t)(c + ctx_offset);
    unsigned int hits = 0;
    if (n_sharp == 0) {  // the common case: everything is in the table
#pragma unroll 4
      for (uint32_t i = i_lo + lane; i < i_hi; i += 32) {
        const Words w = philox4x32_10(i, cw, kCdfTag, offset, rk);
        float rad, cs, sn;
        box_muller_parts(w.x, w.y, rad, cs, sn);
        const float rs = rad * cx.inv_h;  // t' = z inv_h + t0 with z = rad cs | rad sn: one multiply less per call
        const uint32_t p1 = prob_to_u32(cdf_eval_t(fmaf(rs, cs, cx.t0), cx));
        const uint32_t p2 = prob_to_u32(cdf_eval_t(fmaf(rs, sn, cx.t0), cx));
        hits += w.z < p1 ? 1u : 0u;
        hits += (w.w < p2 && i != odd_call) ? 1u : 0u;
      }
    } else {
      for (uint32_t i = i_lo + lane; i < i_hi; i += 32) {
        const Words w = philox4x32_10(i, cw, kCdfTag, offset, rk);
        float z1, z2;
        box_muller(w.x, w.y, z1, z2);
        const uint32_t p1 = prob_to_u32(cdf_eval(z1, cx) * sharp_factor(z1, n_sharp, sharp, s_lnphi));
        const uint32_t p2 = prob_to_u32(cdf_eval(z2, cx) * sharp_factor(z2, n_sharp, sharp, s_lnphi));
        hits += w.z < p1 ? 1u : 0u;
        hits += (w.w < p2 && i != odd_call) ? 1u : 0u;
      }
    }
    calls += lane == 0 ? (unsigned long long)(i_hi - i_lo) : 0ull;
    hits = warp_sum(hits);
    if (lane == 0 && hits) {
      atomicAdd(counts + c, hits);  // zeroed by the build kernel
      total += hits;
    }
  }
  if (stats && lane == 0) {
    if (calls) atomicAdd(stats, calls);
    if (total) atomicAdd(stats + 4, total);
  }
}

// Peak probes (crs_peak_probe kinds 8 / 9): the bare arithmetic of one sampling call (Philox4x32-10,
// one Box-Muller pair, two cubic evaluations on register coefficients, two integer compares) and of
// one table term of the build (one FMA, one ln Phi lookup in shared memory, one add) — no global
// memory, no control flow besides the loop.
__global__ void __launch_bounds__(kCdfThreads)
cdf_probe_kernel(int kind, int iters, const __grid_constant__ RoundKeys rk, float* out) {
  __shared__ float4 s_lnphi[kLnPhiCells];
  stage_lnphi(s_lnphi, threadIdx.x, kCdfThreads);
  __syncthreads();
  const uint32_t smp = blockIdx.x * kCdfThreads + threadIdx.x;
  unsigned int hits = 0;
  if (kind == 8) {
    // a 2 KB table per warp in shared memory, like the sampler's stage
    __shared__ __align__(16) float4 s_tab_raw[(kCdfWarps + 1) * kCdfStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab_base = ((smem_u32(s_tab_raw) + 2047u) & ~2047u) + 2048u * (uint32_t)warp;
    float4* const my_tab = s_tab_raw + (tab_base - smem_u32(s_tab_raw)) / 16u;
    for (int r = 0; r < 4; ++r) my_tab[32 * r + lane] = make_float4(0.3f + 1e-3f * (float)(32 * r + lane), 0.2f, -0.05f, 0.01f);
    __syncwarp();
    CdfCtx cx;
    cx.tab = tab_base;
    cx.inv_h = 10.5f;
    cx.t0 = 63.f;
    cx.tmax = 126.999f;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
      const Words w = philox4x32_10(smp, 7u, (uint32_t)it, 0u, rk);
      float rad, cs, sn;
      box_muller_parts(w.x, w.y, rad, cs, sn);
      const float rs = rad * cx.inv_h;
      const uint32_t p1 = prob_to_u32(cdf_eval_t(fmaf(rs, cs, cx.t0), cx));
      const uint32_t p2 = prob_to_u32(cdf_eval_t(fmaf(rs, sn, cx.t0), cx));
      hits += w.z < p1 ? 1u : 0u;
      hits += (w.w < p2 && (uint32_t)it != 0xffffffffu) ? 1u : 0u;
    }
  }
  if (hits == 0xffffffffu) out[0] = 1.f;  // keep the loop alive
}

// kind 9: the bare table term of the build on the 8-way bank-interleaved table (one FMA, one clamp, the
// index split, one 16-byte look-up, three FMAs, one add), four independent accumulators per thread
__global__ void __launch_bounds__(kBuildThreads, 1)
cdf_build_probe_kernel(int iters, float* out) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  float4* s_tab8 = reinterpret_cast<float4*>(dyn_smem);
  for (int i = threadIdx.x; i < kLnPhiCells * 8; i += kBuildThreads) s_tab8[i] = g_lnphi[i >> 3];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t copy = smem_u32(s_tab8) + 16u * (uint32_t)(lane & 7);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float a = 64.f * (1.f + 1e-3f * (float)lane);
  float b = 576.f + 32.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r] += lnphi8(fmaf(a, -2.f + 1.3f * (float)r + 0.09f * (float)lane, b), copy);
    b += 0.05f;
    b = b > 700.f ? 576.f : b;
  }
  if (__float_as_uint(acc[0] + acc[1] + acc[2] + acc[3]) == 0xffffffffu) out[0] = 1.f;  // keep the loop alive
}

struct CdfDevInfo { bool ready = false; int sms = 0; int build_ctas[2] = {0, 0}; int sample_ctas = 0; };
CdfDevInfo g_cdf_info[64];
std::mutex g_cdf_mutex;

// one-off per device: fill the ln Phi table (a 4-block kernel, synchronised once under the lock)
int cdf_device_ready(CdfDevInfo& info) {
  int dev = 0;
  CRS_CUDA_TRY(cudaGetDevice(&dev), "pcs_cdf get device");
  if (dev < 0 || dev >= 64) return CRS_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(g_cdf_mutex);
  if (!g_cdf_info[dev].ready) {
    cudaStream_t s0;
    CRS_CUDA_TRY(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking), "pcs_cdf init stream");
    lnphi_init_kernel<<<kLnPhiCells / 256, 256, 0, s0>>>();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
    cudaStreamDestroy(s0);
    CRS_CUDA_TRY(e, "pcs_cdf ln Phi table");
    CRS_CUDA_TRY(cudaDeviceGetAttribute(&g_cdf_info[dev].sms, cudaDevAttrMultiProcessorCount, dev), "pcs_cdf sm count");
    // resident CTAs of each kernel: the grids are sized to exactly one wave
    int per_sm = 0;
    CRS_CUDA_TRY(cudaFuncSetAttribute(pcs_cdf_build_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBuildSmem), "pcs_cdf smem attr");
    CRS_CUDA_TRY(cudaFuncSetAttribute(pcs_cdf_build_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBuildSmem), "pcs_cdf smem attr");
    g_cdf_info[dev].build_ctas[0] = g_cdf_info[dev].build_ctas[1] = g_cdf_info[dev].sms;  // 128 KB table: one CTA per SM
    CRS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcs_cdf_sample_kernel, kCdfThreads, 0), "pcs_cdf occupancy");
    g_cdf_info[dev].sample_ctas = g_cdf_info[dev].sms * (per_sm > 0 ? per_sm : 1);
    g_cdf_info[dev].ready = true;
  }
  info = g_cdf_info[dev];
  return CRS_OK;
}

}  // namespace

int64_t pcs_cdf_workspace_bytes(int64_t n_c) {  // tables | headers | work ticket
  return n_c <= 0 ? 0 : n_c * (int64_t)(kCdfStride * sizeof(float4) + kCdfHdr * sizeof(float)) + 64;
}

int launch_pcs_counts_cdf(const void* mean, const void* var, int64_t n_c, int K, int64_t ld, int dtype,
                          int64_t num_samples, uint64_t seed, uint32_t offset, int64_t ctx_offset,
                          uint32_t* counts, uint64_t* stats, void* workspace, cudaStream_t s) {
  if (!mean || !var || !counts || n_c < 0 || K < 2 || ld < K || num_samples < 0) return CRS_ERR_INVALID_ARG;
  if (n_c > 0 && (!workspace || ((uintptr_t)workspace & 15))) return CRS_ERR_INVALID_ARG;
  if (num_samples > 0x1fffffffeLL) return CRS_ERR_UNSUPPORTED;  // 32-bit call counter, two samples per call
  if (n_c == 0) return CRS_OK;
  CdfDevInfo info;
  int rc;
  if ((rc = cdf_device_ready(info)) != CRS_OK) return rc;
  if (stats) CRS_CUDA_TRY(cudaMemsetAsync(stats, 0, 6 * sizeof(uint64_t), s), "pcs_cdf memset stats");
  float4* tables = reinterpret_cast<float4*>(workspace);
  float* headers = reinterpret_cast<float*>(tables + n_c * kCdfStride);
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(headers + n_c * kCdfHdr);
  CRS_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), s), "pcs_cdf memset ticket");
  const int64_t want_b = (n_c + kBuildWarps - 1) / kBuildWarps;
  const int64_t res_b = info.build_ctas[dtype == CRS_F64 ? 0 : 1];
  const int blocks_b = (int)(want_b < res_b ? want_b : res_b);
  if (dtype == CRS_F64)
    pcs_cdf_build_kernel<double><<<blocks_b, kBuildThreads, kBuildSmem, s>>>((const double*)mean, (const double*)var, n_c, K, ld, tables,
                                                                  headers, counts, ticket, (unsigned long long*)stats);
  else if (dtype == CRS_F32)
    pcs_cdf_build_kernel<float><<<blocks_b, kBuildThreads, kBuildSmem, s>>>((const float*)mean, (const float*)var, n_c, K, ld, tables,
                                                                 headers, counts, ticket, (unsigned long long*)stats);
  else
    return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "pcs_cdf build launch");
  if (num_samples == 0) return CRS_OK;
  // sampling: every (context, range) item costs the same, so a static round-robin over warps is
  // balanced as soon as each warp holds many items; contexts are split when they are too few for that
  const int64_t ncalls = (num_samples + 1) >> 1;
  const int64_t warps = (int64_t)info.sample_ctas * kCdfWarps;
  int64_t splits = 1;
  if (n_c < 16 * warps) splits = (16 * warps + n_c - 1) / n_c;  // >= 16 items per warp: the last round is <= 6 % of the work
  const int64_t max_splits = ncalls / 64 > 0 ? ncalls / 64 : 1;  // >= 2 calls per lane and item
  if (splits > max_splits) splits = max_splits;
  const int64_t items = n_c * splits;
  const int64_t want_s = (items + kCdfWarps - 1) / kCdfWarps;
  const int blocks_s = (int)(want_s < info.sample_ctas ? want_s : info.sample_ctas);
  const RoundKeys rk = round_keys((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  pcs_cdf_sample_kernel<<<blocks_s, kCdfThreads, 0, s>>>(tables, headers, n_c, num_samples, rk, offset, ctx_offset,
                                                         (int)splits, counts, (unsigned long long*)stats);
  CRS_CUDA_TRY(cudaGetLastError(), "pcs_cdf sample launch");
  return CRS_OK;
}

int launch_cdf_probe(int kind, int iters, int blocks, void* out, cudaStream_t s) {
  CdfDevInfo info;
  int rc;
  if ((rc = cdf_device_ready(info)) != CRS_OK) return rc;
  if (kind == 9) {
    CRS_CUDA_TRY(cudaFuncSetAttribute(cdf_build_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBuildSmem), "cdf_probe attr");
    cdf_build_probe_kernel<<<info.sms, kBuildThreads, kBuildSmem, s>>>(iters, (float*)out);  // one CTA per SM, like the build
  } else {
    cdf_probe_kernel<<<blocks, kCdfThreads, 0, s>>>(kind, iters, round_keys(1u, 2u), (float*)out);
  }
  CRS_CUDA_TRY(cudaGetLastError(), "cdf_probe launch");
  return CRS_OK;
}

}  // namespace crs
