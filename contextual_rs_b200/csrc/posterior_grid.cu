// crs_posterior_grid (K1): the ModelListGP posterior mean / marginal variance over the full
// arm x context grid — replaces `model.posterior(context_set)` at
// contextual_rs/contextual_rs_strategies.py:133-136 (and generalized_pcs.py:443-444,
// continuous_context.py:68,96, experiments/wsc_experiments/main.py:460).
//
// Mapping: CTA = (tile of 128 contexts) x (tile of <= 8 arms); thread = one context.  An arm's
// prepared state (head, alpha, scaled train inputs, first 32-column panel of o*L^-T) is staged in
// shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier), double-buffered so that the
// next arm of the tile streams in while the current one is computed.  Per (context, arm):
//   phase A  c_j = corr(|| x~* - X~_j ||^2), j = 0..n-1   (direct differences; two observations
//            per iteration for ILP), macc += c_j alpha_j, c_j parked in shared memory [j][thread]
//   phase B  v = (o L^-1) c in register chunks of VB rows: v[i] += linv_t[j][i] c_j for i >= j,
//            rows visited in 8-row segments so that the triangular structure is resolved at
//            compile time (no per-row predicates); q += sum v_i^2
//   mu = ybar + s (m + macc),  var = max(s^2 (o - q), floor)
// The kernel is FP64/FP32-pipe bound (~55 flop per output byte); HBM traffic is the
// context tile in and the two tables out.  Tables are [n_c, ld] arm-fastest, so each thread buffers
// its results in shared memory and writes them as one contiguous run per arm tile.
//
// fp64 transcendental cost dominates at n ~ 20, so exp() is a 64-entry-table + degree-5 polynomial
// (10 FP64 ops) and sqrt() is rsqrt.approx + one coupled Newton step + one correction (branch-free).
#include <type_traits>
#include <utility>
#include <cstdlib>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kGridThreads = 128;
constexpr int kMaxArmTile = 8;
#ifndef CRS_GRID_MIN_BLOCKS
#define CRS_GRID_MIN_BLOCKS 4
#endif
constexpr int kMinBlocks = CRS_GRID_MIN_BLOCKS;
#ifndef CRS_PHASE_A_UNROLL
#define CRS_PHASE_A_UNROLL 4
#endif
constexpr int kPhaseAUnroll = CRS_PHASE_A_UNROLL;  // divides 4 (rows are padded to a multiple of 4)

// ---- fp64 exp(x), x <= 0: x = (64 k + i) ln2/64 + r, exp(x) = 2^k * 2^(i/64) * e^r ------------
__constant__ double kExpTable[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ s_tab) {
  // t = x * 64/ln2 + 1.5*2^52: the integer n = 64 k + i lands in the low mantissa bits of t
  const double t = fma(x, 92.33248261689366, 6755399441055744.0);
  const double nd = t - 6755399441055744.0;
  double r = fma(nd, -0.010830424696249145, x);   // ln2/64 (high part)
  r = fma(nd, -3.623510646634843e-19, r);       // ln2/64 (low part)
  int n = __double2loint(t);
  n = max(n, -64000);                             // exp(-693) ~ 1e-301: clamp instead of underflow
  double p = 1.0 / 120.0;                         // e^r, |r| <= ln2/128: degree 5, error 3.5e-17
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double y = s_tab[n & 63] * p;
  return __hiloint2double(__double2hiint(y) + ((n >> 6) << 20), __double2loint(y));
}

// sqrt(x) for x > 0 (callers add 1e-290, harmless: the result only enters s and exp(-s)).
template <int ITERS>
__device__ __forceinline__ double sqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    if (it + 1 < ITERS) h = fma(h, r, h);  // the last h only scales the 2^-44 residual below: 22 bits suffice
  }
  const double e = fma(-g, g, x);
  return fma(e, h, g);
}

template <typename T> struct Corr;
template <> struct Corr<double> {
  static __device__ __forceinline__ double eval(double r2, int kernel, const double* s_tab) {
    if (kernel == CRS_MATERN52) {
      const double s = sqrt_pos<1>(fma(r2, 5.0, 1e-290));        // sqrt5 * r (1 Newton step + correction: exact on 2e6 samples)
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

// 16-byte shared-memory vector loads (the compiler cannot prove the alignment of the runtime-sized
// stage buffers, so it is spelled out).
__device__ __forceinline__ void lds_vec(const double* p, double (&o)[2]) {
  const double2 t = *reinterpret_cast<const double2*>(p);
  o[0] = t.x; o[1] = t.y;
}
__device__ __forceinline__ void lds_vec(const float* p, float (&o)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
template <typename T> struct VecLen { static constexpr int value = 16 / sizeof(T); };

// v[8 ib .. 8 ib + 7] += row[8 ib ..] * c for the column blocks ib >= FIRST of a VB-wide chunk
template <typename T, int VB, int FIRST>
__device__ __forceinline__ void axpy_blocks(T (&v)[VB], const T* __restrict__ row, T c) {
  constexpr int W = VecLen<T>::value;
#pragma unroll
  for (int i = FIRST * 8; i < VB; i += W) {
    T t[W];
    lds_vec(row + i, t);
#pragma unroll
    for (int u = 0; u < W; ++u) v[i + u] = fma(t[u], c, v[i + u]);
  }
}

// squared scaled distance to one staged training row (D % vector length handled by D >= W or scalar)
template <typename T, int D>
__device__ __forceinline__ T sq_dist(const T (&xq)[D], const T* __restrict__ xrow) {
  constexpr int W = VecLen<T>::value;
  T r = T(0);
  if constexpr (D % W == 0) {
#pragma unroll
    for (int dd = 0; dd < D; dd += W) {
      T t[W];
      lds_vec(xrow + dd, t);
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const T df = xq[dd + u] - t[u];
        r = fma(df, df, r);
      }
    }
  } else {
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      const T df = xq[dd] - xrow[dd];
      r = fma(df, df, r);
    }
  }
  return r;
}

// ---- phase B ---------------------------------------------------------------------------------
// q[context] = || (o L^-1) c ||^2 restricted to the VB rows [p0, p0 + VB) of v, for the 32 contexts
// of one warp.  c lives in s_c[j][context] (row stride CS), the panel in s_panel[j][i - p0].

// FP32: one context per thread, v in registers, rows in 8-row segments so that the triangular
// structure is resolved at compile time.
template <int VB, int CS>
__device__ __forceinline__ float chunk_sqnorm(const float* __restrict__ s_panel, const float* __restrict__ s_c,
                                              int p0, int jend, int tid, float*) {
  constexpr int LD = PanelLd<float>::value;
  float v[VB];
#pragma unroll
  for (int i = 0; i < VB; ++i) v[i] = 0.f;
  for (int j = 0; j < p0; ++j)  // rows above this chunk: every column is below the diagonal
    axpy_blocks<float, VB, 0>(v, s_panel + j * LD, s_c[j * CS + tid]);
  const float* crow = s_c + (size_t)p0 * CS + tid;
  const float* prow = s_panel + (size_t)p0 * LD;
  const int rows_in = jend - p0;
#pragma unroll
  for (int sg = 0; sg < VB / 8; ++sg) {
    if (sg * 8 < rows_in) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (sg * 8 + jj < rows_in) {
          const float cj = crow[(sg * 8 + jj) * CS];
          const float* row = prow + (sg * 8 + jj) * LD;
          if (sg == 0) axpy_blocks<float, VB, 0>(v, row, cj);
          else if (sg == 1) axpy_blocks<float, VB, (VB > 8 ? 1 : 0)>(v, row, cj);
          else if (sg == 2) axpy_blocks<float, VB, (VB > 16 ? 2 : 0)>(v, row, cj);
          else axpy_blocks<float, VB, (VB > 24 ? 3 : 0)>(v, row, cj);
        }
      }
    }
  }
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VB; ++i) q = fmaf(v[i], v[i], q);
  return q;
}

// FP64: the contraction V[VB x 32] = Linv[VB x jend] C[jend x 32] runs on the FP64 tensor path
// (mma.sync.m8n8k4.f64 -> SASS DMMA): same FLOP rate as DFMA on B200 (measured 37.1 vs 36.6
// TFLOP/s) but one fragment load feeds 8 FMAs per lane instead of 2, which takes the kernel off
// the shared-memory-wavefront limit the scalar version sat on (ncu, profiles/r01b).  A fragment:
// Linv[p0 + 8 mt + lane/4][4 ks + lane%4]; B fragment: C[4 ks + lane%4][8 nt + lane/4]; tiles that
// lie entirely above the diagonal are skipped.  Rows >= n of both operands are finite / zero.
template <int VB, int CS>
__device__ __forceinline__ double chunk_sqnorm(const double* __restrict__ s_panel, const double* __restrict__ s_c,
                                               int p0, int jend, int tid, double* s_q) {
  constexpr int LD = PanelLd<double>::value;
  constexpr int MT = VB / 8;
  const int lane = tid & 31, wbase = tid & ~31;
  const int l4 = lane & 3, l8 = lane >> 2;
  double acc[MT][4][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int ksteps = (jend + 3) >> 2;
  const double* bptr = s_c + (size_t)l4 * CS + wbase + l8;
  const double* aptr = s_panel + (size_t)l4 * LD + l8;
  // one k-step (4 rows j) for the row tiles mt >= first (tiles above the diagonal are all zero)
  auto kstep = [&](int ks, auto first_tag) {
    constexpr int FIRST = decltype(first_tag)::value;
    double b[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) b[nt] = bptr[(size_t)ks * 4 * CS + nt * 8];
#pragma unroll
    for (int mt = FIRST; mt < MT; ++mt) {
      const double a = aptr[(size_t)ks * 4 * LD + mt * 8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[mt][nt][0]), "+d"(acc[mt][nt][1]) : "d"(a), "d"(b[nt]));
    }
  };
  const int ks0 = p0 >> 2;  // k-steps above the chunk: every tile is below the diagonal
  for (int ks = 0; ks < ks0; ++ks) kstep(ks, std::integral_constant<int, 0>{});
  // k-steps inside the chunk: steps 2 sg, 2 sg + 1 only reach row tiles mt >= sg
  [&]<int... SG>(std::integer_sequence<int, SG...>) {
    (([&] {
       if (ks0 + 2 * SG < ksteps) kstep(ks0 + 2 * SG, std::integral_constant<int, SG>{});
       if (ks0 + 2 * SG + 1 < ksteps) kstep(ks0 + 2 * SG + 1, std::integral_constant<int, SG>{});
     }()),
     ...);
  }(std::make_integer_sequence<int, MT>{});
  // D fragment: V[p0 + 8 mt + l8][8 nt + 2 l4 + e]; sum of squares over rows = over mt and l8
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    double q0 = 0.0, q1 = 0.0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      q0 = fma(acc[mt][nt][0], acc[mt][nt][0], q0);
      q1 = fma(acc[mt][nt][1], acc[mt][nt][1], q1);
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      q0 += __shfl_xor_sync(kFull, q0, o);
      q1 += __shfl_xor_sync(kFull, q1, o);
    }
    if (l8 == 0) {
      s_q[wbase + nt * 8 + 2 * l4] = q0;
      s_q[wbase + nt * 8 + 2 * l4 + 1] = q1;
    }
  }
  __syncwarp();
  const double q = s_q[tid];
  __syncwarp();
  return q;
}

template <typename T> struct CStride { static constexpr int value = sizeof(T) == 8 ? kGridThreads + 8 : kGridThreads; };

// DIRECT = true: results go straight to the tables (no [2][arm_tile][128] staging buffer), which takes the
// CTA under 45 KB of shared memory at n <= 24 and, with <= 96 registers, lets FIVE CTAs share an SM
// instead of four (20 warps: the FP64 pipe was 50 % busy at 16).  The 8-byte stores of one context's
// arm tile fill the same two 32-byte sectors within microseconds and merge in L2.
template <typename T, int D, int VB, bool DIRECT>
__global__ void __launch_bounds__(kGridThreads, DIRECT ? kMinBlocks + 1 : kMinBlocks)
posterior_grid_kernel(crs_gp_state st, const T* __restrict__ ctx, int64_t n_c, int arm_begin,
                      int arm_end, int arm_tile, int n_rows, T* __restrict__ mean,
                      T* __restrict__ var, int64_t ld, int col_offset) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int LD = PanelLd<T>::value;   // panel row stride
  constexpr int CS = CStride<T>::value;   // row stride of the parked correlations
  const int tid = threadIdx.x;
  const int ncap = st.n_cap, d = st.dim;
  // shared layout: bars[4] u64 | tab[64] T | q[128] T | stage[2] { head[40] | alpha[n_rows] |
  //                xs[n_rows][D] | panel[n_rows][LD] } | c[n_rows][CS] | out[2][arm_tile][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  T* s_tab = reinterpret_cast<T*>(smem_raw + 32);
  T* s_q = s_tab + 64;
  T* s_stage = s_q + kGridThreads;
  const int stage_elems = kHeadElems + n_rows + n_rows * D + n_rows * LD;
  T* s_c = s_stage + 2 * stage_elems;
  T* s_out = s_c + n_rows * CS;

  const int64_t c = (int64_t)blockIdx.x * kGridThreads + tid;
  const bool live = c < n_c;
  const int a0 = arm_begin + blockIdx.y * arm_tile;
  const int a1 = min(a0 + arm_tile, arm_end);
  const int na = a1 - a0;

  const T* g_head = reinterpret_cast<const T*>(st.arm_head);
  const T* g_alpha = reinterpret_cast<const T*>(st.alpha);
  const T* g_xs = reinterpret_cast<const T*>(st.xs);
  const T* g_lt = reinterpret_cast<const T*>(st.linv_t);
  const long long lt_stride = linv_stride(ncap, sizeof(T));

  // TMA producer (thread 0): stage the state of the t-th arm of this tile into buffer t & 1
  auto issue = [&](int t) {
    const int k = a0 + t;
    T* stg = s_stage + (t & 1) * stage_elems;
    const int n = min(min(st.n_obs[k], ncap), n_rows);
    const int n4 = (n + 3) & ~3;
    const int rows0 = min(n4, kPanel);
    uint64_t* bar = &bars[t & 1];
    mbar_expect_tx(bar, (uint32_t)sizeof(T) * (kHeadElems + n4 + n4 * D + rows0 * LD));
    tma_bulk_g2s(stg, g_head + (size_t)k * kHeadElems, sizeof(T) * kHeadElems, bar);
    if (n4) {
      tma_bulk_g2s(stg + kHeadElems, g_alpha + (size_t)k * ncap, sizeof(T) * n4, bar);
      tma_bulk_g2s(stg + kHeadElems + n_rows, g_xs + (size_t)k * ncap * D, sizeof(T) * n4 * D, bar);
      tma_bulk_g2s(stg + kHeadElems + n_rows + n_rows * D, g_lt + (size_t)k * lt_stride,
                   sizeof(T) * rows0 * LD, bar);
    }
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  if (tid < 64) s_tab[tid] = (T)kExpTable[tid];
  __syncthreads();
  griddep_wait();                // crs_gp_prepare of the same decision may still be running (PDL); its
                                 // launch may also carry the H2D copy of `ctx` (crs_gao_decide_host)
  griddep_launch_dependents();   // every CTA of this grid is resident by the time the scan may start
  if (tid == 0) {
    issue(0);
    if (na > 1) issue(1);
  }
  T xraw[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) xraw[dd] = (live && dd < d) ? ctx[c * d + dd] : T(0);
  uint32_t xphase = 0;

  for (int t = 0; t < na; ++t) {
    const int k = a0 + t;
    const int n = min(min(st.n_obs[k], ncap), n_rows);
    const int n4 = (n + 3) & ~3;
    T* stg = s_stage + (t & 1) * stage_elems;
    const T* s_head = stg;
    const T* s_alpha = stg + kHeadElems;
    const T* s_xs = s_alpha + n_rows;
    T* s_panel = stg + kHeadElems + n_rows + n_rows * D;
    mbar_wait(&bars[t & 1], (uint32_t)(t >> 1) & 1u);

    // ---- phase A: correlations, kPhaseAUnroll observations per iteration (ILP) ----------------
    // (runs to n4: the padded rows have zero xs / alpha, so they park a finite value and add 0)
    T xq[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) xq[dd] = (xraw[dd] - s_head[kHeadSub + dd]) * s_head[kHeadMul + dd];
    T macc[kPhaseAUnroll];
#pragma unroll
    for (int u = 0; u < kPhaseAUnroll; ++u) macc[u] = T(0);
#pragma unroll 1
    for (int j = 0; j < n4; j += kPhaseAUnroll) {
      T r2[kPhaseAUnroll], cv[kPhaseAUnroll];
#pragma unroll
      for (int u = 0; u < kPhaseAUnroll; ++u) r2[u] = sq_dist<T, D>(xq, s_xs + (j + u) * D);
#pragma unroll
      for (int u = 0; u < kPhaseAUnroll; ++u) cv[u] = Corr<T>::eval(r2[u], st.kernel, s_tab);
#pragma unroll
      for (int u = 0; u < kPhaseAUnroll; ++u) {
        macc[u] = fma(cv[u], s_alpha[j + u], macc[u]);
        s_c[(j + u) * CS + tid] = cv[u];
      }
    }
    T macc_sum = T(0);
#pragma unroll
    for (int u = 0; u < kPhaseAUnroll; ++u) macc_sum += macc[u];
    __syncwarp();  // phase B reads the correlations of the whole warp (fp64 MMA fragments)

    // ---- phase B: q = || (o L^-1) c ||^2, VB rows of v at a time -----------------------------
    T q = T(0);
    for (int p0 = 0; p0 < n; p0 += kPanel) {
      const int jend = min(n4, p0 + VB);  // rows j < jend touch this chunk
      if (p0 > 0) {                       // later panels of a long arm: staged on demand
        __syncthreads();
        if (tid == 0) {
          fence_proxy_async();
          const uint32_t bytes = (uint32_t)sizeof(T) * min(n4, p0 + kPanel) * LD;
          mbar_expect_tx(&bars[2], bytes);
          tma_bulk_g2s(s_panel, g_lt + (size_t)k * lt_stride + panel_offset(ncap, p0 / kPanel, sizeof(T)),
                       bytes, &bars[2]);
        }
        mbar_wait(&bars[2], xphase);
        xphase ^= 1u;
      }
      q += chunk_sqnorm<VB, CS>(s_panel, s_c, p0, jend, tid, s_q);
    }

    const T o = s_head[kHeadScal + 0], m = s_head[kHeadScal + 1];
    const T ybar = s_head[kHeadScal + 2], ysd = s_head[kHeadScal + 3];
    const T mu = fma(ysd, m + macc_sum, ybar);
    T vr = (ysd * ysd) * (o - q);
    vr = vr > variance_floor<T>() ? vr : variance_floor<T>();  // also maps NaN to the floor
    if constexpr (DIRECT) {
      if (live) {
        const int64_t o_idx = c * ld + col_offset + (k - arm_begin);
        mean[o_idx] = mu;
        var[o_idx] = vr;
      }
    } else {
      s_out[t * kGridThreads + tid] = mu;
      s_out[(arm_tile + t) * kGridThreads + tid] = vr;
    }

    __syncthreads();  // every thread is done with this stage buffer
    if (tid == 0 && t + 2 < na) {
      fence_proxy_async();
      issue(t + 2);
    }
  }

  if (!DIRECT && live) {
    T* mrow = mean + c * ld + col_offset + (a0 - arm_begin);
    T* vrow = var + c * ld + col_offset + (a0 - arm_begin);
    for (int a = 0; a < na; ++a) {
      mrow[a] = s_out[a * kGridThreads + tid];
      vrow[a] = s_out[(arm_tile + a) * kGridThreads + tid];
    }
  }
}

template <typename T, int D, int VB, bool DIRECT>
int launch_variant(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                   int arm_end, int n_rows, void* mean, void* var, int64_t ld, int col_offset,
                   cudaStream_t s) {
  const int narms = arm_end - arm_begin;
  const int64_t ctx_tiles = (n_c + kGridThreads - 1) / kGridThreads;
  // Arm tile: as large as possible (contiguous output runs, context tile reused, TMA prefetch of
  // the next arm) while keeping >= ~4 CTAs per SM in flight on 148 SMs for small problems.
  int arm_tile = kMaxArmTile;
  static const int min_ctas = [] {
    const char* e = getenv("CRS_GRID_MIN_CTAS");  // tuning hook (scripts/gpu_tune_k1.sh)
    return e ? atoi(e) : 148 * 4;
  }();
  while (arm_tile > 1 && ctx_tiles * ((narms + arm_tile - 1) / arm_tile) < min_ctas) arm_tile >>= 1;
  auto smem_for = [&](int at) {
    const size_t stage = kHeadElems + (size_t)n_rows + (size_t)n_rows * D + (size_t)n_rows * PanelLd<T>::value;
    return 32 + sizeof(T) * (64 + kGridThreads + 2 * stage + (size_t)n_rows * CStride<T>::value +
                             (DIRECT ? 0 : 2 * (size_t)at * kGridThreads));
  };
  while (arm_tile > 1 && smem_for(arm_tile) > 227 * 1024) arm_tile >>= 1;
  const size_t smem = smem_for(arm_tile);
  if (smem > 227 * 1024) return CRS_ERR_UNSUPPORTED;
  const int arm_tiles = (narms + arm_tile - 1) / arm_tile;
  if (arm_tiles > 65535) return CRS_ERR_UNSUPPORTED;
  auto kern = posterior_grid_kernel<T, D, VB, DIRECT>;
  if (smem > 40 * 1024)
    CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "grid attr");
  dim3 grid((unsigned)ctx_tiles, (unsigned)arm_tiles);
  CRS_CUDA_TRY(launch_pdl(kern, grid, dim3(kGridThreads), smem, s, *st, reinterpret_cast<const T*>(ctx), n_c, arm_begin,
                          arm_end, arm_tile, n_rows, reinterpret_cast<T*>(mean), reinterpret_cast<T*>(var), ld,
                          col_offset), "posterior_grid launch");
  return CRS_OK;
}

template <typename T, int VB>
int dispatch_dim(const crs_gp_state* st, const void* ctx, int64_t n_c, int a0, int a1, int n_rows,
                 void* mean, void* var, int64_t ld, int col, cudaStream_t s) {
  // fp64, n <= 24, large grids: the 5-CTA/SM variant without output staging (see the kernel's comment)
  static const int direct_mode = [] {
    const char* e = getenv("CRS_GRID_DIRECT");  // tuning hook: 0 = never, 1 = always (where instantiated)
    return e ? atoi(e) : -1;
  }();
  if constexpr (sizeof(T) == 8 && VB <= 24) {
    const bool big = n_c * (int64_t)(a1 - a0) >= (int64_t)1 << 20;  // >= ~1 wave of 5 CTAs x 148 SMs x 8 arms x 128 contexts
    if (direct_mode == 1 || (direct_mode < 0 && big)) {
      switch (st->dim_pad) {
        case 4: return launch_variant<T, 4, VB, true>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
        case 8: return launch_variant<T, 8, VB, true>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
      }
    }
  }
  switch (st->dim_pad) {
    case 1: return launch_variant<T, 1, VB, false>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 2: return launch_variant<T, 2, VB, false>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 4: return launch_variant<T, 4, VB, false>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 8: return launch_variant<T, 8, VB, false>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
    case 16: return launch_variant<T, 16, VB, false>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  }
  return CRS_ERR_UNSUPPORTED;
}

template <typename T>
int dispatch_vb(const crs_gp_state* st, const void* ctx, int64_t n_c, int a0, int a1, void* mean,
                void* var, int64_t ld, int col, cudaStream_t s) {
  // n_max: the caller's bound on n_obs (0 = unknown -> the capacity).  Rows are padded to 4.
  const int n_max = st->n_max > 0 && st->n_max <= st->n_cap ? st->n_max : st->n_cap;
  const int n_rows = (n_max + 3) & ~3;
  if (n_max <= 8) return dispatch_dim<T, 8>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  if (n_max <= 16) return dispatch_dim<T, 16>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  if (n_max <= 24) return dispatch_dim<T, 24>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
  return dispatch_dim<T, 32>(st, ctx, n_c, a0, a1, n_rows, mean, var, ld, col, s);
}

__global__ void debug_corr_kernel(const double* __restrict__ r2, int64_t n, int kernel, int variant,
                                  double* __restrict__ out) {
  __shared__ double s_tab[64];
  if (threadIdx.x < 64) s_tab[threadIdx.x] = kExpTable[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = r2[i];
  if (variant == 0) out[i] = Corr<double>::eval(x, kernel, s_tab);          // product path
  else if (variant == 1) out[i] = correlation<double>(x, kernel);            // libdevice exp / sqrt
  else if (variant == 2) out[i] = sqrt_pos<1>(x + 1e-290);
  else if (variant == 3) out[i] = sqrt_pos<2>(x + 1e-290);
  else out[i] = exp_neg(-x, s_tab);
}

}  // namespace

int launch_debug_corr(const double* r2, int64_t n, int kernel, int variant, double* out, cudaStream_t s) {
  if (!r2 || !out || n < 0) return CRS_ERR_INVALID_ARG;
  if (n == 0) return CRS_OK;
  debug_corr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(r2, n, kernel, variant, out);
  CRS_CUDA_TRY(cudaGetLastError(), "debug_corr launch");
  return CRS_OK;
}

int launch_posterior_grid(const crs_gp_state* st, const void* ctx, int64_t n_c, int arm_begin,
                          int arm_end, void* mean, void* var, int64_t ld, int col_offset,
                          cudaStream_t s) {
  if (!st || !ctx || !mean || !var || n_c < 0 || arm_begin < 0 || arm_end > st->num_arms ||
      arm_begin > arm_end || ld < col_offset + (arm_end - arm_begin))
    return CRS_ERR_INVALID_ARG;
  if (st->n_cap > CRS_MAX_OBS || (st->n_cap & 7) || st->dim_pad != dim_pad_of(st->dim)) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0 || arm_begin == arm_end) return CRS_OK;
  if (st->dtype == CRS_F64) return dispatch_vb<double>(st, ctx, n_c, arm_begin, arm_end, mean, var, ld, col_offset, s);
  if (st->dtype == CRS_F32) return dispatch_vb<float>(st, ctx, n_c, arm_begin, arm_end, mean, var, ld, col_offset, s);
  return CRS_ERR_INVALID_ARG;
}

}  // namespace crs
