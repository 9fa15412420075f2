// crs_pcs_counts (K3): Monte-Carlo correct-selection counts per context — replaces the sampling
// half of estimate_current_generalized_pcs for a ModelListGP
// (contextual_rs/generalized_pcs.py:441-475): draw posterior samples at every (arm, context),
// order arms by posterior mean, delta = y_(1) - max_{j>=2} y_(j), func_I = (delta > 0), mean over
// samples.  The reference materialises an S x K x n_c sample tensor (550 GB at config 3); here
// nothing S-sized ever leaves the SM: each thread owns a quad of 4 samples, draws its normals from
// Philox4x32-10 keyed by (seed; sample/4, context, arm, offset) and keeps only a running max.
//
// Exact screening: the Box-Muller variates are built from 23-bit uniforms, so |z| <= 5.7683.  An
// arm j with  mu_j + ZB*sd_j < mu_b - ZB*sd_b  (ZB = 5.77) can never beat the best arm b in ANY
// draw of this stream, so it is dropped before sampling without changing a single count.  A quad
// also stops early once all of its 4 samples have already failed.  Both shortcuts leave the counts
// identical to the unscreened estimator; executed / skipped draws are reported in `stats`.
#include <math_constants.h>
#include "crs_common.cuh"

namespace crs {
namespace {

constexpr int kPcsThreads = 256;
constexpr int kQuadsPerThread = 16;              // quads of 4 samples per thread per work item
constexpr float kZBound = 5.77f;                 // > sqrt(-2 ln 2^-24) = 5.7683 + MUFU slack
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct Words { uint32_t x, y, z, w; };

__device__ __forceinline__ Words philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return Words{c0, c1, c2, c3};
}

__device__ __forceinline__ float approx_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// (wa, wb) -> two N(0,1) draws; see oracle/philox_ref.py for the stream definition.
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1) {
  // (1 + m 2^-23) - (1 - 2^-24) = m 2^-23 + 2^-24 exactly: one FADD builds the open-interval uniform
  const float a = __uint_as_float(0x3f800000u | (wa >> 9)) - 0.99999994039535522f;
  const float b = __uint_as_float(0x3f800000u | (wb >> 9)) - 1.0f;
  const float rad = approx_sqrt(-1.3862943611198906f * __log2f(a));  // sqrt(-2 ln a)
  float sn, cs;
  __sincosf(6.2831853071795865f * b, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

struct Cand { float mu, sd; int arm; };  // centred mean, std-dev, global arm id

template <typename T>
__global__ void __launch_bounds__(kPcsThreads)
pcs_counts_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int K,
                  int64_t ld, int64_t num_samples, uint32_t k0, uint32_t k1, uint32_t offset,
                  int arm_offset, int64_t ctx_offset, int chunks_per_ctx, uint32_t* counts,
                  unsigned long long* stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cand* cand = reinterpret_cast<Cand*>(smem_raw);  // [K]
  __shared__ T s_bm[kPcsThreads / 32];
  __shared__ int s_bi[kPcsThreads / 32];
  __shared__ int s_m;       // number of competitive arms
  __shared__ int s_best;
  __shared__ float s_sdb;
  __shared__ unsigned int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nq = (num_samples + 3) >> 2;
  const int64_t items = n_c * chunks_per_ctx;
  unsigned long long drawn = 0, skipped = 0;

  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t c = item / chunks_per_ctx;
    const int chunk = (int)(item - c * chunks_per_ctx);
    const T* mrow = mean + c * ld;
    const T* vrow = var + c * ld;
    __syncthreads();
    // ---- predicted best arm: first arg-max of the mean row (generalized_pcs.py:464) -------
    T bm = -CUDART_INF;
    int bi = 0x7fffffff;
    for (int a = tid; a < K; a += kPcsThreads) {
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
    if (lane == 0) {
      s_bm[warp] = bm;
      s_bi[warp] = bi;
    }
    if (tid == 0) {
      s_m = 0;
      s_cnt = 0;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kPcsThreads / 32; ++w)
        if (s_bm[w] > bm || (s_bm[w] == bm && s_bi[w] < bi)) {
          bm = s_bm[w];
          bi = s_bi[w];
        }
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
    // ---- exact screen + compaction -----------------------------------------------------------
    for (int a0 = 0; a0 < K; a0 += kPcsThreads) {
      const int a = a0 + tid;
      bool keep = false;
      float mu = 0.f, sd = 0.f;
      if (a < K && a != best) {
        mu = (float)(mrow[a] - mu_b);  // centred in the grid dtype, then fp32
        sd = (float)sqrt((double)vrow[a]);
        keep = !(fmaf(sd, kZBound, mu) < low_b);
      }
      const unsigned ball = __ballot_sync(kFull, keep);
      int base = 0;
      if (lane == 0 && ball) base = atomicAdd(&s_m, __popc(ball));
      base = __shfl_sync(kFull, base, 0);
      if (keep) {
        const int pos = base + __popc(ball & ((1u << lane) - 1u));
        cand[pos].mu = mu;
        cand[pos].sd = sd;
        cand[pos].arm = a + arm_offset;
      }
    }
    __syncthreads();
    const int M = s_m;
    const uint32_t cw = (uint32_t)(c + ctx_offset);
    // ---- sampling ------------------------------------------------------------------------------
    unsigned int hits = 0;
    const int64_t q0 = (int64_t)chunk * kPcsThreads * kQuadsPerThread;
    for (int qq = 0; qq < kQuadsPerThread; ++qq) {
      const int64_t q = q0 + (int64_t)qq * kPcsThreads + tid;
      const bool active = q < nq;
      if (!__any_sync(kFull, active)) break;
      const int ns = active ? (int)min((int64_t)4, num_samples - 4 * q) : 0;
      float yb[4], mr[4];
      {
        const Words w = philox4x32_10((uint32_t)q, cw, (uint32_t)(best + arm_offset), offset, k0, k1);
        float z[4];
        box_muller(w.x, w.y, z[0], z[1]);
        box_muller(w.z, w.w, z[2], z[3]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          yb[s] = s < ns ? sd_b * z[s] : -CUDART_INF_F;  // padded samples count as failures
          mr[s] = -CUDART_INF_F;
        }
      }
      // two competitor arms per iteration: two independent Philox / Box-Muller chains in flight
      int e = 0;
      for (; e + 1 < M; e += 2) {
        bool alive = false;
#pragma unroll
        for (int s = 0; s < 4; ++s) alive = alive || (yb[s] > mr[s]);
        if (!__any_sync(kFull, alive)) break;  // every sample of every lane already failed
        const Cand c0 = cand[e], c1 = cand[e + 1];
        const Words w0 = philox4x32_10((uint32_t)q, cw, (uint32_t)c0.arm, offset, k0, k1);
        const Words w1 = philox4x32_10((uint32_t)q, cw, (uint32_t)c1.arm, offset, k0, k1);
        float z0[4], z1[4];
        box_muller(w0.x, w0.y, z0[0], z0[1]);
        box_muller(w1.x, w1.y, z1[0], z1[1]);
        box_muller(w0.z, w0.w, z0[2], z0[3]);
        box_muller(w1.z, w1.w, z1[2], z1[3]);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          mr[s] = fmaxf(mr[s], fmaxf(fmaf(c0.sd, z0[s], c0.mu), fmaf(c1.sd, z1[s], c1.mu)));
      }
      if (e < M) {
        bool alive = false;
#pragma unroll
        for (int s = 0; s < 4; ++s) alive = alive || (yb[s] > mr[s]);
        if (__any_sync(kFull, alive)) {
          const Cand cd = cand[e];
          const Words w = philox4x32_10((uint32_t)q, cw, (uint32_t)cd.arm, offset, k0, k1);
          float z[4];
          box_muller(w.x, w.y, z[0], z[1]);
          box_muller(w.z, w.w, z[2], z[3]);
#pragma unroll
          for (int s = 0; s < 4; ++s) mr[s] = fmaxf(mr[s], fmaf(cd.sd, z[s], cd.mu));
          ++e;
        }
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) hits += (s < ns && yb[s] > mr[s]) ? 1u : 0u;
      if (active) {
        drawn += 4ull * (unsigned long long)(e + 1);  // e competitor arms + the best arm
        skipped += 4ull * (unsigned long long)(K - 1 - e);
      }
    }
    hits = warp_sum(hits);
    if (lane == 0 && hits) atomicAdd(&s_cnt, hits);
    __syncthreads();
    if (tid == 0 && s_cnt) atomicAdd(counts + c, s_cnt);
  }
  if (stats) {
    drawn = warp_sum(drawn);
    skipped = warp_sum(skipped);
    if (lane == 0) {
      atomicAdd(stats, drawn);
      atomicAdd(stats + 1, skipped);
    }
  }
}

__global__ void philox_words_kernel(const uint32_t* __restrict__ ctr, int64_t n, uint32_t k0,
                                    uint32_t k1, uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Words w = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1);
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
  if (num_samples > 0xffffffffLL) return CRS_ERR_UNSUPPORTED;
  const size_t smem = (size_t)K * sizeof(Cand);
  if (smem > 200 * 1024) return CRS_ERR_UNSUPPORTED;
  if (n_c == 0) return CRS_OK;
  CRS_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (size_t)n_c, s), "pcs memset");
  if (stats) CRS_CUDA_TRY(cudaMemsetAsync(stats, 0, 2 * sizeof(uint64_t), s), "pcs memset stats");
  if (num_samples == 0) return CRS_OK;
  const int64_t nq = (num_samples + 3) / 4;
  const int64_t per_item = (int64_t)kPcsThreads * kQuadsPerThread;
  const int chunks = (int)((nq + per_item - 1) / per_item);
  const int64_t items = n_c * chunks;
  int blocks = (int)(items < 148 * 8 ? items : 148 * 8);
  const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
  if (dtype == CRS_F64) {
    auto kern = pcs_counts_kernel<double>;
    if (smem > 48 * 1024) CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pcs attr");
    kern<<<blocks, kPcsThreads, smem, s>>>((const double*)mean, (const double*)var, n_c, K, ld, num_samples, k0, k1, offset, arm_offset, ctx_offset, chunks, counts, (unsigned long long*)stats);
  } else if (dtype == CRS_F32) {
    auto kern = pcs_counts_kernel<float>;
    if (smem > 48 * 1024) CRS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pcs attr");
    kern<<<blocks, kPcsThreads, smem, s>>>((const float*)mean, (const float*)var, n_c, K, ld, num_samples, k0, k1, offset, arm_offset, ctx_offset, chunks, counts, (unsigned long long*)stats);
  } else {
    return CRS_ERR_INVALID_ARG;
  }
  CRS_CUDA_TRY(cudaGetLastError(), "pcs_counts launch");
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
