// The small device-side steps of an ARM-sharded GP-C-OCBA decision — the partition BASELINE.json's
// north_star describes (rank g owns arms [a0, a1) for ALL contexts) — so that the
// whole decision stays on the stream between its collectives (no host parsing):
//   crs_arm_pack_best     local per-context best (mean, variance, global arm id) -> [3][n_c] doubles, the
//                         payload of the per-context all-gather (contextual_rs_strategies.py:137-138: the
//                         head of the descending sort, restricted to the local arms)
//   crs_arm_combine_best  [G][3][n_c] gathered bests -> global best per context (largest mean, lowest arm
//                         id on equal means) = the ext_best_* inputs of crs_ocba_scan
//   crs_arm_fold          G gathered 64-byte records of the local zeta minimisers -> the global minimiser
//                         (strategies.py:145-157): smallest zeta, lowest context, larger mean of the zeta
//                         arm, lower arm id (the single-GPU rule); tie counts of all records at the
//                         minimum added; best_arm filled in from the combined bests
//   crs_arm_finish        psi partial sums all-reduced over ranks -> next_arm (strategies.py:195-201)
#include <math_constants.h>
#include "crs_common.cuh"

namespace crs {
namespace {

template <typename T>
__global__ void __launch_bounds__(256)
arm_pack_best_kernel(const T* __restrict__ mean, const T* __restrict__ var, int64_t n_c, int64_t ld, int arm_offset,
                     const int32_t* __restrict__ best_arm, double* __restrict__ pack) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_c) return;
  const int b = best_arm[c];  // global id
  const int64_t idx = c * ld + (b - arm_offset);
  pack[c] = (double)mean[idx];
  pack[n_c + c] = (double)var[idx];
  pack[2 * n_c + c] = (double)b;
}

template <typename T>
__global__ void __launch_bounds__(256)
arm_combine_best_kernel(const double* __restrict__ pack_all, int world, int64_t n_c, int32_t* __restrict__ ext_arm,
                        T* __restrict__ ext_mean, T* __restrict__ ext_var) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_c) return;
  double bm = -CUDART_INF, bv = 0.0;
  int ba = 0x7fffffff;
  for (int r = 0; r < world; ++r) {
    const double* p = pack_all + (size_t)r * 3 * n_c;
    const double m = p[c];
    const int a = (int)p[2 * n_c + c];
    if (m > bm || (m == bm && a < ba)) {
      bm = m;
      bv = p[n_c + c];
      ba = a;
    }
  }
  ext_arm[c] = ba;
  ext_mean[c] = (T)bm;  // exact: the values were T before the gather
  ext_var[c] = (T)bv;
}

// the mean of the zeta arm at the local minimiser rides in psi_best of the rank's record (unused until select)
template <typename T>
__global__ void arm_record_mean_kernel(crs_decision* decision, const T* __restrict__ mean, int64_t ld, int arm_offset) {
  if (threadIdx.x == 0 && decision->context >= 0)
    decision->psi_best = (double)mean[(int64_t)decision->context * ld + (decision->zeta_arm - arm_offset)];
}

__global__ void arm_fold_kernel(const crs_decision* __restrict__ records, int world, const int32_t* __restrict__ ext_arm,
                                crs_decision* __restrict__ out) {
  if (threadIdx.x != 0) return;
  crs_decision best;
  best.min_zeta = CUDART_INF;
  best.context = -1;
  best.zeta_arm = best.best_arm = best.next_arm = -1;
  best.tie_count = 0;
  best.psi_best = best.psi_rest = 0.0;
  long long ties = 0;
  int bad = 0;
  for (int r = 0; r < world; ++r) {
    const crs_decision q = records[r];
    if (q.reserved[0] != 0 && bad == 0) bad = q.reserved[0];
    if (q.context < 0) continue;
    bool take = false;
    if (best.context < 0 || q.min_zeta < best.min_zeta) {
      take = true;
      ties = q.tie_count;
    } else if (q.min_zeta == best.min_zeta) {
      ties += q.tie_count;
      if (q.context < best.context) take = true;
      else if (q.context == best.context)  // same context on two shards: larger mean, then lower arm id
        take = q.psi_best > best.psi_best || (q.psi_best == best.psi_best && q.zeta_arm < best.zeta_arm);
    }
    if (take) best = q;
  }
  best.tie_count = ties;
  best.best_arm = best.context >= 0 ? ext_arm[best.context] : -1;
  best.psi_best = best.psi_rest = 0.0;
  best.next_arm = -1;
  best.reserved[0] = bad;
  best.reserved[1] = best.reserved[2] = best.reserved[3] = 0;
  *out = best;
}

__global__ void arm_finish_kernel(crs_decision* decision, const double* __restrict__ psi, const int32_t* __restrict__ bad) {
  if (threadIdx.x != 0) return;
  decision->psi_best = psi[0];
  decision->psi_rest = psi[1];
  decision->next_arm = psi[0] < psi[1] ? decision->best_arm : decision->zeta_arm;
  if (bad && *bad != 0) decision->reserved[0] = *bad;
}

}  // namespace

int launch_arm_pack_best(const void* mean, const void* var, int64_t n_c, int64_t ld, int dtype, int arm_offset,
                         const int32_t* best_arm, double* pack, cudaStream_t s) {
  if (!mean || !var || !best_arm || !pack || n_c < 0) return CRS_ERR_INVALID_ARG;
  if (n_c == 0) return CRS_OK;
  const unsigned blocks = (unsigned)((n_c + 255) / 256);
  if (dtype == CRS_F64) arm_pack_best_kernel<double><<<blocks, 256, 0, s>>>((const double*)mean, (const double*)var, n_c, ld, arm_offset, best_arm, pack);
  else if (dtype == CRS_F32) arm_pack_best_kernel<float><<<blocks, 256, 0, s>>>((const float*)mean, (const float*)var, n_c, ld, arm_offset, best_arm, pack);
  else return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "arm_pack_best launch");
  return CRS_OK;
}

int launch_arm_combine_best(const double* pack_all, int world, int64_t n_c, int dtype, int32_t* ext_arm, void* ext_mean,
                            void* ext_var, cudaStream_t s) {
  if (!pack_all || world < 1 || !ext_arm || !ext_mean || !ext_var || n_c < 0) return CRS_ERR_INVALID_ARG;
  if (n_c == 0) return CRS_OK;
  const unsigned blocks = (unsigned)((n_c + 255) / 256);
  if (dtype == CRS_F64) arm_combine_best_kernel<double><<<blocks, 256, 0, s>>>(pack_all, world, n_c, ext_arm, (double*)ext_mean, (double*)ext_var);
  else if (dtype == CRS_F32) arm_combine_best_kernel<float><<<blocks, 256, 0, s>>>(pack_all, world, n_c, ext_arm, (float*)ext_mean, (float*)ext_var);
  else return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "arm_combine_best launch");
  return CRS_OK;
}

int launch_arm_record_mean(crs_decision* decision, const void* mean, int64_t ld, int dtype, int arm_offset, cudaStream_t s) {
  if (!decision || !mean) return CRS_ERR_INVALID_ARG;
  if (dtype == CRS_F64) arm_record_mean_kernel<double><<<1, 32, 0, s>>>(decision, (const double*)mean, ld, arm_offset);
  else if (dtype == CRS_F32) arm_record_mean_kernel<float><<<1, 32, 0, s>>>(decision, (const float*)mean, ld, arm_offset);
  else return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "arm_record_mean launch");
  return CRS_OK;
}

int launch_arm_fold(const crs_decision* records, int world, const int32_t* ext_arm, crs_decision* out, cudaStream_t s) {
  if (!records || world < 1 || !ext_arm || !out) return CRS_ERR_INVALID_ARG;
  arm_fold_kernel<<<1, 32, 0, s>>>(records, world, ext_arm, out);
  CRS_CUDA_TRY(cudaGetLastError(), "arm_fold launch");
  return CRS_OK;
}

int launch_arm_finish(crs_decision* decision, const double* psi, const int32_t* bad, cudaStream_t s) {
  if (!decision || !psi) return CRS_ERR_INVALID_ARG;
  arm_finish_kernel<<<1, 32, 0, s>>>(decision, psi, bad);
  CRS_CUDA_TRY(cudaGetLastError(), "arm_finish launch");
  return CRS_OK;
}

}  // namespace crs
