// crs_peer_sum_u64: sum up to 8 unsigned 64-bit counters over the ranks of a peer-mailbox group —
// the "one allreduce" that combines the PCS success counts of a context-sharded estimate
// (contextual_rs/generalized_pcs.py:470-479 with rho = mean) — as ONE single-CTA
// launch on the caller's stream: post the 64-byte record to every peer over NVLink, wait for theirs,
// add, and publish the result to device memory and (optionally) mapped pinned host memory with a
// completion tag written last (crs_wait_u64 polls it; no stream synchronise, no NCCL launch).
#include <time.h>
#include "crs_common.cuh"
#include "peer_mailbox.cuh"

namespace crs {
namespace {

__global__ void __launch_bounds__(64)
peer_sum_kernel(const unsigned long long* __restrict__ values, int n, unsigned long long* __restrict__ out,
                unsigned long long* __restrict__ host, crs_peer_exchange px) {
  __shared__ int s_timeout;
  const int t = threadIdx.x;
  if (t == 0) s_timeout = 0;
  __syncthreads();
  if (px.world > 1 && t < px.world) {
    alignas(16) unsigned long long rec[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rec[i] = i < n ? values[i] : 0ull;
    if (!mailbox::post_and_wait(px, t, reinterpret_cast<const int4*>(rec))) atomicExch(&s_timeout, 1);
  }
  __syncthreads();
  if (t == 0) {
    unsigned long long sum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = (px.world > 1 || i >= n) ? 0ull : values[i];
    if (px.world > 1) {
      for (int r = 0; r < px.world; ++r) {
        alignas(16) unsigned long long q[8];
        mailbox::read_slot(px, r, reinterpret_cast<int4*>(q));
#pragma unroll
        for (int i = 0; i < 8; ++i) sum[i] += q[i];
      }
    }
    if (out)
      for (int i = 0; i < n; ++i) out[i] = sum[i];
    if (host) {
#pragma unroll
      for (int i = 0; i < 8; ++i) host[i] = sum[i];
      host[9] = s_timeout ? 1ull : 0ull;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(host + 8) = (unsigned long long)(unsigned int)px.seq;
    }
  }
}

}  // namespace

int launch_peer_sum(const uint64_t* values, int n, uint64_t* out, uint64_t* host_mirror, const crs_peer_exchange* px,
                    cudaStream_t s) {
  if (!values || n < 1 || n > 8 || !mailbox::valid(px)) return CRS_ERR_INVALID_ARG;
  peer_sum_kernel<<<1, 64, 0, s>>>(reinterpret_cast<const unsigned long long*>(values), n,
                                   reinterpret_cast<unsigned long long*>(out),
                                   reinterpret_cast<unsigned long long*>(host_mirror), *px);
  CRS_CUDA_TRY(cudaGetLastError(), "peer_sum launch");
  return CRS_OK;
}

}  // namespace crs
