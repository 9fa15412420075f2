// Peer mailboxes: the device-side exchange of one 64-byte record per rank between the GPUs of one
// NVLink / NVSwitch domain (one process per GPU), used where the decision path has a real exchange
// step (the per-rank local minimiser of a sharded decision, the PCS count totals).
//
// Every rank owns a mailbox of 2 x world slots of 128 bytes in ITS OWN memory, peer-mapped by all
// ranks (CUDA IPC, crs_ipc_export / crs_ipc_open).  In exchange number `seq` the posting CTA stores its
// record straight into slot [seq & 1][rank] of EVERY rank's mailbox (a 64-byte store over NVLink per
// peer), flags each with `seq` (release, system scope), then waits until the world slots of its own
// mailbox carry `seq` (acquire).  No NCCL launch and no host in the loop; NVLink traffic per exchange
// and rank is (world - 1) x 68 bytes.  The double buffer is enough because a rank cannot post
// exchange s + 2 before every peer posted s + 1, i.e. finished reading s.  Ranks are on different
// GPUs, so the wait cannot starve the peer it waits for; a peer that never shows up ends the wait after
// ~2 s with a time-out flag instead of hanging the GPU.
#pragma once
#include "crs_common.cuh"

namespace crs {
namespace mailbox {

constexpr int kSlotBytes = 128;   // 64-byte record | int32 flag at +64 | pad
constexpr int kMaxWorld = 64;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ const unsigned char* my_slot(const crs_peer_exchange& px, int r) {
  return reinterpret_cast<const unsigned char*>(px.mailboxes[px.rank]) +
         ((size_t)(px.seq & 1) * (size_t)px.world + (size_t)r) * kSlotBytes;
}

// Called by threads 0 .. world-1 of ONE CTA (thread t serves peer t): post `rec` (64 bytes, 16-byte
// aligned, in registers/local memory of the caller) to peer t and wait for peer t's record in the own
// mailbox.  Returns false on time-out.
__device__ __forceinline__ bool post_and_wait(const crs_peer_exchange& px, int t, const int4 rec[4]) {
  unsigned char* slot = reinterpret_cast<unsigned char*>(px.mailboxes[t]) +
                        ((size_t)(px.seq & 1) * (size_t)px.world + (size_t)px.rank) * kSlotBytes;
  int4* dst = reinterpret_cast<int4*>(slot);
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[i] = rec[i];
  __threadfence_system();
  st_release_sys(reinterpret_cast<int*>(slot + 64), px.seq);
  const int* flag = reinterpret_cast<const int*>(my_slot(px, t) + 64);
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) != px.seq) {
    __nanosleep(32);
    if (global_ns() - t0 > 2000000000ull) return false;
  }
  return true;
}

__device__ __forceinline__ void read_slot(const crs_peer_exchange& px, int r, int4 out[4]) {
  const int4* src = reinterpret_cast<const int4*>(my_slot(px, r));
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = __ldcv(src + i);
}

inline bool valid(const crs_peer_exchange* px) {
  return px && px->world >= 1 && px->world <= kMaxWorld && px->rank >= 0 && px->rank < px->world && px->seq != 0 &&
         (px->world == 1 || px->mailboxes != nullptr);
}

}  // namespace mailbox
}  // namespace crs
