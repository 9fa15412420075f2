// Philox4x32-10 and the uniform / normal variate helpers shared by the PCS kernels
// (pcs_philox.cu: flat per-arm stream; pcs_tree.cu: max-tree stream).  Stream definitions:
// oracle/philox_ref.py, oracle/pcs_tree_ref.py.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace crs {
namespace rng {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct Words { uint32_t x, y, z, w; };

// round keys k + r * W, computed on the host and read from the kernel-parameter constant bank:
// the key schedule costs no instruction
struct RoundKeys { uint32_t k0[10], k1[10]; };

__host__ __device__ inline RoundKeys round_keys(uint32_t k0, uint32_t k1) {
  RoundKeys rk;
  for (int r = 0; r < 10; ++r) {
    rk.k0[r] = k0 + (uint32_t)r * kPhiloxW0;
    rk.k1[r] = k1 + (uint32_t)r * kPhiloxW1;
  }
  return rk;
}

__device__ __forceinline__ Words philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const RoundKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ rk.k0[r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ rk.k1[r];
    c3 = lo0;
  }
  return Words{c0, c1, c2, c3};
}

__device__ __forceinline__ float approx_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ uint32_t mantissa_bits(uint32_t w) {
  uint32_t r;
  asm("mad.hi.u32 %0, %1, 0x800000, 0x3f800000;" : "=r"(r) : "r"(w));
  return r;
}

__device__ __forceinline__ float approx_log2(float x) {  // x >= 2^-24: never subnormal
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// (wa, wb) -> two N(0,1) draws; see oracle/philox_ref.py for the stream definition.
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1) {
  // (1 + m 2^-23) - (1 - 2^-24) = m 2^-23 + 2^-24 exactly: one FADD builds the open-interval uniform
  // 0x3f800000 | (w >> 9) as one multiply-add (hi32(w * 2^23) + 0x3f800000): runs on the FMA pipe and
  // keeps the ALU pipe — the busiest one of this kernel — for the Philox XORs
  const float a = __uint_as_float(mantissa_bits(wa)) - 0.99999994039535522f;
  const float b = __uint_as_float(mantissa_bits(wb)) - 1.0f;
  const float rad = approx_sqrt(-1.3862943611198906f * approx_log2(a));  // sqrt(-2 ln a)
  float sn, cs;
  __sincosf(6.2831853071795865f * b, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

// the same pair in parts: z0 = rad * cs, z1 = rad * sn (lets a caller fold a scale into the radius)
__device__ __forceinline__ void box_muller_parts(uint32_t wa, uint32_t wb, float& rad, float& cs, float& sn) {
  const float a = __uint_as_float(mantissa_bits(wa)) - 0.99999994039535522f;
  const float b = __uint_as_float(mantissa_bits(wb)) - 1.0f;
  rad = approx_sqrt(-1.3862943611198906f * approx_log2(a));
  __sincosf(6.2831853071795865f * b, &sn, &cs);
}

// log2 of the open-interval uniform u(w) = ((w >> 9) + 1/2) 2^-23; -ln 2 times this is an Exp(1) variate
__device__ __forceinline__ float log2_unit_open(uint32_t w) {
  return approx_log2(__uint_as_float(mantissa_bits(w)) - 0.99999994039535522f);
}

// first (cosine) Box-Muller variate of (wa, wb) — identical to z0 of box_muller
__device__ __forceinline__ float box_muller_cos(uint32_t wa, uint32_t wb) {
  const float b = __uint_as_float(mantissa_bits(wb)) - 1.0f;
  const float rad = approx_sqrt(-1.3862943611198906f * log2_unit_open(wa));
  return rad * __cosf(6.2831853071795865f * b);
}

}  // namespace rng
}  // namespace crs
