// crs_peak_probe: micro-kernels that measure the pipe peaks the roofline fractions are quoted
// against (MEASURED_PEAKS.json has no fp64 / fp32-FMA / Philox figure): kind 0 = FP64 FMA,
// 1 = FP32 FMA, 2 = FP64 mma.sync m8n8k4, 5 = Philox4x32-10 + Box-Muller + FMA + max (the bare PCS
// draw sequence, pcs_philox.cu).
// Each thread runs `iters` rounds of 16 independent accumulator chains (kinds 0/1).
#include "crs_common.cuh"

namespace crs {
namespace {

// OP: 0 = fma, 1 = add, 2 = mul
template <typename T, int OP>
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, T seed, T* out) {
  T a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + T(i) * T(1e-3) + T(threadIdx.x) * T(1e-6);
  const T m = T(0.999999), b = T(1e-7);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = OP == 0 ? fma(a[i], m, b) : (OP == 1 ? a[i] + b : a[i] * m);
    }
  }
  T s = T(0);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == T(-1)) out[0] = s;  // keep the chains alive
}

// FP64 tensor path: mma.sync.m8n8k4.f64 (SASS DMMA), 8 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double seed, double* out) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = seed * i; c[i][1] = seed + i; }
  const double a = 0.999999 + threadIdx.x * 1e-9, b = 1e-7 + threadIdx.x * 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == -1.0) out[0] = s;
}

// kind 10: even warps run the DFMA chains, odd warps the DMMA tiles, same iteration counts as kinds 0 / 2:
// if the launch takes as long as the slower of the two alone, the FP64 FMA and FP64 MMA paths are separate
// pipes; if it takes the sum, they share the units.
__global__ void __launch_bounds__(256) dfma_dmma_mix_kernel(int iters, double seed, double* out) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = seed * i; c[i][1] = seed + i; }
    const double a = 0.999999 + threadIdx.x * 1e-9, b = 1e-7 + threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  } else {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i * 1e-3 + threadIdx.x * 1e-6;
    const double m = 0.999999, b = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, b);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
  }
  if (s == -1.0) out[0] = s;
}

}  // namespace

int launch_peak_probe(int kind, int iters, int blocks, void* out, cudaStream_t s) {
  if (!out || iters < 1 || blocks < 1) return CRS_ERR_INVALID_ARG;
  if (kind == 0) fma_probe_kernel<double, 0><<<blocks, 256, 0, s>>>(iters, 1.0, (double*)out);
  else if (kind == 1) fma_probe_kernel<float, 0><<<blocks, 256, 0, s>>>(iters, 1.0f, (float*)out);
  else if (kind == 3) fma_probe_kernel<double, 1><<<blocks, 256, 0, s>>>(iters, 1.0, (double*)out);
  else if (kind == 4) fma_probe_kernel<double, 2><<<blocks, 256, 0, s>>>(iters, 1.0, (double*)out);
  else if (kind == 2) dmma_probe_kernel<<<blocks, 256, 0, s>>>(iters, 1.0, (double*)out);  // 16 DMMA / iter / warp
  else if (kind == 10) dfma_dmma_mix_kernel<<<blocks, 256, 0, s>>>(iters, 1.0, (double*)out);
  else if (kind == 5) return launch_pcs_probe(iters, blocks, out, s);  // Philox + Box-Muller + FMA + max
  else if (kind == 6 || kind == 7) return launch_tree_probe(kind, iters, blocks, out, s);  // max-tree PCS: root test / expansion
  else if (kind == 8 || kind == 9) return launch_cdf_probe(kind, iters, blocks, out, s);  // cdf PCS: sampling call / 4 table terms
  else return CRS_ERR_INVALID_ARG;
  CRS_CUDA_TRY(cudaGetLastError(), "peak_probe launch");
  return CRS_OK;
}

}  // namespace crs
