// Shared helpers for the PFN sm_100a kernels (error reporting, dtype traits, warp reductions).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace pfn {

// ---------------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry returns 0 on success, non-zero on failure and stores a
// message readable through pfn_last_error(). No C++ exception ever crosses the ABI.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define PFN_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::pfn::set_error(__VA_ARGS__);                               \
      return 1;                                                    \
    }                                                              \
  } while (0)

#define PFN_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::pfn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define PFN_LAUNCH_OK()                                                                \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ::pfn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 3;                                                                        \
    }                                                                                  \
  } while (0)

enum DType : int { kF32 = 0, kBF16 = 1 };

__host__ __device__ inline size_t dtype_size(int dt) { return dt == kF32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------------
// Device-side scalar conversion
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (the reference uses activation='gelu' => erf form, transformer.py:17)
__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float u) {
  const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

// Fast erf-GELU for the bf16 tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, far below bf16
// resolution) with one MUFU.RCP and one MUFU.EX2; exp(-u^2/2) is shared between erf and the Gaussian pdf of GELU'.
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// returns erf(|u| / sqrt(2)) and exp(-u^2 / 2)
__device__ __forceinline__ float erf_abs_as(float u, float& gauss) {
  const float ax = fabsf(u) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  gauss = fast_ex2(u * u * -0.72134752044448170f);        // exp(-u^2/2) = 2^(-u^2 * log2(e) / 2)
  return fmaf(-poly * t, gauss, 1.0f);
}
__device__ __forceinline__ float gelu_fast(float u) {
  float gs;
  const float e = erf_abs_as(u, gs);
  return 0.5f * u * (1.0f + copysignf(e, u));
}
__device__ __forceinline__ float gelu_grad_fast(float u) {
  float gs;
  const float e = erf_abs_as(u, gs);
  return fmaf(0.39894228040143268f * u, gs, 0.5f * (1.0f + copysignf(e, u)));
}

int num_sms();

}  // namespace pfn
