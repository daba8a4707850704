// Elementwise dropout (+ residual) with regenerated counter-based masks, and the mask dump used by the tests' oracle.
// HBM-bound: one read of x (and of the residual), one write; 8 elements (two hash evaluations) per thread.
#include "common.cuh"
#include "dropout.cuh"
#include "../../include/pfn_b200.h"

namespace pfn {

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ res, int ldr, T* __restrict__ out, int ldo, int rows,
               int cols, uint32_t seed, int thr) {
  const int groups = cols >> 3;                       // 8-element groups per row
  const long long total = static_cast<long long>(rows) * groups;
  const float sc = drop_scale(thr);
  for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(g / groups);
    const int c0 = static_cast<int>(g - static_cast<long long>(r) * groups) << 3;
    float v[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = to_f32<T>(x[static_cast<size_t>(r) * ldx + c0 + i]);
    if (res != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = to_f32<T>(res[static_cast<size_t>(r) * ldr + c0 + i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.f;
    }
    const uint32_t h0 = drop_hash(seed, static_cast<uint32_t>(r), static_cast<uint32_t>(c0 >> 2));
    const uint32_t h1 = drop_hash(seed, static_cast<uint32_t>(r), static_cast<uint32_t>(c0 >> 2) + 1u);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool keep = drop_keep_byte(i < 4 ? h0 : h1, i & 3, thr);
      out[static_cast<size_t>(r) * ldo + c0 + i] = from_f32<T>((keep ? v[i] * sc : 0.f) + a[i]);
    }
  }
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* __restrict__ out, int rows, int cols, uint32_t seed, int thr) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t r = static_cast<uint32_t>(e / cols), c = static_cast<uint32_t>(e % cols);
    out[e] = drop_keep(seed, r, c, thr) ? 1 : 0;
  }
}

}  // namespace pfn

using namespace pfn;

extern "C" int pfn_dropout(const void* x, int ldx, const void* residual, int ldr, void* out, int ldo, int rows, int cols,
                           int dtype, uint32_t seed, int thr, void* stream) {
  PFN_CHECK_ARG(x != nullptr && out != nullptr && rows > 0 && cols > 0, "dropout: bad arguments");
  PFN_CHECK_ARG(cols % 8 == 0, "dropout: cols %d must be a multiple of 8", cols);
  PFN_CHECK_ARG(thr >= 0 && thr <= 255, "dropout: threshold %d outside [0,255]", thr);
  PFN_CHECK_ARG(dtype == PFN_F32 || dtype == PFN_BF16, "dropout: bad dtype %d", dtype);
  const long long total = static_cast<long long>(rows) * (cols / 8);
  long long grid = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (grid > cap) grid = cap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == PFN_F32)
    dropout_kernel<float><<<static_cast<int>(grid), 256, 0, s>>>(static_cast<const float*>(x), ldx, static_cast<const float*>(residual), ldr,
                                                             static_cast<float*>(out), ldo, rows, cols, seed, thr);
  else
    dropout_kernel<__nv_bfloat16><<<static_cast<int>(grid), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                                     static_cast<const __nv_bfloat16*>(residual), ldr,
                                                                     static_cast<__nv_bfloat16*>(out), ldo, rows, cols, seed, thr);
  PFN_LAUNCH_OK();
  return 0;
}

extern "C" int pfn_dropout_keep_mask(uint8_t* out, int rows, int cols, uint32_t seed, int thr, void* stream) {
  PFN_CHECK_ARG(out != nullptr && rows > 0 && cols > 0, "dropout_keep_mask: bad arguments");
  PFN_CHECK_ARG(thr >= 0 && thr <= 255, "dropout_keep_mask: threshold %d outside [0,255]", thr);
  const long long total = static_cast<long long>(rows) * cols;
  long long grid = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (grid > cap) grid = cap;
  dropout_mask_kernel<<<static_cast<int>(grid), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, rows, cols, seed, thr);
  PFN_LAUNCH_OK();
  return 0;
}
