// Microbenchmark: fp32 FMA issue rate on sm_100a -- scalar FFMA (three register operands) against the packed
// fma.rn.f32x2 (FFMA2, two FMAs per lane and instruction).  The GP sampler's update product and the GELU / softmax
// epilogues are made of these.  Prints FMA/clk/SM for 4, 8 and 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\t"
               "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
               "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
               "mov.b64 {%0, %1}, rc;\n\t}"
               : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

template <int KIND>
__global__ void __launch_bounds__(512, 1) k(long long* out, float seed, int reps) {
  float2 acc[8];
  float2 a[2], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(seed * i, seed + i);
  a[0] = make_float2(seed, seed * 0.5f); a[1] = make_float2(seed * 0.25f, seed * 2.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = make_float2(seed + j, seed + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (KIND == 0) {                      // 16 scalar FFMA: 4x4 micro-tile step
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i * 4 + j].x = fmaf(a[i].x, b[j].x, acc[i * 4 + j].x);
            acc[i * 4 + j].y = fmaf(a[i].y, b[j].x, acc[i * 4 + j].y);
          }
      } else {                              // 8 FFMA2: the same 16 FMAs
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) ffma2(acc[i * 4 + j], a[i], b[j]);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  if (s == 12345.f) out[1] = 1;
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 20000;
  for (int kind = 0; kind < 2; ++kind)
    for (int warps : {4, 8, 16}) {
      if (kind == 0) k<0><<<148, warps * 32>>>(d, 1.0001f, reps); else k<1><<<148, warps * 32>>>(d, 1.0001f, reps);
      cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double fma_per_clk_sm = double(reps) * 4 * 16 * warps * 32 / double(h);
      printf("%s, %2d warps/SM: %.1f FMA/clk/SM\n", kind == 0 ? "FFMA  (scalar)" : "FFMA2 (f32x2) ", warps, fma_per_clk_sm);
    }
  return 0;
}
