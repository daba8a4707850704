// Microbenchmark: clocks per tcgen05.mma (kind::f16, M=128) for several N, SS (A,B smem) and TS (A tmem), back to back.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

template <int N, bool TS, bool BMN>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t a_addr = tc::smem_u32(smem), b_addr = tc::smem_u32(smem + 16384);
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, N, 0, BMN ? 1 : 0);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      if (tc::elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a_desc = tc::umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t b_desc = BMN ? tc::umma_smem_desc(b_addr + k * 2048, 8192, 1024) : tc::umma_smem_desc(b_addr + k * 32, 16, 1024);
            if (TS) tc::umma_bf16_ts(tmem, tmem + 256 + k * 8, b_desc, idesc, 1u);
            else tc::umma_bf16_ss(tmem, a_desc, b_desc, idesc, 1u);
          }
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int N, bool TS, bool BMN>
void run(const char* name, int nctas) {
  long long* d; cudaMalloc(&d, 8);
  const int reps = 256;
  auto k = mma_rate<N, TS, BMN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k<<<nctas, 128, 60000>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s ctas=%3d : %.1f clk/MMA (ideal %d)  [%s]\n", name, nctas, double(h) / (reps * 4), N / 2, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int nctas : {1, 148}) {
    run<64, false, false>("SS N=64  (B K-major)", nctas);
    run<128, false, false>("SS N=128 (B K-major)", nctas);
    run<256, false, false>("SS N=256 (B K-major)", nctas);
    run<128, false, true>("SS N=128 (B MN-major)", nctas);
    run<64, true, true>("TS N=64  (B MN-major)", nctas);
    run<128, true, true>("TS N=128 (B MN-major)", nctas);
    run<256, true, true>("TS N=256 (B MN-major)", nctas);
    run<128, true, false>("TS N=128 (B K-major)", nctas);
  }
  return 0;
}
