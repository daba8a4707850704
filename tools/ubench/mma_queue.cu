// Microbenchmark: how deep is the tcgen05.mma issue queue, and what does an idle gap between MMA groups cost?
//  (a) issue n SS N=64 MMAs on an empty pipe: clocks until the issuing thread is past the last MMA vs until the commit lands
//  (b) steady state of [16 SS N=64][gap][4 TS N=128][gap] with a software delay of `gap` clocks in the issuing warp
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

__device__ __forceinline__ void spin_clocks(int n) {
  const long long t = clock64();
  while (clock64() - t < n) {}
}

__global__ void __launch_bounds__(128, 1) mma_queue(long long* out, int n_mma, int gap, int reps, int boundary) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t ready_bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&ready_bar, 1); tc::mbar_fence_init(); tc::mbar_arrive(&ready_bar); }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t q_addr = tc::smem_u32(smem), k_addr = tc::smem_u32(smem + 32768);
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(128, 128, 0, 1);
    uint32_t phase = 0;
    if (n_mma > 0) {
      long long t0 = 0, t1 = 0, t2 = 0;
      for (int pass = 0; pass < 3; ++pass) {
        t0 = clock64();
        if (tc::elect_one()) {
          for (int i = 0; i < n_mma; ++i)
            tc::umma_bf16_ss(tmem + (i & 1) * 64, tc::umma_smem_desc(q_addr + (i & 3) * 32, 16, 1024),
                             tc::umma_smem_desc(k_addr + (i & 3) * 32, 16, 1024), idesc_s, 1u);
          t1 = clock64();
          tc::umma_commit(&bar);
        }
        __syncwarp();
        tc::mbar_wait(&bar, phase & 1); ++phase;
        t2 = clock64();
      }
      t1 = __shfl_sync(0xffffffffu, t1, 0);   // elected lane is lane 0 in practice
      if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)
            tc::umma_bf16_ss(tmem + (kk >> 3) * 64, tc::umma_smem_desc(q_addr + ((kk >> 2) & 1) * 16384 + (kk & 3) * 32, 16, 1024),
                             tc::umma_smem_desc(k_addr + ((kk >> 2) & 1) * 8192 + (kk & 3) * 32, 16, 1024), idesc_s, (kk & 7) > 0);
        }
        __syncwarp();
        if (gap) spin_clocks(gap);
        if (boundary == 1 || boundary >= 3) { if (boundary >= 2) tc::mbar_wait(&ready_bar, 0); tc::tc_fence_after(); }
        else if (boundary == 2) tc::mbar_wait(&ready_bar, 0);
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc::umma_bf16_ts(tmem + 384, tmem + 128 + kk * 8, tc::umma_smem_desc(k_addr + kk * 2048, 8192, 1024), idesc_o, 1u);
        }
        __syncwarp();
        if (gap) spin_clocks(gap);
        if (boundary == 1 || boundary >= 3) { if (boundary >= 2) tc::mbar_wait(&ready_bar, 0); tc::tc_fence_after(); }
        else if (boundary == 2) tc::mbar_wait(&ready_bar, 0);
      }
      if (tc::elect_one()) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, 0);
      long long t2 = clock64();
      if (lane == 0) { out[0] = 0; out[1] = (t2 - t0) / reps; }
    }
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 2048);
  for (int n : {1, 2, 3, 4, 6, 8, 12, 16, 24, 32}) {
    mma_queue<<<148, 128, 3 * 32768 + 2048>>>(d, n, 0, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("n=%2d SS N=64 on an empty pipe: issue done after %5lld clk, commit seen after %5lld clk [%s]\n", n, h[0], h[1], cudaGetErrorString(e));
  }
  for (int gap : {0, 50, 100, 200, 400}) {
    mma_queue<<<148, 128, 3 * 32768 + 2048>>>(d, 0, gap, 200, 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("[16 SS N=64][gap %3d][4 TS N=128][gap %3d]: %lld clk/block [%s]\n", gap, gap, h[1], cudaGetErrorString(e));
  }
  const char* names[] = {"none", "tcgen05.fence::after_thread_sync", "mbar_wait(ready)", "mbar_wait(ready)+fence"};
  for (int b = 0; b < 4; ++b) {
    mma_queue<<<148, 128, 3 * 32768 + 2048>>>(d, 0, 0, 200, b);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("[16 SS N=64][B][4 TS N=128][B], B = %-36s: %lld clk/block [%s]\n", names[b], h[1], cudaGetErrorString(e));
  }
  return 0;
}
