// Microbenchmark: does concurrent bulk-copy (TMA engine) traffic INTO shared memory slow tcgen05.mma operand reads?
//  warp 1 issues a long stream of SS MMAs (GEMM shape: M=128 N=256 K=16, or attention score shape N=64);
//  warp 0 (optional) keeps `depth` 32 KB cp.async.bulk global->shared copies in flight into other smem regions.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_vs_tma(long long* out, const uint8_t* __restrict__ src, size_t src_bytes, int reps,
                                                     int depth, int chunk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, cbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) tc::mbar_init(&cbar[i], 1); tc::mbar_fence_init(); stop = 0; }
  if (warp == 2) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t a_addr = tc::smem_u32(smem), b_addr = tc::smem_u32(smem + 16384);
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    if (tc::elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_bf16_ss(tmem + (r & 1) * 256, tc::umma_smem_desc(a_addr + k * 32, 16, 1024), tc::umma_smem_desc(b_addr + k * 32, 16, 1024), idesc, 1u);
      }
      tc::umma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; stop = 1; }
  } else if (warp == 0 && depth > 0) {
    // copy stream: `depth` copies of `chunk` bytes in flight, destinations above the operand tiles
    uint32_t issued = 0, waited = 0;
    long long bytes = 0;
    size_t off = (static_cast<size_t>(blockIdx.x) * 262144) % (src_bytes - 4 * 65536);
    while (!stop) {
      if (issued - waited < static_cast<uint32_t>(depth)) {
        const int s = issued % depth;
        if (lane == 0) {
          tc::mbar_expect_tx(&cbar[s], chunk);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           tc::smem_u32(smem + 49152 + s * chunk)),
                       "l"(src + off), "r"(chunk), "r"(tc::smem_u32(&cbar[s]))
                       : "memory");
        }
        off += chunk; if (off + chunk > src_bytes) off = 0;
        ++issued;
      } else {
        const int s = waited % depth;
        tc::mbar_wait(&cbar[s], (waited / depth) & 1);
        ++waited; bytes += chunk;
      }
    }
    while (waited < issued) { const int s = waited % depth; tc::mbar_wait(&cbar[s], (waited / depth) & 1); ++waited; bytes += chunk; }
    if (lane == 0) out[blockIdx.x * 2 + 1] = bytes;
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int N>
void run(const char* name, const uint8_t* src, size_t src_bytes) {
  long long* d; cudaMalloc(&d, 148 * 16);
  auto k = mma_vs_tma<N>;
  const int smem = 49152 + 4 * 32768 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2000;
  for (int depth : {0, 1, 2, 4}) {
    for (int chunk : {16384, 32768}) {
      if (depth == 0 && chunk != 16384) continue;
      cudaMemset(d, 0, 148 * 16);
      k<<<148, 128, smem>>>(d, src, src_bytes, reps, depth, chunk);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-22s copies in flight %d x %2d KB: %.1f clk/MMA (ideal %d), copy stream %.1f B/clk/SM [%s]\n", name, depth, chunk / 1024,
             double(h[0]) / (reps * 4), N / 2, double(h[1]) / double(h[0]), cudaGetErrorString(e));
    }
  }
  cudaFree(d);
}
int main() {
  uint8_t* src; const size_t bytes = size_t(48) << 20;   // L2-resident source: the copy stream runs at L2 speed, like the re-used GEMM operand tiles
  cudaMalloc(&src, bytes); cudaMemset(src, 0x3c, bytes);
  run<256>("SS M=128 N=256 (GEMM)", src, bytes);
  run<64>("SS M=128 N=64 (scores)", src, bytes);
  return 0;
}
