// Microbenchmark (round 2): tensor-pipe clocks per key block when the RESIDENT operand of the score MMAs (Q / dO in the dQ
// kernel, K / V in the dK,dV kernel, Q in the forward) is read from TMEM (TS mode) instead of shared memory (SS mode).
// SS reads A (128 x 16 bf16 = 4 KB) + B per MMA from smem; TS only B.  No data dependencies: pure pipe occupancy.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

// MODE 0: SS scores   MODE 1: TS scores (A in TMEM columns 448..511 region, packed bf16)
// NS = keys per block (64 / 128); NSCORE = score GEMMs per block (1 fwd, 2 dQ); NACC = accumulate GEMMs (1 fwd/dQ, 2 dKV)
template <int MODE, int NS, int NSCORE, int NACC>
__global__ void __launch_bounds__(320, 1) attn_pattern2(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp >= 2 && warp < 6) {       // fill the TMEM A region with something
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = 0x3c003c00u;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    for (int c = 0; c < 512; c += 32) tc::tmem_st_32x32b_x32(tmem + lane_off + c, v);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  if (warp == 1) {
    const uint32_t q_addr = tc::smem_u32(smem), k_addr = tc::smem_u32(smem + 32768), v_addr = tc::smem_u32(smem + 65536);
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(128, NS, 0, 0);
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(128, 128, 0, 1);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      if (tc::elect_one()) {
        for (int r = 0; r < reps; ++r) {
          const uint32_t sb = (r & 1) * NS;          // S double buffer at columns 0 / NS (<= 256 used)
#pragma unroll
          for (int sc = 0; sc < NSCORE; ++sc) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t bdesc = tc::umma_smem_desc((sc ? v_addr : k_addr) + (kk >> 2) * (NS * 128) + (kk & 3) * 32, 16, 1024);
              if (MODE == 0)
                tc::umma_bf16_ss(tmem + sc * 128 + (sb & 127), tc::umma_smem_desc(q_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), bdesc, idesc_s, kk > 0);
              else
                tc::umma_bf16_ts(tmem + sc * 128 + (sb & 127), tmem + 384 + sc * 64 + kk * 8, bdesc, idesc_s, kk > 0);
            }
          }
#pragma unroll
          for (int ac = 0; ac < NACC; ++ac) {
#pragma unroll
            for (int kk = 0; kk < NS / 16; ++kk)
              tc::umma_bf16_ts(tmem + 256 + (ac & 0) * 0, tmem + (sb & 127) + kk * 8, tc::umma_smem_desc((ac ? v_addr : k_addr) + (kk & 3) * 2048 + (kk >> 2) * 16384 * 0, 8192, 1024), idesc_o, 1u);
          }
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if (lane == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int MODE, int NS, int NSCORE, int NACC>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 200;
  auto k = attn_pattern2<MODE, NS, NSCORE, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 2048);
  k<<<148, 320, 3 * 32768 + 2048>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double flop = 2.0 * 128 * NS * 128 * (NSCORE + NACC);
  printf("%-54s %s scores: %7.0f clk/block  = %5.0f flop/clk/SM  (%.2f of 8192) [%s]\n", name, MODE ? "TS" : "SS", double(h) / reps,
         flop / (double(h) / reps), flop / (double(h) / reps) / 8192.0, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<0, 64, 1, 1>("fwd  64-key: 8 score N=64  + 4 acc N=128");
  run<1, 64, 1, 1>("fwd  64-key: 8 score N=64  + 4 acc N=128");
  run<0, 128, 1, 1>("fwd 128-key: 8 score N=128 + 8 acc N=128");
  run<1, 128, 1, 1>("fwd 128-key: 8 score N=128 + 8 acc N=128");
  run<0, 64, 2, 1>("dQ   64-key: 16 score N=64  + 4 acc N=128");
  run<1, 64, 2, 1>("dQ   64-key: 16 score N=64  + 4 acc N=128");
  run<0, 128, 2, 1>("dQ  128-key: 16 score N=128 + 8 acc N=128");
  run<1, 128, 2, 1>("dQ  128-key: 16 score N=128 + 8 acc N=128");
  run<0, 64, 2, 2>("dKV  64-row: 16 score N=64  + 8 acc N=128");
  run<1, 64, 2, 2>("dKV  64-row: 16 score N=64  + 8 acc N=128");
  run<0, 128, 2, 2>("dKV 128-row: 16 score N=128 + 16 acc N=128");
  run<1, 128, 2, 2>("dKV 128-row: 16 score N=128 + 16 acc N=128");
  return 0;
}
