// Microbenchmark: what does each synchronisation primitive cost the MMA-issuing warp BETWEEN two tcgen05.mma batches?
// Pattern per block = the dQ kernel's: 16 TS N=64 score MMAs + commit, then 4 TS N=128 accumulate MMAs + commit.
// Variants insert, between the batches, the primitives the real kernel needs (all barriers are already complete, so any
// extra time is pure issue-side overhead that shows up as tensor-pipe idle time).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

template <int VAR>
__global__ void __launch_bounds__(320, 1) mma_gap(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, ready[4], dummy[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&ready[i], 1); tc::mbar_init(&dummy[i], 1); }
    tc::mbar_fence_init();
    stop = 0;
    for (int i = 0; i < 4; ++i) tc::mbar_arrive(&ready[i]);     // phase 0 of every `ready` barrier is complete
  }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t kv16 = tc::smem_u32(smem) >> 4;
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(128, 128, 0, 1);
    constexpr uint32_t hi = static_cast<uint32_t>(tc::umma_smem_desc_hi(1024) >> 32);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      int stage = 0;
      for (int r = 0; r < reps; ++r) {
        const uint32_t sb = (r & 1) * 64;
        const uint32_t k16 = kv16 + stage * 2048, v16 = k16 + 1024;
        // ---- before the score batch
        if (VAR == 2 || VAR == 4 || VAR == 5 || VAR >= 7) tc::mbar_wait(&ready[stage], 0);           // all 32 lanes poll a complete barrier
        if (VAR == 6) tc::mbar_wait_warp(&ready[stage], 0);                               // one lane polls
        if (VAR == 3 || VAR == 4 || VAR == 5 || VAR >= 7) tc::tc_fence_after();
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc::umma_bf16_ts(tmem + sb, tmem + 384 + kk * 8, (static_cast<uint64_t>(hi) << 32) | ((1u << 16) | (k16 + (kk >> 2) * 512 + (kk & 3) * 2)), idesc_s, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc::umma_bf16_ts(tmem + 128 + sb, tmem + 448 + kk * 8, (static_cast<uint64_t>(hi) << 32) | ((1u << 16) | (v16 + (kk >> 2) * 512 + (kk & 3) * 2)), idesc_s, kk > 0);
          if (VAR >= 1) tc::umma_commit(&dummy[0]);
        }
        __syncwarp();
        // ---- before the accumulate batch
        if (VAR == 2 || VAR == 4 || VAR == 5 || VAR >= 7) tc::mbar_wait(&ready[(stage + 1) & 3], 0);
        if (VAR == 6) tc::mbar_wait_warp(&ready[(stage + 1) & 3], 0);
        if (VAR == 3 || VAR == 4 || VAR == 5 || VAR >= 7) tc::tc_fence_after();
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc::umma_bf16_ts(tmem + 256, tmem + sb + (kk >> 1) * 32 + (kk & 1) * 8, (static_cast<uint64_t>(hi) << 32) | ((512u << 16) | (k16 + kk * 128)), idesc_o, 1u);
          if (VAR >= 1) { tc::umma_commit(&dummy[1]); }
          if (VAR == 5 || VAR >= 7) tc::umma_commit(&dummy[2]);
        }
        __syncwarp();
        if (++stage == 4) stage = 0;
      }
      if (tc::elect_one()) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if (lane == 0) { out[0] = t1 - t0; stop = 1; }
  } else if (warp >= 2 && (VAR >= 7)) {
    // the row warps of the real kernel: 2 x tcgen05.ld x32 + (VAR 8: 32 ex2 per thread) + tcgen05.st x16 per iteration
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int half = (warp - 2) >> 2;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t a[32], b[32], pk[16];
      tc::tmem_ld_32x32b_x32(tmem + lane_off + half * 32, a);
      tc::tmem_ld_32x32b_x32(tmem + lane_off + 128 + half * 32, b);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (VAR >= 8) {
          const float p0 = tc::fast_exp2(fmaf(__uint_as_float(a[2 * c]), 0.1f, -3.f)), p1 = tc::fast_exp2(fmaf(__uint_as_float(a[2 * c + 1]), 0.1f, -3.f));
          pk[c] = tc::pack_bf16x2(p0 * fmaf(__uint_as_float(b[2 * c]), 0.1f, -1.f), p1 * fmaf(__uint_as_float(b[2 * c + 1]), 0.1f, -1.f));
        } else pk[c] = a[2 * c] ^ b[2 * c + 1];
      }
      tc::tmem_st_32x32b_x16(tmem + lane_off + 320 + half * 16, pk);
      tc::tmem_st_wait();
      acc ^= pk[3];
    }
    if (acc == 0x12345) out[1] = acc;
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int VAR>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 400;
  auto k = mma_gap<VAR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 2048);
  k<<<148, 320, 4 * 32768 + 2048>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-78s: %7.0f clk/block (pipe floor 768) [%s]\n", name, double(h) / reps, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<0>("0 runtime-stage descriptors, no commits, nothing between batches");
  run<1>("1 + one tcgen05.commit after each batch");
  run<2>("2 + all-lane poll of a COMPLETE mbarrier before each batch");
  run<3>("3 commits + tcgen05.fence::after_thread_sync before each batch (no polls)");
  run<4>("4 commits + poll + fence before each batch (what the dQ kernel does)");
  run<5>("5 = 4 + a second commit after the accumulate batch");
  run<6>("6 commits + ONE-lane poll (mbar_wait_warp) before each batch");
  run<7>("7 = 5 + 8 row warps looping tcgen05.ld x32 x2 / st x16 on other TMEM columns");
  run<8>("8 = 7 + the row arithmetic (32 ex2 + fma + pack per thread and iteration)");
  return 0;
}
