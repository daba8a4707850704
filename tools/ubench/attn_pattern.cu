// Microbenchmark: tensor-pipe clocks per key block for the MMA sequences of the attention kernels, with and without
// elementwise warps hammering TMEM (tcgen05.ld / st) at the same time.  No data dependencies: pure pipe occupancy.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

// pattern 0: dQ kernel, 64-key blocks : S(8 SS N=64) + dP(8 SS N=64) + dQ(4 TS N=128)
// pattern 1: fwd kernel, 64-key blocks: S(8 SS N=64) + PV(4 TS N=128)
// pattern 2: dQ kernel, 128-key blocks: S(8 SS N=128) + dP(8 SS N=128) + dQ(8 TS N=128)
// pattern 3: fwd kernel, 128-key blocks: S(8 SS N=128) + PV(8 TS N=128)
// pattern 4: dKV kernel, 64-row blocks : ST(8 SS N=64) + dPT(8 SS N=64) + dV(4 TS N=128) + dK(4 TS N=128)
template <int PATTERN>
__global__ void __launch_bounds__(320, 1) attn_pattern(long long* out, int reps, int contend, int commits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (3 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) tc::mbar_init(&dummy[i], 1); tc::mbar_fence_init(); stop = 0; }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t q_addr = tc::smem_u32(smem), k_addr = tc::smem_u32(smem + 32768), v_addr = tc::smem_u32(smem + 65536);
    constexpr int NS = (PATTERN == 2 || PATTERN == 3) ? 128 : 64;
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(128, NS, 0, 0);
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(128, 128, 0, 1);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      if (tc::elect_one()) {
        for (int r = 0; r < reps; ++r) {
          const uint32_t sb = (r & 1) * NS;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc::umma_bf16_ss(tmem + sb, tc::umma_smem_desc(q_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                             tc::umma_smem_desc(k_addr + (kk >> 2) * (NS * 128) + (kk & 3) * 32, 16, 1024), idesc_s, kk > 0);
          if (PATTERN == 0 || PATTERN == 2 || PATTERN == 4) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              tc::umma_bf16_ss(tmem + 256 + sb, tc::umma_smem_desc(q_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                               tc::umma_smem_desc(v_addr + (kk >> 2) * (NS * 128) + (kk & 3) * 32, 16, 1024), idesc_s, kk > 0);
          }
          if (commits >= 1) tc::umma_commit(&dummy[0]);
#pragma unroll
          for (int kk = 0; kk < NS / 16; ++kk)
            tc::umma_bf16_ts(tmem + 384, tmem + sb + kk * 8, tc::umma_smem_desc(k_addr + kk * 2048, 8192, 1024), idesc_o, 1u);
          if (commits >= 2) tc::umma_commit(&dummy[1]);
          if (commits >= 3) tc::umma_commit(&dummy[2]);
          if (PATTERN == 4) {
#pragma unroll
            for (int kk = 0; kk < NS / 16; ++kk)
              tc::umma_bf16_ts(tmem + 128, tmem + 256 + sb + kk * 8, tc::umma_smem_desc(v_addr + kk * 2048, 8192, 1024), idesc_o, 1u);
          }
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if (lane == 0) { out[0] = t1 - t0; stop = 1; }
  } else if (warp >= 2 && contend >= 3) {
    // waiting-primitive study: the elementwise warps wait on an mbarrier that never completes while the MMA warp runs
    const uint32_t addr = tc::smem_u32(&dummy[3]);
    const int nwarps = (contend == 7) ? 1 : 8;
    if (warp - 2 < nwarps) {
      while (!stop) {
        uint32_t ok = 0;
        if (contend == 3 || contend == 6 || contend == 7) {
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(addr), "r"(0u) : "memory");
          if (contend == 6) __nanosleep(40);
        } else if (contend == 4) {
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(addr), "r"(0u) : "memory");
        } else if (contend == 5) {
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(addr), "r"(0u), "r"(2000u) : "memory");
        }
        if (ok) break;
      }
    }
  } else if (warp >= 2 && contend) {
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int half = (warp - 2) >> 2;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t s[32], d[32], pk[16];
      tc::tmem_ld_32x32b_x32(tmem + lane_off + half * 32, s);
      tc::tmem_ld_32x32b_x32(tmem + lane_off + 256 + half * 32, d);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) pk[c] = s[2 * c] ^ d[2 * c + 1];
      if (contend > 1) {
        tc::tmem_st_32x32b_x16(tmem + lane_off + 64 + half * 16, pk);
        tc::tmem_st_wait();
      }
      acc ^= pk[3];
    }
    if (acc == 0x12345) out[1] = acc;
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int PATTERN>
void run(const char* name, int ideal) {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 200;
  auto k = attn_pattern<PATTERN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 2048);
  for (int contend = 0; contend <= 7; ++contend) {
    if (contend == 1) continue;
    const int commits = 3;
    k<<<148, 320, 3 * 32768 + 2048>>>(d, reps, contend, commits);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-40s mode=%d : %.0f clk/block (math-ideal %d) [%s]\n", name, contend, double(h) / reps, ideal, cudaGetErrorString(e));
  }
  cudaFree(d);
}
// modes: 0 idle warps | 2 tcgen05.ld/st loop | 3 test_wait spin x8 warps | 4 try_wait x8 | 5 try_wait hint 2000ns x8 | 6 test_wait+nanosleep(40) x8 | 7 test_wait x1 warp
int main() {
  run<0>("dQ  64-key : 16 SS N=64 + 4 TS N=128", 768);
  run<1>("fwd 64-key :  8 SS N=64 + 4 TS N=128", 512);
  run<2>("dQ  128-key: 16 SS N=128 + 8 TS N=128", 1536);
  run<3>("fwd 128-key:  8 SS N=128 + 8 TS N=128", 1024);
  run<4>("dKV 64-row : 16 SS N=64 + 8 TS N=128", 1024);
  return 0;
}
