// Microbenchmark: tcgen05.ld / tcgen05.st throughput and latency per SM (how many bytes per clock the row warps of the
// attention kernels can pull out of / push into TMEM), and the cost of one mbarrier hand-off between two warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transformerscandobayesianinference_b200/csrc/tc_common.cuh"
using namespace pfn;

__global__ void __launch_bounds__(512, 1) tmem_bw(long long* out, int reps, int mode, int nwarps) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bar_a, bar_b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { tc::mbar_init(&bar_a, 1); tc::mbar_init(&bar_b, 1); tc::mbar_fence_init(); }
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (mode <= 3 && warp < nwarps) {
    for (int r = 0; r < reps; ++r) {
      if (mode == 0) {            // ld x32, wait each
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(tmem + lane_off + ((r * 32) & 255), v);
        tc::tmem_ld_wait();
        acc ^= v[0] ^ v[31];
      } else if (mode == 1) {     // 2 x ld x32, one wait (the attention row pattern)
        uint32_t a[32], b[32];
        tc::tmem_ld_32x32b_x32(tmem + lane_off + ((r * 64) & 255), a);
        tc::tmem_ld_32x32b_x32(tmem + lane_off + 256 + ((r * 64) & 255), b);
        tc::tmem_ld_wait();
        acc ^= a[0] ^ b[31] ^ a[17];
      } else if (mode == 2) {     // st x16 + wait
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = r + i;
        tc::tmem_st_32x32b_x16(tmem + lane_off + ((r * 16) & 255), v);
        tc::tmem_st_wait();
      } else {                    // st x32 + wait
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = r + i;
        tc::tmem_st_32x32b_x32(tmem + lane_off + ((r * 32) & 255), v);
        tc::tmem_st_wait();
      }
    }
  } else if (mode == 4) {         // mbarrier ping-pong between warp 0 and warp 1 (one lane arrives, whole warp polls)
    if (warp == 0) {
      for (int r = 0; r < reps; ++r) {
        tc::mbar_arrive_warp(&bar_a);
        tc::mbar_wait(&bar_b, r & 1);
      }
    } else if (warp == 1) {
      for (int r = 0; r < reps; ++r) {
        tc::mbar_wait(&bar_a, r & 1);
        tc::mbar_arrive_warp(&bar_b);
      }
    }
  } else if (mode == 5) {         // same with the tcgen05 fences + a TMEM st/ld on each side (the real hand-off sequence)
    if (warp == 0) {
      for (int r = 0; r < reps; ++r) {
        uint32_t v[16];
        for (int i = 0; i < 16; ++i) v[i] = r + i;
        tc::tmem_st_32x32b_x16(tmem + lane_off, v);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive_warp(&bar_a);
        tc::mbar_wait(&bar_b, r & 1);
        tc::tc_fence_after();
      }
    } else if (warp == 1) {
      for (int r = 0; r < reps; ++r) {
        tc::mbar_wait(&bar_a, r & 1);
        tc::tc_fence_after();
        uint32_t v[16];
        tc::tmem_ld_32x32b_x16(tmem + lane_off, v);
        tc::tmem_ld_wait();
        acc ^= v[3];
        tc::tc_fence_before();
        tc::mbar_arrive_warp(&bar_b);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678) out[1] = acc;
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 2000;
  const char* names[] = {"ld x32 + wait", "2 x ld x32 + wait", "st x16 + wait", "st x32 + wait", "mbarrier ping-pong (round trip)", "st+fence+arrive -> wait+fence+ld round trip"};
  const int bytes[] = {4096, 8192, 2048, 4096, 0, 0};
  for (int mode = 0; mode < 6; ++mode) {
    for (int nw : {1, 4, 8, 16}) {
      if (mode >= 4 && nw != 4) continue;
      tmem_bw<<<148, 512>>>(d, reps, mode, nw);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double clk = double(h) / reps;
      if (mode < 4) printf("%-46s warps=%2d : %7.1f clk/iter  -> %6.1f B/clk/SM [%s]\n", names[mode], nw, clk, bytes[mode] * nw / clk, cudaGetErrorString(e));
      else printf("%-46s          : %7.1f clk per round trip [%s]\n", names[mode], clk, cudaGetErrorString(e));
    }
  }
  return 0;
}
