// Shared definitions of the tcgen05 attention-backward kernels (attention_bwd_tc.cu: delta + dK/dV; attention_bwd_dq.cu: dQ).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/pfn_b200.h"

namespace pfn {

constexpr int AB_DH = 128;
constexpr int AB_THREADS = 320;                   // warp 0 TMA, warp 1 MMA, warps 2..9 elementwise (2 per TMEM lane quarter)
constexpr int AB_EW_THREADS = 256;
constexpr int AB_EW_WARPS = 8;
constexpr int AB_TILE_BYTES = 128 * AB_DH * 2;     // 32 KB : 128-row operand tile (2 chunks of 16 KB)
constexpr int AB_BLK_BYTES = 64 * AB_DH * 2;       // 16 KB : 64-row operand block (2 chunks of 8 KB)
constexpr int AB_KS = 5;                           // depth of the 64-row block ring (TMA runs 3 blocks ahead of the MMAs)
constexpr int AB_SMEM = 2 * AB_TILE_BYTES + AB_KS * 2 * AB_BLK_BYTES + 1024 /*lse/delta*/ + 256 + 1024;   // both kernels                      // dQ kernel

struct AttnBwdParams {
  int T, B, H, sep;
  float scale, scale_log2;
  const __nv_bfloat16* qkv; int ld_qkv;
  const __nv_bfloat16* out; int ld_out;
  const __nv_bfloat16* dout; int ld_dout;
  __nv_bfloat16* dqkv; int ld_dqkv;
  const float* lse;
  float* delta;
  int delta_tm;          // 1: delta is token-major [T*B, H] and already filled (GEMM ROWDOT epilogue); 0: [B*H, T] scratch
  float* dq_colsum;      // optional [H*dh]: += column sums of dQ (dQ kernel epilogue)
  uint32_t drop_seed; int drop_thr;   // dropout on the attention probabilities (thr 0 = off), csrc/dropout.cuh
  int n_tiles;
  int total_work;
  int batch_major;
  long long* trace; int trace_cap;
};

__device__ __forceinline__ void ab_load32(const __nv_bfloat16* p, float (&v)[32]) {
  const uint4* src = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 pk = src[q];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(h[j]);
      v[q * 8 + 2 * j] = t.x;
      v[q * 8 + 2 * j + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void ab_store32(__nv_bfloat16* p, const float (&v)[32]) {
#pragma unroll
  for (int e = 0; e < 32; e += 8) {
    uint4 pk;
    pk.x = tc::pack_bf16x2(v[e], v[e + 1]);
    pk.y = tc::pack_bf16x2(v[e + 2], v[e + 3]);
    pk.z = tc::pack_bf16x2(v[e + 4], v[e + 5]);
    pk.w = tc::pack_bf16x2(v[e + 6], v[e + 7]);
    *reinterpret_cast<uint4*>(p + e) = pk;
  }
}
__device__ __forceinline__ float ab_dot128(const __nv_bfloat16* a, const __nv_bfloat16* b) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const uint4 pa = reinterpret_cast<const uint4*>(a)[c];
    const uint4 pb = reinterpret_cast<const uint4*>(b)[c];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&pa);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&pb);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(ha[j]);
      const float2 y = __bfloat1622float2(hb[j]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  return acc;
}

// 128x64x128 SS MMA: D[tmem] = A[128 rows x 128 dh, K-major tile] * B[64 rows x 128 dh, K-major block]^T
// (all ab_mma_* helpers must be called by exactly one elected lane of a converged warp)
__device__ __forceinline__ void ab_mma_ss_128x64(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr) {
  constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 64, 0, 0);
#pragma unroll
  for (int kk = 0; kk < AB_DH / 16; ++kk) {
    const uint64_t a_desc = tc::umma_smem_desc(a_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
    const uint64_t b_desc = tc::umma_smem_desc(b_addr + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
    tc::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, kk > 0 ? 1u : 0u);
  }
}
// 128x128x64 TS MMA: D[tmem] (+)= A[tmem, 128 x 64 packed bf16] * B[64 rows x 128 dh block read MN-major]
__device__ __forceinline__ void ab_mma_ts_128x128(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, bool accumulate) {
  constexpr uint32_t idesc = tc::umma_idesc_bf16(128, 128, 0, 1);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint64_t b_desc = tc::umma_smem_desc(b_addr + kk * 2048, 8192, 1024);
    tc::umma_bf16_ts(d_tmem, a_tmem + kk * 8, b_desc, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}


__device__ __forceinline__ int ab_tile_block_plan(int i0, int sep, int T, int nblk, int (&dstart)[2]) {
  int nd = 0;
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int lo = i0 + 64 * jj;
    if (lo < T && lo + 63 >= sep) dstart[nd++] = lo;
  }
  return nblk + nd;
}

// 32 consecutive elements (columns c0..c0+31) of row r of a [128 x 128] bf16 tile stored as two 64-column chunks with
// the TMA 128-byte swizzle (16-byte unit u of a row sits at position u ^ (r & 7)).  Conflict-free for one row per lane.
__device__ __forceinline__ void ab_load32_swz(const uint8_t* tile, int r, int c0, float (&v)[32]) {
  const uint8_t* base = tile + (c0 >> 6) * 16384 + r * 128;
  const int u0 = (c0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 pk = *reinterpret_cast<const uint4*>(base + (((u0 + q) ^ (r & 7)) << 4));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(h[j]);
      v[q * 8 + 2 * j] = t.x;
      v[q * 8 + 2 * j + 1] = t.y;
    }
  }
}

// launchers of the individual backward kernels (called by pfn_attention_bwd_tc)
int launch_attn_bwd_dq(const AttnBwdParams& p, const pfn_attn_desc* d, cudaStream_t stream);

}  // namespace pfn
