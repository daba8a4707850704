// tcgen05 backward of the single_eval_pos-masked attention (head dim 128, bf16), two kernels:
//
//  (1) attn_bwd_dq_tc_kernel   one CTA per (batch, head, 128-row query tile), loop over 64-key blocks j < sep:
//          S_j  = Q K_j^T          dP_j = dO V_j^T                      (SS MMAs, 128x64x128, TMEM double-buffered)
//          dS_j = exp2(S_j c - lse) * (dP_j - delta) * scale   -> bf16 -> TMEM   (one thread per query row)
//          dQ  += dS_j K_j                                                (TS MMA, K_j read MN-major from the same smem)
//      plus, per row, delta_i = dO_i . O_i (stored for kernel 2) and the diagonal key of a query row (i >= sep):
//      its dq contribution and — since nobody else attends to that key — the complete dK_i, dV_i.
//
//  (2) attn_bwd_dkv_tc_kernel  one CTA per (batch, head, 128-key tile of the train keys), loop over 64-row blocks i:
//          S^T_i = K Q_i^T         dP^T_i = V dO_i^T                      (lane = key, column = query row)
//          P^T, dS^T -> bf16 -> TMEM ;  dV += P^T dO_i ;  dK += dS^T Q_i  (TS MMAs; Q_i / dO_i blocks read MN-major
//                                                                          from the very smem bytes used K-major above)
//
// Same warp roles / mbarrier protocol as the forward kernel (attention_tc.cu): warp 0 TMA, warp 1 MMA issue,
// warps 2..5 one thread per TMEM lane.  All tiles come straight out of the packed [T*B, 3E] qkv / [T*B, E] dO
// buffers through 3-D TMA maps (column, batch, time); no transposes, no atomics, deterministic.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/pfn_b200.h"

namespace pfn {

int check_attn_desc_public(const pfn_attn_desc* d, bool bwd, const char* who);

constexpr int AB_DH = 128;
constexpr int AB_THREADS = 192;
constexpr int AB_TILE_BYTES = 128 * AB_DH * 2;     // 32 KB : 128-row operand tile (2 chunks of 16 KB)
constexpr int AB_BLK_BYTES = 64 * AB_DH * 2;       // 16 KB : 64-row operand block (2 chunks of 8 KB)
constexpr int AB_SMEM = 2 * AB_TILE_BYTES + 2 * 2 * AB_BLK_BYTES + 1024 /*lse/delta*/ + 256 + 1024;

struct AttnBwdParams {
  int T, B, H, sep;
  float scale, scale_log2;
  const __nv_bfloat16* qkv; int ld_qkv;
  const __nv_bfloat16* out; int ld_out;
  const __nv_bfloat16* dout; int ld_dout;
  __nv_bfloat16* dqkv; int ld_dqkv;
  const float* lse;
  float* delta;
  int n_tiles;
  int total_work;
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

// =====================================================================================================================
// Kernel 1: dQ (+ delta, + diagonal-key dK/dV of query rows)
// =====================================================================================================================
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO128, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + AB_TILE_BYTES;
  uint8_t* sKV = smem + 2 * AB_TILE_BYTES;               // stage s: K at +s*32K, V at +16K
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * AB_TILE_BYTES + 4 * AB_BLK_BYTES + 1024);
  uint64_t* qdo_full = bars + 0;
  uint64_t* qdo_empty = bars + 1;
  uint64_t* kv_full = bars + 2;     // [2]
  uint64_t* kv_empty = bars + 4;    // [2]
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* ds_ready = bars + 8;    // [2]
  uint64_t* dq_done = bars + 10;
  uint64_t* dq_empty = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int E = p.H * AB_DH;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQKV128);
    tc::tma_prefetch_desc(&tmQKV64);
    tc::tma_prefetch_desc(&tmDO128);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(qdo_full, 1);
    tc::mbar_init(qdo_empty, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&kv_full[s], 1);
      tc::mbar_init(&kv_empty[s], 1);
      tc::mbar_init(&s_full[s], 1);
      tc::mbar_init(&ds_ready[s], 128);
    }
    tc::mbar_init(dq_done, 1);
    tc::mbar_init(dq_empty, 128);
    tc::mbar_fence_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nblk = (p.sep + 63) / 64;
  // TMEM columns: S[2] @0,64 | dP[2] @128,192 | dQ @256..383

  if (warp == 0) {
    if (lane == 0 && nblk > 0) {
      uint32_t g = 0, tcount = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        const int bh = w / p.n_tiles;
        const int qt = w - bh * p.n_tiles;
        const int b = bh / p.H, h = bh - b * p.H;
        const int i0 = qt * 128;
        tc::mbar_wait(qdo_empty, (tcount & 1) ^ 1);
        tc::mbar_expect_tx(qdo_full, 2 * AB_TILE_BYTES);
        tc::tma_load_3d(sQ, &tmQKV128, qdo_full, h * AB_DH, b, i0);
        tc::tma_load_3d(sQ + 16384, &tmQKV128, qdo_full, h * AB_DH + 64, b, i0);
        tc::tma_load_3d(sDO, &tmDO128, qdo_full, h * AB_DH, b, i0);
        tc::tma_load_3d(sDO + 16384, &tmDO128, qdo_full, h * AB_DH + 64, b, i0);
        for (int j = 0; j < nblk; ++j, ++g) {
          const int st = g & 1;
          tc::mbar_wait(&kv_empty[st], ((g >> 1) & 1) ^ 1);
          tc::mbar_expect_tx(&kv_full[st], 2 * AB_BLK_BYTES);
          uint8_t* kdst = sKV + st * 2 * AB_BLK_BYTES;
          uint8_t* vdst = kdst + AB_BLK_BYTES;
          const int j0 = j * 64;
          tc::tma_load_3d(kdst, &tmQKV64, &kv_full[st], E + h * AB_DH, b, j0);
          tc::tma_load_3d(kdst + 8192, &tmQKV64, &kv_full[st], E + h * AB_DH + 64, b, j0);
          tc::tma_load_3d(vdst, &tmQKV64, &kv_full[st], 2 * E + h * AB_DH, b, j0);
          tc::tma_load_3d(vdst + 8192, &tmQKV64, &kv_full[st], 2 * E + h * AB_DH + 64, b, j0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && nblk > 0) {
      const uint32_t q_addr = tc::smem_u32(sQ), do_addr = tc::smem_u32(sDO);
      uint32_t g = 0, tcount = 0;
      auto issue_scores = [&](uint32_t gg) {
        const uint32_t k_addr = tc::smem_u32(sKV + (gg & 1) * 2 * AB_BLK_BYTES);
        ab_mma_ss_128x64(tmem_base + (gg & 1) * 64, q_addr, k_addr);                       // S  = Q K^T
        ab_mma_ss_128x64(tmem_base + 128 + (gg & 1) * 64, do_addr, k_addr + AB_BLK_BYTES); // dP = dO V^T
      };
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        tc::mbar_wait(qdo_full, tcount & 1);
        tc::mbar_wait(&kv_full[g & 1], (g >> 1) & 1);
        tc::tc_fence_after();
        issue_scores(g);
        tc::umma_commit(&s_full[g & 1]);
        for (int j = 0; j < nblk; ++j, ++g) {
          if (j + 1 < nblk) {
            const uint32_t gn = g + 1;
            tc::mbar_wait(&kv_full[gn & 1], (gn >> 1) & 1);
            tc::tc_fence_after();
            issue_scores(gn);
            tc::umma_commit(&s_full[gn & 1]);
          } else {
            tc::umma_commit(qdo_empty);
          }
          tc::mbar_wait(&ds_ready[g & 1], (g >> 1) & 1);
          if (j == 0) tc::mbar_wait(dq_empty, (tcount & 1) ^ 1);
          tc::tc_fence_after();
          const uint32_t k_addr = tc::smem_u32(sKV + (g & 1) * 2 * AB_BLK_BYTES);
          ab_mma_ts_128x128(tmem_base + 256, tmem_base + (g & 1) * 64, k_addr, j > 0);     // dQ += dS K
          tc::umma_commit(&kv_empty[g & 1]);
          tc::umma_commit(dq_done);
        }
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t g = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const int bh = w / p.n_tiles;
      const int qt = w - bh * p.n_tiles;
      const int b = bh / p.H, h = bh - b * p.H;
      const int i = qt * 128 + row;
      const bool valid = i < p.T;
      const bool is_query = valid && i >= p.sep;
      const size_t tok = static_cast<size_t>(valid ? i : 0) * p.B + b;
      const __nv_bfloat16* qrow = p.qkv + tok * p.ld_qkv + h * AB_DH;
      const __nv_bfloat16* dorow = p.dout + tok * p.ld_dout + h * AB_DH;
      float lse2 = INFINITY, delta = 0.f, ds_ii = 0.f, p_ii = 0.f;
      if (valid) {
        lse2 = p.lse[static_cast<size_t>(bh) * p.T + i] * 1.4426950408889634f;
        delta = ab_dot128(dorow, p.out + tok * p.ld_out + h * AB_DH);
        p.delta[static_cast<size_t>(bh) * p.T + i] = delta;
        if (is_query) {
          const float s_ii = ab_dot128(qrow, qrow + E);
          p_ii = tc::fast_exp2(fmaf(s_ii, p.scale_log2, -lse2));
          const float dp_ii = ab_dot128(dorow, qrow + 2 * E);
          ds_ii = p_ii * (dp_ii - delta) * p.scale;
        }
      }
      for (int j = 0; j < nblk; ++j, ++g) {
        const uint32_t buf = g & 1;
        tc::mbar_wait(&s_full[buf], (g >> 1) & 1);
        tc::tc_fence_after();
        const int kmax = p.sep - j * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t s[32], dp[32], pk[16];
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + buf * 64 + half * 32, s);
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + 128 + buf * 64 + half * 32, dp);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int k0 = half * 32 + 2 * c;
            float d0 = 0.f, d1 = 0.f;
            if (k0 < kmax) {
              const float pr = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c]), p.scale_log2, -lse2));
              d0 = pr * (__uint_as_float(dp[2 * c]) - delta) * p.scale;
            }
            if (k0 + 1 < kmax) {
              const float pr = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), p.scale_log2, -lse2));
              d1 = pr * (__uint_as_float(dp[2 * c + 1]) - delta) * p.scale;
            }
            pk[c] = tc::pack_bf16x2(d0, d1);
          }
          tc::tmem_st_32x32b_x16(tmem_base + lane_off + buf * 64 + half * 16, pk);
        }
        tc::tmem_st_wait();
        // observe every dq_done phase in order (see attention_tc.cu: parity waits must not run two phases ahead)
        if (j > 0) tc::mbar_wait(dq_done, (g - 1) & 1);
        tc::tc_fence_before();
        tc::mbar_arrive(&ds_ready[buf]);
      }
      if (nblk > 0) {
        tc::mbar_wait(dq_done, (g - 1) & 1);
        tc::tc_fence_after();
      }
      __nv_bfloat16* dq_out = p.dqkv + tok * p.ld_dqkv + h * AB_DH;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float acc[32];
        if (nblk > 0) {
          uint32_t raw[32];
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + c * 32, raw);
          tc::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(raw[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = 0.f;
        }
        if (is_query) {
          float kk[32], qq[32], dd[32];
          ab_load32(qrow + E + c * 32, kk);
          ab_load32(qrow + c * 32, qq);
          ab_load32(dorow + c * 32, dd);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            acc[e] = fmaf(ds_ii, kk[e], acc[e]);
            qq[e] *= ds_ii;        // dK_i = dS_ii q_i
            dd[e] *= p_ii;         // dV_i = P_ii dO_i
          }
          ab_store32(dq_out + E + c * 32, qq);
          ab_store32(dq_out + 2 * E + c * 32, dd);
        }
        if (valid) ab_store32(dq_out + c * 32, acc);
      }
      if (nblk > 0) {
        tc::tc_fence_before();
        tc::mbar_arrive(dq_empty);
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// Kernel 2: dK, dV of the train keys
// =====================================================================================================================
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                       const __grid_constant__ CUtensorMap tmDO64, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE_BYTES;
  uint8_t* sQD = smem + 2 * AB_TILE_BYTES;               // stage s: Q block at +s*32K, dO block at +16K
  float* sStat = reinterpret_cast<float*>(smem + 2 * AB_TILE_BYTES + 4 * AB_BLK_BYTES);   // [2][2][64]: lse2, delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * AB_TILE_BYTES + 4 * AB_BLK_BYTES + 1024);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qd_full = bars + 2;     // [2]
  uint64_t* qd_empty = bars + 4;    // [2]
  uint64_t* st_full = bars + 6;     // [2]
  uint64_t* pds_ready = bars + 8;   // [2]
  uint64_t* acc_done = bars + 10;
  uint64_t* acc_empty = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int E = p.H * AB_DH;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQKV128);
    tc::tma_prefetch_desc(&tmQKV64);
    tc::tma_prefetch_desc(&tmDO64);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(kv_full, 1);
    tc::mbar_init(kv_empty, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&qd_full[s], 1);
      tc::mbar_init(&qd_empty[s], 1);
      tc::mbar_init(&st_full[s], 1);
      tc::mbar_init(&pds_ready[s], 128);
    }
    tc::mbar_init(acc_done, 1);
    tc::mbar_init(acc_empty, 128);
    tc::mbar_fence_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nq = (p.T + 63) / 64;
  // TMEM columns: S^T[2] @0,64 | dP^T[2] @128,192 | dV @256 | dK @384

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0, tcount = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        const int bh = w / p.n_tiles;
        const int kt = w - bh * p.n_tiles;
        const int b = bh / p.H, h = bh - b * p.H;
        const int j0 = kt * 128;
        tc::mbar_wait(kv_empty, (tcount & 1) ^ 1);
        tc::mbar_expect_tx(kv_full, 2 * AB_TILE_BYTES);
        tc::tma_load_3d(sK, &tmQKV128, kv_full, E + h * AB_DH, b, j0);
        tc::tma_load_3d(sK + 16384, &tmQKV128, kv_full, E + h * AB_DH + 64, b, j0);
        tc::tma_load_3d(sV, &tmQKV128, kv_full, 2 * E + h * AB_DH, b, j0);
        tc::tma_load_3d(sV + 16384, &tmQKV128, kv_full, 2 * E + h * AB_DH + 64, b, j0);
        for (int i = 0; i < nq; ++i, ++g) {
          const int st = g & 1;
          tc::mbar_wait(&qd_empty[st], ((g >> 1) & 1) ^ 1);
          tc::mbar_expect_tx(&qd_full[st], 2 * AB_BLK_BYTES);
          uint8_t* qdst = sQD + st * 2 * AB_BLK_BYTES;
          uint8_t* ddst = qdst + AB_BLK_BYTES;
          const int i0 = i * 64;
          tc::tma_load_3d(qdst, &tmQKV64, &qd_full[st], h * AB_DH, b, i0);
          tc::tma_load_3d(qdst + 8192, &tmQKV64, &qd_full[st], h * AB_DH + 64, b, i0);
          tc::tma_load_3d(ddst, &tmDO64, &qd_full[st], h * AB_DH, b, i0);
          tc::tma_load_3d(ddst + 8192, &tmDO64, &qd_full[st], h * AB_DH + 64, b, i0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t k_addr = tc::smem_u32(sK), v_addr = tc::smem_u32(sV);
      uint32_t g = 0, tcount = 0;
      auto issue_scores = [&](uint32_t gg) {
        const uint32_t q_addr = tc::smem_u32(sQD + (gg & 1) * 2 * AB_BLK_BYTES);
        ab_mma_ss_128x64(tmem_base + (gg & 1) * 64, k_addr, q_addr);                        // S^T  = K Q^T
        ab_mma_ss_128x64(tmem_base + 128 + (gg & 1) * 64, v_addr, q_addr + AB_BLK_BYTES);   // dP^T = V dO^T
      };
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        tc::mbar_wait(kv_full, tcount & 1);
        tc::mbar_wait(&qd_full[g & 1], (g >> 1) & 1);
        tc::tc_fence_after();
        issue_scores(g);
        tc::umma_commit(&st_full[g & 1]);
        for (int i = 0; i < nq; ++i, ++g) {
          if (i + 1 < nq) {
            const uint32_t gn = g + 1;
            tc::mbar_wait(&qd_full[gn & 1], (gn >> 1) & 1);
            tc::tc_fence_after();
            issue_scores(gn);
            tc::umma_commit(&st_full[gn & 1]);
          } else {
            tc::umma_commit(kv_empty);
          }
          tc::mbar_wait(&pds_ready[g & 1], (g >> 1) & 1);
          if (i == 0) tc::mbar_wait(acc_empty, (tcount & 1) ^ 1);
          tc::tc_fence_after();
          const uint32_t q_addr = tc::smem_u32(sQD + (g & 1) * 2 * AB_BLK_BYTES);
          ab_mma_ts_128x128(tmem_base + 256, tmem_base + (g & 1) * 64, q_addr + AB_BLK_BYTES, i > 0);   // dV += P^T dO
          ab_mma_ts_128x128(tmem_base + 384, tmem_base + 128 + (g & 1) * 64, q_addr, i > 0);            // dK += dS^T Q
          tc::umma_commit(&qd_empty[g & 1]);
          tc::umma_commit(acc_done);
        }
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // key within the tile
    const int st_tid = threadIdx.x - 64;           // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t g = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const int bh = w / p.n_tiles;
      const int kt = w - bh * p.n_tiles;
      const int b = bh / p.H, h = bh - b * p.H;
      const int j = kt * 128 + row;
      const bool key_ok = j < p.sep;               // sep <= T
      for (int i = 0; i < nq; ++i, ++g) {
        const uint32_t buf = g & 1;
        float* stat = sStat + buf * 128;
        {
          const int r = i * 64 + (st_tid & 63);
          float v;
          if (st_tid < 64) v = r < p.T ? p.lse[static_cast<size_t>(bh) * p.T + r] * 1.4426950408889634f : INFINITY;
          else v = r < p.T ? p.delta[static_cast<size_t>(bh) * p.T + r] : 0.f;
          stat[st_tid] = v;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tc::mbar_wait(&st_full[buf], (g >> 1) & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t s[32], dp[32], pkp[16], pkd[16];
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + buf * 64 + half * 32, s);
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + 128 + buf * 64 + half * 32, dp);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = half * 32 + 2 * c;
            const float2 l2 = *reinterpret_cast<const float2*>(&stat[col]);
            const float2 dl = *reinterpret_cast<const float2*>(&stat[64 + col]);
            float p0 = 0.f, p1 = 0.f;
            if (key_ok) {
              p0 = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c]), p.scale_log2, -l2.x));
              p1 = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), p.scale_log2, -l2.y));
            }
            const float d0 = p0 * (__uint_as_float(dp[2 * c]) - dl.x) * p.scale;
            const float d1 = p1 * (__uint_as_float(dp[2 * c + 1]) - dl.y) * p.scale;
            pkp[c] = tc::pack_bf16x2(p0, p1);
            pkd[c] = tc::pack_bf16x2(d0, d1);
          }
          tc::tmem_st_32x32b_x16(tmem_base + lane_off + buf * 64 + half * 16, pkp);
          tc::tmem_st_32x32b_x16(tmem_base + lane_off + 128 + buf * 64 + half * 16, pkd);
        }
        tc::tmem_st_wait();
        if (i > 0) tc::mbar_wait(acc_done, (g - 1) & 1);   // keep in step with every acc_done phase
        tc::tc_fence_before();
        tc::mbar_arrive(&pds_ready[buf]);
      }
      tc::mbar_wait(acc_done, (g - 1) & 1);
      tc::tc_fence_after();
      const bool store_ok = j < p.sep && j < p.T;
      __nv_bfloat16* drow = p.dqkv + (static_cast<size_t>(store_ok ? j : 0) * p.B + b) * p.ld_dqkv + h * AB_DH;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t raw[32];
        float acc[32];
        tc::tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + c * 32, raw);   // c < 4: dV, c >= 4: dK
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(raw[e]);
        if (store_ok) {
          if (c < 4) ab_store32(drow + 2 * E + c * 32, acc);
          else ab_store32(drow + E + (c - 4) * 32, acc);
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acc_empty);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

static int make_map3d(CUtensorMap* tm, const void* base, int ld, int width, int B, int T, int box_rows) {
  uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(B), static_cast<uint64_t>(T)};
  uint64_t strides[3] = {0, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * B};
  uint32_t box[3] = {64, 1, static_cast<uint32_t>(box_rows)};
  return make_tensor_map_bf16(tm, base, 3, dims, strides, box, true);
}

}  // namespace pfn

using namespace pfn;

extern "C" int pfn_attention_bwd_tc(const pfn_attn_desc* d, void* stream) {
  if (int rc = check_attn_desc_public(d, true, "attention_bwd_tc")) return rc;
  PFN_CHECK_ARG(d->dtype == PFN_BF16, "attention_bwd_tc: bf16 only");
  PFN_CHECK_ARG(d->dh == AB_DH, "attention_bwd_tc: head dim %d unsupported (built for 128)", d->dh);
  PFN_CHECK_ARG(d->ld_qkv % 8 == 0 && d->ld_out % 8 == 0 && d->ld_dout % 8 == 0 && d->ld_dqkv % 8 == 0,
                "attention_bwd_tc: leading dims must be multiples of 8");
  PFN_CHECK_ARG(((reinterpret_cast<uintptr_t>(d->qkv) | reinterpret_cast<uintptr_t>(d->out) |
                  reinterpret_cast<uintptr_t>(d->dout) | reinterpret_cast<uintptr_t>(d->dqkv)) & 15) == 0,
                "attention_bwd_tc: buffers must be 16-byte aligned");
  const int E = d->H * d->dh;
  CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDO64;
  if (int rc = make_map3d(&tmQKV128, d->qkv, d->ld_qkv, 3 * E, d->B, d->T, 128)) return rc;
  if (int rc = make_map3d(&tmQKV64, d->qkv, d->ld_qkv, 3 * E, d->B, d->T, 64)) return rc;
  if (int rc = make_map3d(&tmDO128, d->dout, d->ld_dout, E, d->B, d->T, 128)) return rc;
  if (int rc = make_map3d(&tmDO64, d->dout, d->ld_dout, E, d->B, d->T, 64)) return rc;
  AttnBwdParams p;
  p.T = d->T; p.B = d->B; p.H = d->H; p.sep = d->sep;
  p.scale = d->scale; p.scale_log2 = d->scale * 1.4426950408889634f;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(d->qkv); p.ld_qkv = d->ld_qkv;
  p.out = reinterpret_cast<const __nv_bfloat16*>(d->out); p.ld_out = d->ld_out;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(d->dout); p.ld_dout = d->ld_dout;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(d->dqkv); p.ld_dqkv = d->ld_dqkv;
  p.lse = d->lse; p.delta = d->delta;
  static bool attr_set = false;
  if (!attr_set) {
    PFN_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    PFN_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_set = true;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {
    p.n_tiles = (d->T + 127) / 128;
    p.total_work = p.n_tiles * d->B * d->H;
    int grid = num_sms() < p.total_work ? num_sms() : p.total_work;
    attn_bwd_dq_tc_kernel<<<grid, AB_THREADS, AB_SMEM, s>>>(tmQKV128, tmQKV64, tmDO128, p);
    PFN_LAUNCH_OK();
  }
  if (d->sep > 0) {
    p.n_tiles = (d->sep + 127) / 128;
    p.total_work = p.n_tiles * d->B * d->H;
    int grid = num_sms() < p.total_work ? num_sms() : p.total_work;
    attn_bwd_dkv_tc_kernel<<<grid, AB_THREADS, AB_SMEM, s>>>(tmQKV128, tmQKV64, tmDO64, p);
    PFN_LAUNCH_OK();
  }
  return 0;
}
