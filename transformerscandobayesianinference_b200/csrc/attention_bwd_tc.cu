// tcgen05 backward of the single_eval_pos-masked attention (head dim 128, bf16), three kernels:
//
//  (0) attn_bwd_delta_kernel   delta[b,h,i] = dO_i . O_i   (one warp per token row, HBM-bound, 16 B vectors)
//
//  (1) attn_bwd_dq_tc_kernel   [attention_bwd_dq.cu] one CTA per (batch, head, 128-row query tile); blocks = 64-key blocks of the train keys
//      followed by up to two "diagonal" blocks (the tile's own rows as keys; row i keeps only key i, cf. attention_tc.cu):
//          S_j  = Q K_j^T          dP_j = dO V_j^T                      (SS MMAs, 128x64x128, TMEM double-buffered)
//          dS_j = exp2(S_j c - lse) * (dP_j - delta) * scale   -> bf16 -> TMEM   (one thread per query row)
//          dQ  += dS_j K_j                                                (TS MMA, K_j read MN-major from the same smem)
//      For a query row the diagonal key is attended by that row only, so dK_i = dS_ii q_i and dV_i = P_ii dO_i are
//      complete: the owning thread scales its own q / dO row (read back from the swizzled smem tiles) and stores them.
//      Q and dO tiles are double-buffered so the next tile's loads overlap this tile's epilogue.
//
//  (2) attn_bwd_dkv_tc_kernel  one CTA per (batch, head, 128-key tile of the train keys), loop over 64-row blocks i:
//          S^T_i = K Q_i^T         dP^T_i = V dO_i^T                      (lane = key, column = query row)
//          P^T, dS^T -> bf16 -> TMEM ;  dV += P^T dO_i ;  dK += dS^T Q_i  (TS MMAs; Q_i / dO_i blocks read MN-major
//                                                                          from the very smem bytes used K-major above)
//
// Same warp roles / mbarrier protocol as the forward kernel (attention_tc.cu): warp 0 TMA, warp 1 MMA issue,
// warps 2..5 one thread per TMEM lane.  All tiles come straight out of the packed [T*B, 3E] qkv / [T*B, E] dO
// buffers through 3-D TMA maps (column, batch, time); no transposes, no atomics, deterministic.
#include <stdlib.h>

#include "attention_bwd_common.cuh"
#include "dropout.cuh"

namespace pfn {

int check_attn_desc_public(const pfn_attn_desc* d, bool bwd, const char* who);

// =====================================================================================================================
// Kernel 0: delta = rowsum(dO * O) per (batch, head, row)
// =====================================================================================================================
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ out, int ld_out, const __nv_bfloat16* __restrict__ dout,
                      int ld_dout, float* __restrict__ delta, int T, int B, int H, int batch_major) {
  // one warp per token; a 128-wide head = 16 lanes x 8 elements, so a warp covers two heads per pass
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long rows = static_cast<long long>(T) * B;
  for (long long tok = warp; tok < rows; tok += nwarps) {
    const int t = batch_major ? static_cast<int>(tok % T) : static_cast<int>(tok / B);
    const int b = batch_major ? static_cast<int>(tok / T) : static_cast<int>(tok % B);
    for (int h0 = 0; h0 < H; h0 += 2) {
      const int h = h0 + (lane >> 4);
      float acc = 0.f;
      if (h < H) {
        const int col = h * AB_DH + (lane & 15) * 8;
        const uint4 pa = *reinterpret_cast<const uint4*>(out + tok * ld_out + col);
        const uint4 pb = *reinterpret_cast<const uint4*>(dout + tok * ld_dout + col);
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
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((lane & 15) == 0 && h < H) delta[(static_cast<size_t>(b) * H + h) * T + t] = acc;
    }
  }
}

// =====================================================================================================================
// Kernel 2: dK, dV of the train keys
// =====================================================================================================================
__global__ void __launch_bounds__(AB_THREADS + 32, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                       const __grid_constant__ CUtensorMap tmDO64, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET on the __shared__ symbol (an integer round trip of the pointer makes every access through it a
  // generic LD.E / ST.E instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_TILE_BYTES;
  uint8_t* sQD = smem + 2 * AB_TILE_BYTES;               // stage s: Q block at +s*32K, dO block at +16K
  float* sStat = reinterpret_cast<float*>(smem + 2 * AB_TILE_BYTES + AB_KS * 2 * AB_BLK_BYTES);   // [2][2][64]: lse2, delta*scale
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * AB_TILE_BYTES + AB_KS * 2 * AB_BLK_BYTES + 1024);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qd_full = bars + 2;                  // [AB_KS]
  uint64_t* qd_empty = bars + 2 + AB_KS;         // [AB_KS]
  uint64_t* st_full = bars + 2 + 2 * AB_KS;      // S^T / dP^T of a block are in TMEM (one phase per block)
  uint64_t* s_consumed = st_full + 1;            // every row thread has read them into registers (one phase per block)
  uint64_t* pds_ready = st_full + 2;             // [2] P^T / dS^T written (bf16, TMEM)
  uint64_t* pd_free = st_full + 4;               // [2] the dV / dK MMAs that read that P^T / dS^T buffer are complete
  uint64_t* acc_done = st_full + 6;
  uint64_t* acc_empty = st_full + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_full + 8);

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
    for (int s = 0; s < AB_KS; ++s) {
      tc::mbar_init(&qd_full[s], 1);
      tc::mbar_init(&qd_empty[s], 1);
    }
    tc::mbar_init(st_full, 1);
    tc::mbar_init(s_consumed, AB_EW_WARPS);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&pds_ready[s], AB_EW_WARPS);
      tc::mbar_init(&pd_free[s], 1);
    }
    tc::mbar_init(acc_done, 1);
    tc::mbar_init(acc_empty, AB_EW_WARPS);
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
  // TMEM columns: S^T @0 | dP^T @64 | P^T[2] @128,192 (32 cols, bf16) | dS^T[2] @160,224 | dV @256 | dK @384.
  // The fp32 score buffer is single: it is free again as soon as the row threads have pulled it into registers
  // (s_consumed), so the next block's score MMAs run under this block's exp/convert work, while the bf16 P^T / dS^T
  // operands of the accumulate MMAs are double-buffered on their own columns.

  if (warp == 0) {
    {
      uint32_t g = 0, tcount = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        const int bh = w / p.n_tiles;
        const int kt = w - bh * p.n_tiles;
        const int b = bh / p.H, h = bh - b * p.H;
        const int j0 = kt * 128;
        tc::mbar_wait(kv_empty, (tcount & 1) ^ 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(kv_full, 2 * AB_TILE_BYTES);
          tc::tma_load_3d(sK, &tmQKV128, kv_full, E + h * AB_DH, b, j0);
          tc::tma_load_3d(sK + 16384, &tmQKV128, kv_full, E + h * AB_DH + 64, b, j0);
          tc::tma_load_3d(sV, &tmQKV128, kv_full, 2 * E + h * AB_DH, b, j0);
          tc::tma_load_3d(sV + 16384, &tmQKV128, kv_full, 2 * E + h * AB_DH + 64, b, j0);
        }
        __syncwarp();
        for (int i = 0; i < nq; ++i, ++g) {
          const int st = g % AB_KS;
          tc::mbar_wait(&qd_empty[st], ((g / AB_KS) & 1) ^ 1);
          uint8_t* qdst = sQD + st * 2 * AB_BLK_BYTES;
          uint8_t* ddst = qdst + AB_BLK_BYTES;
          const int i0 = i * 64;
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&qd_full[st], 2 * AB_BLK_BYTES);
            tc::tma_load_3d(qdst, &tmQKV64, &qd_full[st], h * AB_DH, b, i0);
            tc::tma_load_3d(qdst + 8192, &tmQKV64, &qd_full[st], h * AB_DH + 64, b, i0);
            tc::tma_load_3d(ddst, &tmDO64, &qd_full[st], h * AB_DH, b, i0);
            tc::tma_load_3d(ddst + 8192, &tmDO64, &qd_full[st], h * AB_DH + 64, b, i0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      // Score issuer (S^T, dP^T of every 64-row block).  The accumulate MMAs are issued by warp 10: with a single issuing
      // thread the score issue, the wait for P^T / dS^T, the accumulate issue and the barrier polls were served in series
      // (see attention_bwd_dq.cu); nothing orders the two streams beyond the barriers that already exist (s_consumed guards
      // the single fp32 score buffer, pd_free the bf16 operand buffers, qd_empty the Q / dO ring).
      const uint32_t k_addr = tc::smem_u32(sK), v_addr = tc::smem_u32(sV);
      const uint32_t qd_addr0 = tc::smem_u32(sQD);
      int sst = 0;               // ring stage of the block whose scores are issued next
      uint32_t sph = 0;
      uint32_t nsc = 0;          // score batches issued so far
      uint32_t tcount = 0;
      tc::KernelTrace tr = tc::trace_make(p.trace, p.trace_cap, 1);
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        tc::mbar_wait(kv_full, tcount & 1);
        for (int i = 0; i < nq; ++i) {
          const uint32_t q_addr = qd_addr0 + static_cast<uint32_t>(sst) * (2 * AB_BLK_BYTES);
          tc::mbar_wait(&qd_full[sst], sph);
          if (nsc > 0) tc::mbar_wait(s_consumed, (nsc - 1) & 1);      // the previous block's scores are in registers
          tc::tc_fence_after();
          if (lane == 0) tr.log(11, tcount, i);
          if (tc::elect_one()) {
            ab_mma_ss_128x64(tmem_base, k_addr, q_addr);                             // S^T  = K Q^T
            ab_mma_ss_128x64(tmem_base + 64, v_addr, q_addr + AB_BLK_BYTES);         // dP^T = V dO^T
            tc::umma_commit(st_full);
            if (i + 1 == nq) tc::umma_commit(kv_empty);                              // K / V tile no longer read
          }
          __syncwarp();
          if (lane == 0) tr.log(14, tcount, i);
          ++nsc;
          if (++sst == AB_KS) { sst = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    {
      // Accumulate issuer: dV += P^T dO_i, dK += dS^T Q_i
      const uint32_t qd_addr0 = tc::smem_u32(sQD);
      int ast = 0;
      uint32_t g = 0, tcount = 0;
      tc::KernelTrace tr = tc::trace_make(p.trace, p.trace_cap, 10);
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        for (int i = 0; i < nq; ++i, ++g) {
          const uint32_t buf = g & 1;
          tc::mbar_wait(&pds_ready[buf], (g >> 1) & 1);
          if (i == 0) tc::mbar_wait(acc_empty, (tcount & 1) ^ 1);
          tc::tc_fence_after();
          if (lane == 0) tr.log(12, tcount, i);
          const uint32_t q_addr = qd_addr0 + static_cast<uint32_t>(ast) * (2 * AB_BLK_BYTES);
          if (tc::elect_one()) {
            ab_mma_ts_128x128(tmem_base + 256, tmem_base + 128 + buf * 64, q_addr + AB_BLK_BYTES, i > 0);   // dV += P^T dO
            ab_mma_ts_128x128(tmem_base + 384, tmem_base + 160 + buf * 64, q_addr, i > 0);                  // dK += dS^T Q
            tc::umma_commit(&qd_empty[ast]);
            tc::umma_commit(&pd_free[buf]);
            if (i + 1 == nq) tc::umma_commit(acc_done);    // one phase per tile
          }
          __syncwarp();
          if (lane == 0) tr.log(13, tcount, i);
          if (++ast == AB_KS) ast = 0;
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;              // which 32 of a block's 64 query-row columns this warp owns
    const int row = quarter * 32 + lane;           // key within the tile
    const int st_tid = threadIdx.x - 64;           // 0..255
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    tc::KernelTrace tr = tc::trace_make(p.trace, p.trace_cap, warp);
    uint32_t g = 0, tcount = 0;
    // Row statistics (lse * log2e, delta * scale) of a 64-row block, one value per thread 0..127.  They are fetched one
    // block AHEAD into a register and parked in the other smem buffer after this block's bar.sync, so the global-load
    // latency is off the per-block critical path.
    // load_stat returns the RAW global value (no arithmetic on it, so nothing waits for the load where it is issued);
    // finish_stat applies the scaling / row masking when the value is parked in shared memory a block later.
    auto load_stat = [&](int ww, int ii) -> float {
      if (st_tid >= 128 || ww >= p.total_work) return 0.f;
      const int bh2 = ww / p.n_tiles;
      const int r = min(ii * 64 + (st_tid & 63), p.T - 1);
      if (st_tid >= 64 && p.delta_tm) return __ldg(p.delta + (static_cast<size_t>(r) * p.B + bh2 / p.H) * p.H + bh2 % p.H);
      const float* src = st_tid < 64 ? p.lse : p.delta;
      return __ldg(src + static_cast<size_t>(bh2) * p.T + r);
    };
    auto finish_stat = [&](float raw, int ii) -> float {
      const bool ok = ii * 64 + (st_tid & 63) < p.T;
      if (st_tid < 64) return ok ? raw * 1.4426950408889634f : INFINITY;
      return ok ? raw * p.scale : 0.f;
    };
    if (st_tid < 128 && static_cast<int>(blockIdx.x) < p.total_work) sStat[st_tid] = finish_stat(load_stat(blockIdx.x, 0), 0);
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
      const int bh = w / p.n_tiles;
      const int kt = w - bh * p.n_tiles;
      const int b = bh / p.H, h = bh - b * p.H;
      const int j = kt * 128 + row;
      const bool key_ok = j < p.sep;               // sep <= T
      for (int i = 0; i < nq; ++i, ++g) {
        const uint32_t buf = g & 1;
        float* stat = sStat + buf * 128;
        const float stat_next = (i + 1 < nq) ? load_stat(w, i + 1) : load_stat(w + gridDim.x, 0);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // stat[buf] visible; everyone is done reading stat[buf ^ 1]
        if (lane == 0) tr.log(24 + 100 * warp, tcount, i);
        tc::mbar_wait_rows(st_full, g & 1);
        if (lane == 0) tr.log(20 + 100 * warp, tcount, i);
        tc::tc_fence_after();
        {
          uint32_t s[32], dp[32], pkp[16], pkd[16];
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + half * 32, s);
          tc::tmem_ld_32x32b_x32(tmem_base + lane_off + 64 + half * 32, dp);
          tc::tmem_ld_wait();
          if (lane == 0) tr.log(25 + 100 * warp, tcount, i);
          tc::tc_fence_before();
          tc::mbar_arrive_warp(s_consumed);          // the score buffer may be overwritten by the next block's MMAs
          if (key_ok && p.drop_thr == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int col = half * 32 + 4 * c;
              const float4 l2 = *reinterpret_cast<const float4*>(&stat[col]);
              const float4 dl = *reinterpret_cast<const float4*>(&stat[64 + col]);
              const float p0 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * c]), p.scale_log2, -l2.x));
              const float p1 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * c + 1]), p.scale_log2, -l2.y));
              const float p2 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * c + 2]), p.scale_log2, -l2.z));
              const float p3 = tc::fast_exp2(fmaf(__uint_as_float(s[4 * c + 3]), p.scale_log2, -l2.w));
              pkp[2 * c] = tc::pack_bf16x2(p0, p1);
              pkp[2 * c + 1] = tc::pack_bf16x2(p2, p3);
              pkd[2 * c] = tc::pack_bf16x2(p0 * fmaf(__uint_as_float(dp[4 * c]), p.scale, -dl.x),
                                           p1 * fmaf(__uint_as_float(dp[4 * c + 1]), p.scale, -dl.y));
              pkd[2 * c + 1] = tc::pack_bf16x2(p2 * fmaf(__uint_as_float(dp[4 * c + 2]), p.scale, -dl.z),
                                               p3 * fmaf(__uint_as_float(dp[4 * c + 3]), p.scale, -dl.w));
            }
          } else if (key_ok) {
            // dropout on the probabilities (csrc/dropout.cuh): the keep bit of (query row i, key j) is byte j & 3 of the hash
            // of (row id (b*H + h)*T + i, j >> 2); Pd = P m / (1 - p) feeds dV, dS = P (m dP / (1 - p) - delta) scale feeds dK
            const float dsc = drop_scale(p.drop_thr);
            const uint32_t rowbase = static_cast<uint32_t>(bh) * p.T + i * 64 + half * 32;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float pv[2], dv[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int cc = 2 * c + e;
                const bool keep = drop_keep_byte(drop_hash(p.drop_seed, rowbase + cc, static_cast<uint32_t>(j) >> 2), j & 3, p.drop_thr);
                const float mk = keep ? dsc : 0.f;
                const float pr = tc::fast_exp2(fmaf(__uint_as_float(s[cc]), p.scale_log2, -stat[half * 32 + cc]));
                pv[e] = pr * mk;
                dv[e] = pr * fmaf(__uint_as_float(dp[cc]) * mk, p.scale, -stat[64 + half * 32 + cc]);
              }
              pkp[c] = tc::pack_bf16x2(pv[0], pv[1]);
              pkd[c] = tc::pack_bf16x2(dv[0], dv[1]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) { pkp[c] = 0u; pkd[c] = 0u; }
          }
          if (lane == 0) tr.log(26 + 100 * warp, tcount, i);
          tc::mbar_wait(&pd_free[buf], ((g >> 1) & 1) ^ 1);      // accumulate MMAs of block g-2 no longer read this buffer
          if (lane == 0) tr.log(27 + 100 * warp, tcount, i);
          tc::tc_fence_after();
          tc::tmem_st_32x32b_x16(tmem_base + lane_off + 128 + buf * 64 + half * 16, pkp);
          tc::tmem_st_32x32b_x16(tmem_base + lane_off + 160 + buf * 64 + half * 16, pkd);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive_warp(&pds_ready[buf]);
        if (st_tid < 128) sStat[(buf ^ 1) * 128 + st_tid] = finish_stat(stat_next, i + 1 < nq ? i + 1 : 0);   // parked for the next block
        if (lane == 0) tr.log(21 + 100 * warp, tcount, i);
      }
      tc::mbar_wait(acc_done, tcount & 1);                    // committed once per tile, after its last dV/dK MMA
      if (lane == 0) tr.log(22 + 100 * warp, tcount, 0);
      tc::tc_fence_after();
      const bool store_ok = j < p.sep && j < p.T;
      const size_t krow = p.batch_major ? static_cast<size_t>(b) * p.T + (store_ok ? j : 0) : static_cast<size_t>(store_ok ? j : 0) * p.B + b;
      __nv_bfloat16* drow = p.dqkv + krow * p.ld_dqkv + h * AB_DH;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        float acc[32];
        // half 0 stores dV (TMEM columns 256..383), half 1 stores dK (384..511)
        tc::tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + half * 128 + c * 32, raw);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(raw[e]);
        if (store_ok) ab_store32(drow + (half == 0 ? 2 * E : E) + c * 32, acc);
      }
      tc::tc_fence_before();
      tc::mbar_arrive_warp(acc_empty);
      if (lane == 0) tr.log(23 + 100 * warp, tcount, 0);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

static int make_map3d(CUtensorMap* tm, const void* base, int ld, int width, int B, int T, int box_rows, int batch_major) {
  uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(B), static_cast<uint64_t>(T)};
  uint64_t strides[3] = {0, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * B};
  if (batch_major) { strides[1] = static_cast<uint64_t>(ld) * 2 * T; strides[2] = static_cast<uint64_t>(ld) * 2; }
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
  if (int rc = make_map3d(&tmQKV128, d->qkv, d->ld_qkv, 3 * E, d->B, d->T, 128, d->batch_major)) return rc;
  if (int rc = make_map3d(&tmQKV64, d->qkv, d->ld_qkv, 3 * E, d->B, d->T, 64, d->batch_major)) return rc;
  if (int rc = make_map3d(&tmDO128, d->dout, d->ld_dout, E, d->B, d->T, 128, d->batch_major)) return rc;
  if (int rc = make_map3d(&tmDO64, d->dout, d->ld_dout, E, d->B, d->T, 64, d->batch_major)) return rc;
  AttnBwdParams p;
  p.T = d->T; p.B = d->B; p.H = d->H; p.sep = d->sep;
  p.scale = d->scale; p.scale_log2 = d->scale * 1.4426950408889634f;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(d->qkv); p.ld_qkv = d->ld_qkv;
  p.out = reinterpret_cast<const __nv_bfloat16*>(d->out); p.ld_out = d->ld_out;
  p.dout = reinterpret_cast<const __nv_bfloat16*>(d->dout); p.ld_dout = d->ld_dout;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(d->dqkv); p.ld_dqkv = d->ld_dqkv;
  p.lse = d->lse; p.delta = d->delta; p.dq_colsum = d->dq_colsum; p.delta_tm = d->delta_token_major;
  PFN_CHECK_ARG(!(d->delta_token_major && d->batch_major), "attention_bwd_tc: a token-major delta implies the reference token order");
  p.drop_seed = d->drop_seed; p.drop_thr = d->drop_thr;
  p.batch_major = d->batch_major;
  p.trace = nullptr; p.trace_cap = g_trace_cap;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    PFN_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // debug only (tools/time_kernels.py): pfn_debug_attention_trace(NULL, 0, 21|22|23) runs just dK/dV | dQ | delta
  const int only = (g_trace_ptr == nullptr && g_trace_which >= 21 && g_trace_which <= 23) ? g_trace_which : 0;
  if ((only == 0 || only == 23) && !d->delta_token_major) {
    const long long rows = static_cast<long long>(d->T) * d->B;
    long long grid = (rows + 7) / 8;
    if (grid > 8LL * num_sms()) grid = 8LL * num_sms();
    attn_bwd_delta_kernel<<<static_cast<int>(grid), 256, 0, s>>>(p.out, p.ld_out, p.dout, p.ld_dout, p.delta, d->T, d->B, d->H, d->batch_major);
    PFN_LAUNCH_OK();
  }
  if (only == 0 || only == 22) {
    p.trace = g_trace_which == 1 ? g_trace_ptr : nullptr;
    if (int rc = launch_attn_bwd_dq(p, d, s)) return rc;
    p.trace = nullptr;
    PFN_LAUNCH_OK();
  }
  if (d->sep > 0 && (only == 0 || only == 21)) {
    p.n_tiles = (d->sep + 127) / 128;
    p.total_work = p.n_tiles * d->B * d->H;
    int grid = num_sms() < p.total_work ? num_sms() : p.total_work;
    p.trace = g_trace_which == 2 ? g_trace_ptr : nullptr;
    attn_bwd_dkv_tc_kernel<<<grid, AB_THREADS + 32, AB_SMEM, s>>>(tmQKV128, tmQKV64, tmDO64, p);
    p.trace = nullptr;
    PFN_LAUNCH_OK();
  }
  return 0;
}
