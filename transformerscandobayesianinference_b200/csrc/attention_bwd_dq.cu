// tcgen05 backward of the single_eval_pos-masked attention, dQ part (head dim 128, bf16) -- round-2 kernel.
//
// One persistent CTA per SM walks (batch, head, 128-row query tile) work items.  For a tile, blocks = up to two
// "diagonal" 64-key blocks (the tile's own rows as keys: query row i keeps only key i) FIRST, then the 64-key blocks of the
// train keys [0, sep):
//     S_j  = Q K_j^T          dP_j = dO V_j^T                    (TS MMAs 128x64x128: Q / dO are read from TMEM)
//     dS_j = exp2(S_j c - lse) (dP_j - delta) scale -> bf16 -> TMEM   (one thread per (row, 32-key half))
//     dQ  += dS_j K_j                                               (TS MMA, K_j read MN-major from the same smem block)
// For a query row the diagonal key is attended by that row only: dK_i = dS_ii q_i, dV_i = P_ii dO_i are complete and are
// written by the thread that owns the row (reference semantics: transformer.py:35-41 mask, torch SDPA backward).
//
// What changed against the round-1 kernel (profiles/r2_attn_*.md has the measurements that motivated each item):
//   * The resident operands Q and dO are copied once per tile from the TMA-written smem tile into TMEM, so the 16 score MMAs
//     of a block run in TS mode: they read 2 KB instead of 6 KB of shared memory each and reach the 32-clock floor of an
//     N = 64 MMA (tools/ubench/attn_pattern2.cu: 770 vs 1004 clocks per block for the whole MMA sequence).
//   * Diagonal blocks come first, so the smem Q / dO tile is dead early and the NEXT tile's Q / dO loads overlap this tile's
//     dense blocks; the next tile's TMEM copy, its first score MMAs and this tile's epilogue overlap as well (the round-1
//     kernel spent ~35 % of its time in the exposed tile prologue / epilogue).
//   * The MMA warp carries ring positions as (stage, phase) counters and builds UMMA descriptors by 32-bit adds on a
//     precomputed low word instead of `g % 5`, `g / 5`, `g % 3` and a full descriptor rebuild per block.
//   * Row statistics (lse, delta) of the next tile are requested a tile ahead.
//   * dQ leaves through a swizzled smem staging tile and one TMA store instead of 16-byte stores to rows 1.5 MB apart.
//
// Tried and measured slower (round 2, same box): a 128-KEY-BLOCK variant (SS scores at N = 128, S double- / dP single-buffered
// with early release, K/V ring of two 64 KB stages -- all that fits next to the Q / dO tile and the staging tile): 1.59 ms vs
// 1.31 ms.  Half the MMA batches and polls per FLOP, but the next-but-one block's K/V can only be requested when this block's
// dQ MMA has finished, and that 64 KB load is then exposed on every block (a third stage does not fit in 227 KB).
//
// TMEM map (512 columns):  S[2] @0,64 | dP[2] @128,192 | dQ @256..383 | Q (bf16 pairs) @384..447 | dO @448..511.
// dS_j (packed bf16) aliases its own S buffer: keys 0..31 -> columns +0..15, keys 32..63 -> columns +32..47 (each half is
// written by the warps that read exactly those score columns, so no thread overwrites scores another thread still needs).
#include "attention_bwd_common.cuh"
#include "dropout.cuh"

#ifdef PFN_DQ_TRACE
#define DQ_LOG(tr, ...) (tr).log(__VA_ARGS__)
#else
#define DQ_LOG(tr, ...) ((void)0)
#endif

namespace pfn {

constexpr int DQ_KS = 4;                                   // K/V block ring depth
constexpr int DQ_THREADS = AB_THREADS + 32;                // + warp 10: issuer of the dQ-accumulate MMAs
constexpr int DQ_SMEM_Q = 0;
constexpr int DQ_SMEM_DO = AB_TILE_BYTES;
constexpr int DQ_SMEM_KV = 2 * AB_TILE_BYTES;              // stage s: K at +s*32K, V at +16K
constexpr int DQ_SMEM_OUT = DQ_SMEM_KV + DQ_KS * 2 * AB_BLK_BYTES;   // dQ staging tile, 32 KB
constexpr int DQ_SMEM_BARS = DQ_SMEM_OUT + AB_TILE_BYTES;
constexpr int DQ_SMEM_SELF = DQ_SMEM_BARS + 512;           // [2][64] float2: (dS_ii, P_ii) of a diagonal block's rows
constexpr int DQ_SMEM = DQ_SMEM_SELF + 1024 + 1024;        // + 1 KB alignment slack

constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_Q = 384, TM_DO = 448;

// low / high 32-bit words of the sm_100 shared-memory matrix descriptor (128-byte swizzle, SBO = 1024 B)
constexpr uint32_t kDescHi = static_cast<uint32_t>(tc::umma_smem_desc_hi(1024) >> 32);
__device__ __forceinline__ uint64_t dq_desc(uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; }
// K-major operand (LBO = 16 B), MN-major operand (LBO = 8192 B = next 64-wide chunk); addr16 = smem byte address >> 4
__device__ __forceinline__ uint32_t dq_lo_kmajor(uint32_t addr16) { return ((16u >> 4) << 16) | addr16; }
__device__ __forceinline__ uint32_t dq_lo_mnmajor(uint32_t addr16) { return ((8192u >> 4) << 16) | addr16; }

__global__ void __launch_bounds__(DQ_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDQ,
                      const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET on the __shared__ symbol (an integer round trip of the pointer makes every access through it a
  // generic LD.E / ST.E instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + DQ_SMEM_Q;
  uint8_t* sDO = smem + DQ_SMEM_DO;
  uint8_t* sKV = smem + DQ_SMEM_KV;
  uint8_t* sOut = smem + DQ_SMEM_OUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_SMEM_BARS);
  float2* sSelf = reinterpret_cast<float2*>(smem + DQ_SMEM_SELF);
  uint64_t* qdo_full = bars + 0;                 // TMA: Q / dO tile landed
  uint64_t* qdo_free = bars + 1;                 // row warps: smem Q / dO tile no longer read
  uint64_t* qt_ready = bars + 2;                 // row warps: Q / dO copied into TMEM
  uint64_t* dq_done = bars + 3;                  // MMA: last dQ MMA of the tile complete
  uint64_t* dq_empty = bars + 4;                 // row warps: dQ accumulator read out
  uint64_t* kv_full = bars + 5;                  // [DQ_KS]
  uint64_t* kv_empty = kv_full + DQ_KS;          // [DQ_KS]
  uint64_t* s_full = kv_empty + DQ_KS;           // [2]
  uint64_t* ds_ready = s_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ds_ready + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int E = p.H * AB_DH;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQKV128);
    tc::tma_prefetch_desc(&tmQKV64);
    tc::tma_prefetch_desc(&tmDO128);
    tc::tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(qdo_full, 1);
    tc::mbar_init(qdo_free, AB_EW_WARPS);
    tc::mbar_init(qt_ready, AB_EW_WARPS);
    tc::mbar_init(dq_done, 1);
    tc::mbar_init(dq_empty, AB_EW_WARPS);
    for (int s = 0; s < DQ_KS; ++s) {
      tc::mbar_init(&kv_full[s], 1);
      tc::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&s_full[s], 1);
      tc::mbar_init(&ds_ready[s], AB_EW_WARPS);
    }
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

  if (warp == 0) {
    // =============================================================== TMA producer
    int stage = 0;
    uint32_t sphase = 0;
    uint32_t tcount = 0;
    tc::KernelTrace tr = tc::trace_make(p.trace, p.trace_cap, 0);
    auto load_qdo = [&](int w, uint32_t tc_) {
      const int bh = w / p.n_tiles;
      const int qt = w - bh * p.n_tiles;
      const int b = bh / p.H, h = bh - b * p.H;
      if (tc_ > 0) tc::mbar_wait(qdo_free, (tc_ - 1) & 1);       // the previous tile's rows are done with the smem tile
      if (lane == 0) DQ_LOG(tr, 1, tc_, 0);
      if (tc::elect_one()) {
        tc::mbar_expect_tx(qdo_full, 2 * AB_TILE_BYTES);
        tc::tma_load_3d(sQ, &tmQKV128, qdo_full, h * AB_DH, b, qt * 128);
        tc::tma_load_3d(sQ + 16384, &tmQKV128, qdo_full, h * AB_DH + 64, b, qt * 128);
        tc::tma_load_3d(sDO, &tmDO128, qdo_full, h * AB_DH, b, qt * 128);
        tc::tma_load_3d(sDO + 16384, &tmDO128, qdo_full, h * AB_DH + 64, b, qt * 128);
      }
      __syncwarp();
    };
    if (static_cast<int>(blockIdx.x) < p.total_work) load_qdo(blockIdx.x, 0);
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
      const int bh = w / p.n_tiles;
      const int qt = w - bh * p.n_tiles;
      const int b = bh / p.H, h = bh - b * p.H;
      const int i0 = qt * 128;
      int dstart[2];
      const int nb = ab_tile_block_plan(i0, p.sep, p.T, nblk, dstart);
      const int nd = nb - nblk;
      // the NEXT tile's Q / dO are requested once the loads of this tile's diagonal blocks (+1 dense block) are out: the
      // row warps release the smem tile after the last diagonal block, for which they only need data that has already been
      // requested, so waiting for `qdo_free` here cannot deadlock
      const int jq = min(nb - 1, nd);      // >= nd - 1: every diagonal block's K / V has been requested before we wait
      for (int j = 0; j < nb; ++j) {
        tc::mbar_wait(&kv_empty[stage], sphase ^ 1);
        uint8_t* kdst = sKV + stage * 2 * AB_BLK_BYTES;
        uint8_t* vdst = kdst + AB_BLK_BYTES;
        const int j0 = j < nd ? dstart[j] : (j - nd) * 64;
        if (lane == 0) DQ_LOG(tr, 2, tcount, j);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&kv_full[stage], 2 * AB_BLK_BYTES);
          tc::tma_load_3d(kdst, &tmQKV64, &kv_full[stage], E + h * AB_DH, b, j0);
          tc::tma_load_3d(kdst + 8192, &tmQKV64, &kv_full[stage], E + h * AB_DH + 64, b, j0);
          tc::tma_load_3d(vdst, &tmQKV64, &kv_full[stage], 2 * E + h * AB_DH, b, j0);
          tc::tma_load_3d(vdst + 8192, &tmQKV64, &kv_full[stage], 2 * E + h * AB_DH + 64, b, j0);
        }
        __syncwarp();
        if (++stage == DQ_KS) { stage = 0; sphase ^= 1; }
        if (j == jq && w + static_cast<int>(gridDim.x) < p.total_work) load_qdo(w + gridDim.x, tcount + 1);
      }
    }
  } else if (warp == 1 || warp == 10) {
    // =============================================================== MMA issuers (converged warps, one elected lane issues)
    // TWO issuing warps: warp 1 issues the score batches (S_j, dP_j), warp 10 the dQ-accumulate batches.  With a single
    // issuer the tcgen05 queue is shallow enough that every issue blocks until the pipe has nearly caught up, so one thread
    // served, in series, the score issue (~610 clk), the wait for dS, the accumulate issue (~540 clk) and the barrier
    // polls (~120 clk each on a complete barrier): 1 830 clk per 64-key block against a pipe floor of 768
    // (profiles/r2_trace_attn_bwd_dq_clock64.txt).  Split, the two issue streams and their polls overlap; the only ordering
    // the single thread provided implicitly -- the scores of block n+2 overwrite the S / dP buffers whose dS aliases feed
    // the dQ MMA of block n -- is now the kv_empty barrier of block n (committed behind that MMA), which warp 1 polls.
    // (tools/ubench/mma_gap.cu: descriptor arithmetic, tcgen05.commit and tcgen05.fence are free; all lanes poll.)
    const uint32_t kv16 = tc::smem_u32(sKV) >> 4;
    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_q = tc::umma_idesc_bf16(128, 128, 0, 1);
    int stage = 0;            // ring position (= score buffer parity) of the block this warp handles next
    uint32_t sphase = 0;      // phase of kv_full[stage] / kv_empty[stage]
    uint32_t tcount = 0;
    tc::KernelTrace tr = tc::trace_make(p.trace, p.trace_cap, warp == 1 ? 1 : 10);
    if (warp == 1) {
      uint32_t n = 0;         // blocks issued so far
      int st2 = 0;            // ring position / phase of block n - 2
      uint32_t ph2 = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        const int qt = w % p.n_tiles;
        int dstart[2];
        const int nb = ab_tile_block_plan(qt * 128, p.sep, p.T, nblk, dstart);
        tc::mbar_wait(qt_ready, tcount & 1);          // the row warps have copied this tile's Q / dO into TMEM
        for (int j = 0; j < nb; ++j, ++n) {
          tc::mbar_wait(&kv_full[stage], sphase);
          if (n >= 2) {                                // the dQ MMA of block n-2 no longer reads the dS over this S buffer
            tc::mbar_wait(&kv_empty[st2], ph2);
            if (++st2 == DQ_KS) { st2 = 0; ph2 ^= 1; }
          }
          tc::tc_fence_after();
          if (lane == 0) DQ_LOG(tr, 11, tcount, j);
          const uint32_t sb = stage & 1;
          const uint32_t k16 = kv16 + static_cast<uint32_t>(stage) * (2 * AB_BLK_BYTES >> 4);
          const uint32_t v16 = k16 + (AB_BLK_BYTES >> 4);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              tc::umma_bf16_ts(tmem_base + TM_S + sb * 64, tmem_base + TM_Q + kk * 8,
                               dq_desc(dq_lo_kmajor(k16 + (kk >> 2) * (8192 >> 4) + (kk & 3) * 2)), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              tc::umma_bf16_ts(tmem_base + TM_DP + sb * 64, tmem_base + TM_DO + kk * 8,
                               dq_desc(dq_lo_kmajor(v16 + (kk >> 2) * (8192 >> 4) + (kk & 3) * 2)), idesc_s, kk > 0 ? 1u : 0u);
            tc::umma_commit(&s_full[sb]);
          }
          __syncwarp();
          if (lane == 0) DQ_LOG(tr, 14, tcount, j);
          if (++stage == DQ_KS) { stage = 0; sphase ^= 1; }
        }
      }
    } else {
      uint32_t aphase = 0;    // phase of ds_ready[stage & 1]: toggles every second block
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
        const int qt = w % p.n_tiles;
        int dstart[2];
        const int nb = ab_tile_block_plan(qt * 128, p.sep, p.T, nblk, dstart);
        for (int j = 0; j < nb; ++j) {
          tc::mbar_wait(&ds_ready[stage & 1], aphase);
          if (j == 0 && tcount > 0) tc::mbar_wait(dq_empty, (tcount - 1) & 1);   // previous tile's dQ read out of TMEM
          tc::tc_fence_after();
          if (lane == 0) DQ_LOG(tr, 12, tcount, j);
          const uint32_t k16 = kv16 + static_cast<uint32_t>(stage) * (2 * AB_BLK_BYTES >> 4);
          if (tc::elect_one()) {
            const uint32_t a0 = tmem_base + TM_S + (stage & 1) * 64;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc::umma_bf16_ts(tmem_base + TM_DQ, a0 + (kk >> 1) * 32 + (kk & 1) * 8,
                               dq_desc(dq_lo_mnmajor(k16 + kk * (2048 >> 4))), idesc_q, (j > 0 || kk > 0) ? 1u : 0u);
            tc::umma_commit(&kv_empty[stage]);
            if (j + 1 == nb) tc::umma_commit(dq_done);
          }
          __syncwarp();
          if (lane == 0) DQ_LOG(tr, 13, tcount, j);
          if (stage & 1) aphase ^= 1;
          if (++stage == DQ_KS) { stage = 0; sphase ^= 1; }
        }
      }
    }
  } else {
    // =============================================================== row warps: one thread per (query row, 32-key half)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;             // which 32 of a block's 64 key columns / which 64 of the 128 dh columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const bool issuer = (warp == 2 && lane == 0);  // owns the dQ bulk-store groups
    bool store_pending = false;
    uint32_t buf = 0, bphase = 0;
    uint32_t tcount = 0;
    tc::KernelTrace tr = tc::trace_make(p.trace, p.trace_cap, warp);   // one region per row warp (2..9)
    // copy this thread's 64 dh columns of its Q and dO rows from the swizzled smem tiles into TMEM (raw bf16 pairs)
    auto copy_to_tmem = [&](uint32_t parity) {
      tc::mbar_wait(qdo_full, parity);
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint8_t* src = (which ? sDO : sQ) + half * 16384 + row * 128;
        uint32_t v[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 pk = *reinterpret_cast<const uint4*>(src + ((u ^ (row & 7)) << 4));
          v[4 * u] = pk.x; v[4 * u + 1] = pk.y; v[4 * u + 2] = pk.z; v[4 * u + 3] = pk.w;
        }
        tc::tmem_st_32x32b_x32(tmem_base + lane_off + (which ? TM_DO : TM_Q) + half * 32, v);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive_warp(qt_ready);
    };
    auto load_stats = [&](int w, float& lse_raw, float& delta_raw, bool& ok) {
      ok = false; lse_raw = 0.f; delta_raw = 0.f;
      if (w < p.total_work) {
        const int bh = w / p.n_tiles;
        const int i = (w - bh * p.n_tiles) * 128 + row;
        if (i < p.T) {
          ok = true;
          lse_raw = __ldg(p.lse + static_cast<size_t>(bh) * p.T + i);
          delta_raw = p.delta_tm ? __ldg(p.delta + (static_cast<size_t>(i) * p.B + bh / p.H) * p.H + bh % p.H)
                                 : __ldg(p.delta + static_cast<size_t>(bh) * p.T + i);
        }
      }
    };
    float lse_raw, delta_raw;
    bool stat_ok;
    load_stats(blockIdx.x, lse_raw, delta_raw, stat_ok);
    if (static_cast<int>(blockIdx.x) < p.total_work) copy_to_tmem(0);
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcount) {
      const int bh = w / p.n_tiles;
      const int qt = w - bh * p.n_tiles;
      const int b = bh / p.H, h = bh - b * p.H;
      const int i0 = qt * 128;
      const int i = i0 + row;
      const bool valid = i < p.T;
      const bool is_query = valid && i >= p.sep;
      int dstart[2];
      const int nb = ab_tile_block_plan(i0, p.sep, p.T, nblk, dstart);
      const int nd = nb - nblk;
      const float lse2 = stat_ok ? lse_raw * 1.4426950408889634f : INFINITY;
      const float dls = stat_ok ? delta_raw * p.scale : 0.f;
      // statistics of the NEXT tile: in flight during this whole tile
      float nlse, ndelta;
      bool nok;
      load_stats(w + gridDim.x, nlse, ndelta, nok);
      if (nd == 0) tc::mbar_arrive_warp(qdo_free);           // no diagonal block: the smem Q / dO tile is already dead
      for (int j = 0; j < nb; ++j) {
        tc::mbar_wait_rows(&s_full[buf], bphase);
        if (lane == 0) DQ_LOG(tr, 20 + 100 * warp, tcount, j);
        tc::tc_fence_after();
        const bool diag = j < nd;
        const int kmax = diag ? 0 : p.sep - (j - nd) * 64;
        uint32_t s[32], dp[32], pk[16];
        float dself = 0.f, pself = 0.f;     // diagonal block: dS_ii and P_ii (x dropout factor) of this thread's own key
        bool has_self = false;
        tc::tmem_ld_32x32b_x32(tmem_base + lane_off + TM_S + buf * 64 + half * 32, s);
        tc::tmem_ld_32x32b_x32(tmem_base + lane_off + TM_DP + buf * 64 + half * 32, dp);
        tc::tmem_ld_wait();
        if (lane == 0) DQ_LOG(tr, 25 + 100 * warp, tcount, j);
        if (!diag && p.drop_thr > 0) {
          // dropout on the probabilities (csrc/dropout.cuh): dS = P (m dP / (1 - p) - delta) scale, keep bit of key j = byte j & 3
          // of the hash of (row id (b*H + h)*T + i, j >> 2); covers full and partial dense blocks
          const float dsc = drop_scale(p.drop_thr);
          const uint32_t rid = static_cast<uint32_t>(bh) * p.T + i;
          const int kb = (j - nd) * 64 + half * 32;           // first key of this thread's 32 columns
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const uint32_t hsh = drop_hash(p.drop_seed, rid, static_cast<uint32_t>(kb >> 2) + q4);
            float dv4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * q4 + e;
              const float mk = drop_keep_byte(hsh, e, p.drop_thr) ? dsc : 0.f;
              const float pr = tc::fast_exp2(fmaf(__uint_as_float(s[c]), p.scale_log2, -lse2));
              dv4[e] = (half * 32 + c < kmax) ? pr * fmaf(__uint_as_float(dp[c]) * mk, p.scale, -dls) : 0.f;
            }
            pk[2 * q4] = tc::pack_bf16x2(dv4[0], dv4[1]);
            pk[2 * q4 + 1] = tc::pack_bf16x2(dv4[2], dv4[3]);
          }
        } else         if (!diag && kmax >= 64) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float p0 = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c]), p.scale_log2, -lse2));
            const float p1 = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), p.scale_log2, -lse2));
            pk[c] = tc::pack_bf16x2(p0 * fmaf(__uint_as_float(dp[2 * c]), p.scale, -dls),
                                    p1 * fmaf(__uint_as_float(dp[2 * c + 1]), p.scale, -dls));
          }
        } else if (!diag) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int k0 = half * 32 + 2 * c;
            float d0 = 0.f, d1 = 0.f;
            if (k0 < kmax)
              d0 = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c]), p.scale_log2, -lse2)) * fmaf(__uint_as_float(dp[2 * c]), p.scale, -dls);
            if (k0 + 1 < kmax)
              d1 = tc::fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), p.scale_log2, -lse2)) * fmaf(__uint_as_float(dp[2 * c + 1]), p.scale, -dls);
            pk[c] = tc::pack_bf16x2(d0, d1);
          }
        } else {
          int cl = -1;                                    // own column inside this warp's half, if any
          {
            const int c = i - dstart[j] - half * 32;
            if (is_query && c >= 0 && c < 32) cl = c;
          }
          float sv = 0.f, dv = 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            sv = (c == cl) ? __uint_as_float(s[c]) : sv;
            dv = (c == cl) ? __uint_as_float(dp[c]) : dv;
          }
          if (cl >= 0) {
            // the diagonal key is attended by this row only: dK_i = dS_ii q_i and dV_i = P_ii dO_i are complete (written
            // below, after dS has been published)
            const float pii = tc::fast_exp2(fmaf(sv, p.scale_log2, -lse2));
            float mk = 1.f;                                 // dropout keep factor of the diagonal key (column i of row i)
            if (p.drop_thr > 0) mk = drop_keep(p.drop_seed, static_cast<uint32_t>(bh) * p.T + i, static_cast<uint32_t>(i), p.drop_thr) ? drop_scale(p.drop_thr) : 0.f;
            dself = pii * fmaf(dv * mk, p.scale, -dls);
            pself = pii * mk;
            has_self = true;
          }
          {
            // every row of this 64-row diagonal block has one owner thread (the half whose 32 columns hold the row's own
            // key); it leaves (dS_ii, P_ii) -- or a NaN marker for a row without a diagonal key -- for the warps that write
            // dK_i / dV_i below
            const int c64 = i0 + row - dstart[j];
            if (c64 >= 0 && c64 < 64 && (c64 >> 5) == half)
              sSelf[(j & 1) * 64 + c64] = make_float2(has_self ? dself : __int_as_float(0x7fc00000), pself);
          }
          const uint32_t lo = tc::pack_bf16x2(dself, 0.f), hi = tc::pack_bf16x2(0.f, dself);
          const int cw = cl >> 1;                         // -1 >> 1 == -1: matches nothing
          const uint32_t word = (cl & 1) ? hi : lo;
#pragma unroll
          for (int c = 0; c < 16; ++c) pk[c] = (c == cw) ? word : 0u;
        }
        if (lane == 0) DQ_LOG(tr, 26 + 100 * warp, tcount, j);
        tc::tmem_st_32x32b_x16(tmem_base + lane_off + TM_S + buf * 64 + half * 32, pk);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive_warp(&ds_ready[buf]);
        if (lane == 0) DQ_LOG(tr, 21 + 100 * warp, tcount, j);
        if (diag) {
          // dK_i = dS_ii q_i, dV_i = P_ii dO_i for the 64 rows of this diagonal block, AFTER dS went out, shared by all 8
          // row warps (8 rows each).  A warp writes one row per step: lanes 0..15 the 256-byte dK row, lanes 16..31 the dV
          // row (16 bytes each, read back from the still-live swizzled smem Q / dO tiles) -- two fully used 256-byte
          // segments per store instruction.  (One thread per row wrote 32 half-used sectors per instruction, only the 2
          // warps owning the block's rows worked, and the tile stood still for ~7 000 clocks per diagonal block:
          // profiles/r2_trace_attn_bwd_dq_clock64.txt.)
          asm volatile("bar.sync 1, 256;" ::: "memory");          // sSelf of this block complete
          const int which = lane >> 4, u = lane & 15;
          const uint8_t* tile = (which ? sDO : sQ) + (u >> 3) * 16384;
          const int rbase = dstart[j] - i0;                        // tile row of the block's first row (0 or 64)
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int c64 = (warp - 2) * 8 + r8;
            const float2 sp = sSelf[(j & 1) * 64 + c64];           // broadcast
            if (sp.x == sp.x) {                                    // warp-uniform: the row has a diagonal key
              const float f = which ? sp.y : sp.x;
              const int rr = rbase + c64;
              const int ii = i0 + rr;
              const uint4 raw = *reinterpret_cast<const uint4*>(tile + rr * 128 + (((u & 7) ^ (rr & 7)) << 4));
              const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
              uint4 o;
              float2 t = __bfloat1622float2(hh[0]); o.x = tc::pack_bf16x2(t.x * f, t.y * f);
              t = __bfloat1622float2(hh[1]); o.y = tc::pack_bf16x2(t.x * f, t.y * f);
              t = __bfloat1622float2(hh[2]); o.z = tc::pack_bf16x2(t.x * f, t.y * f);
              t = __bfloat1622float2(hh[3]); o.w = tc::pack_bf16x2(t.x * f, t.y * f);
              const size_t tokq = p.batch_major ? static_cast<size_t>(b) * p.T + ii : static_cast<size_t>(ii) * p.B + b;
              *reinterpret_cast<uint4*>(p.dqkv + tokq * p.ld_dqkv + (which ? 2 * E : E) + h * AB_DH + u * 8) = o;
            }
          }
        }
        if (j + 1 == nd) tc::mbar_arrive_warp(qdo_free);     // last diagonal block: smem Q / dO rows no longer needed
        buf ^= 1;
        if (buf == 0) bphase ^= 1;
      }
      // ---- all score MMAs of this tile are complete (this thread saw the last s_full): the TMEM copies of Q / dO may be
      //      replaced by the next tile's, whose first score MMAs then overlap the epilogue below
      if (w + static_cast<int>(gridDim.x) < p.total_work) copy_to_tmem((tcount + 1) & 1);
      // ---- epilogue: dQ -> bf16 -> swizzled smem staging -> one TMA store per 64-column chunk
      if (lane == 0) DQ_LOG(tr, 24 + 100 * warp, tcount, 0);      // next tile's TMEM copy done
      tc::mbar_wait(dq_done, tcount & 1);
      if (lane == 0) DQ_LOG(tr, 22 + 100 * warp, tcount, 0);      // epilogue start
      tc::tc_fence_after();
      if (store_pending) {                                   // the previous tile's bulk store must have read the staging tile
        if (issuer) tc::tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        store_pending = false;
      }
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;                         // 32-column chunk of the 128 dh columns
        uint32_t raw[32];
        tc::tmem_ld_32x32b_x32(tmem_base + lane_off + TM_DQ + c * 32, raw);
        tc::tmem_ld_wait();
        uint8_t* rowp = sOut + (c >> 1) * 16384 + row * 128;
        const int u0 = (c & 1) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = tc::pack_bf16x2(__uint_as_float(raw[8 * q]), __uint_as_float(raw[8 * q + 1]));
          o.y = tc::pack_bf16x2(__uint_as_float(raw[8 * q + 2]), __uint_as_float(raw[8 * q + 3]));
          o.z = tc::pack_bf16x2(__uint_as_float(raw[8 * q + 4]), __uint_as_float(raw[8 * q + 5]));
          o.w = tc::pack_bf16x2(__uint_as_float(raw[8 * q + 6]), __uint_as_float(raw[8 * q + 7]));
          *reinterpret_cast<uint4*>(rowp + (((u0 + q) ^ (row & 7)) << 4)) = o;
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive_warp(dq_empty);                        // the dQ accumulator may be overwritten by the next tile
      tc::fence_proxy_async_smem();                          // generic-proxy staging writes -> visible to the TMA
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (issuer) {
        tc::tma_store_3d(&tmDQ, sOut, h * AB_DH, b, i0);     // rows >= T are clipped by the tensor map
        tc::tma_store_3d(&tmDQ, sOut + 16384, h * AB_DH + 64, b, i0);
        tc::tma_store_commit();
      }
      store_pending = true;
      if (p.dq_colsum != nullptr) {
        // q third of the in-projection bias gradient: column sums of the staged bf16 tile (rows >= T hold zeros), two
        // threads per column, 64 rows each -- instead of a separate pass re-reading dqkv from HBM
        const int tid = threadIdx.x - 64;                    // 0..255
        const int c = tid & 127, rh = tid >> 7;
        const uint32_t colp = tc::smem_u32(sOut) + (c >> 6) * 16384 + (c & 7) * 2;   // explicit shared-space loads
        const int u = (c & 63) >> 3;
        float acc = 0.f;
#pragma unroll 8
        for (int r = rh * 64; r < rh * 64 + 64; ++r) {
          unsigned short bits;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(bits) : "r"(colp + r * 128 + ((u ^ (r & 7)) << 4)));
          acc += __uint_as_float(static_cast<uint32_t>(bits) << 16);
        }
        atomicAdd(p.dq_colsum + h * AB_DH + c, acc);
      }
      if (lane == 0) DQ_LOG(tr, 23 + 100 * warp, tcount, 0);      // epilogue end
      lse_raw = nlse; delta_raw = ndelta; stat_ok = nok;
    }
    if (issuer && store_pending) tc::tma_store_wait<0>();    // all bulk stores complete before the CTA exits
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

static int dq_make_map3d(CUtensorMap* tm, const void* base, int ld, int width, int B, int T, int box_rows, int batch_major) {
  uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(B), static_cast<uint64_t>(T)};
  uint64_t strides[3] = {0, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * B};
  if (batch_major) { strides[1] = static_cast<uint64_t>(ld) * 2 * T; strides[2] = static_cast<uint64_t>(ld) * 2; }
  uint32_t box[3] = {64, 1, static_cast<uint32_t>(box_rows)};
  return make_tensor_map_bf16(tm, base, 3, dims, strides, box, true);
}

int launch_attn_bwd_dq(const AttnBwdParams& p_in, const pfn_attn_desc* d, cudaStream_t stream) {
  const int E = d->H * d->dh;
  CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDQ;
  if (int rc = dq_make_map3d(&tmQKV128, d->qkv, d->ld_qkv, 3 * E, d->B, d->T, 128, d->batch_major)) return rc;
  if (int rc = dq_make_map3d(&tmQKV64, d->qkv, d->ld_qkv, 3 * E, d->B, d->T, 64, d->batch_major)) return rc;
  if (int rc = dq_make_map3d(&tmDO128, d->dout, d->ld_dout, E, d->B, d->T, 128, d->batch_major)) return rc;
  // dQ occupies columns [0, E) of dqkv; declaring only those keeps a stray store from ever touching dK / dV
  if (int rc = dq_make_map3d(&tmDQ, d->dqkv, d->ld_dqkv, E, d->B, d->T, 128, d->batch_major)) return rc;
  AttnBwdParams p = p_in;
  p.n_tiles = (d->T + 127) / 128;
  p.total_work = p.n_tiles * d->B * d->H;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set))
    PFN_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
  const int grid = num_sms() < p.total_work ? num_sms() : p.total_work;
  attn_bwd_dq_tc_kernel<<<grid, DQ_THREADS, DQ_SMEM, stream>>>(tmQKV128, tmQKV64, tmDO128, tmDQ, p);
  PFN_LAUNCH_OK();
  return 0;
}

}  // namespace pfn
