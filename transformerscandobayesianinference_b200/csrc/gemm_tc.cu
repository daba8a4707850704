// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] (+)= epilogue( sum_k A(m,k) * B(n,k) )      bf16 operands, fp32 accumulation in TMEM
//
// Replaces the cuBLASLt addmm calls the reference reaches through nn.Linear / in_proj / out_proj
// (reference transformer.py:17-18,23,84-85; torch nn/functional.py:6478 `_in_projection_packed`).
//
// Operand storage ("major"):
//   K-major  : element (i,k) at base + i*ld + k        (activations x, weights W[N,K] for y = x W^T)
//   MN-major : element (i,k) at base + k*ld + i        (W used for dgrad, dY / X used for wgrad)
// so forward, dgrad and wgrad all run on the same kernel without any transposed copies in HBM.
//
// Structure (one CTA per SM, persistent over work items = output tile x k-split):
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma 128 x BLOCK_N x 16, commits to mbarriers)
//   warps 2..9  : epilogue       (tcgen05.ld TMEM->regs, bias / GELU / residual / GELU'; bf16 tiles are staged in
//                                 128B-swizzled smem and written with TMA stores, fp32/atomic outputs go direct)
//   TMEM        : 2 accumulator stages x BLOCK_N fp32 columns (epilogue of tile i overlaps MMA of i+1)
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/pfn_b200.h"

namespace pfn {

struct GemmTcParams {
  int M, N, K;
  const float* bias;             // [N] fp32 or null
  const __nv_bfloat16* aux;      // residual (epi add) or pre-activation u (GELU') : [M, ld_aux] bf16, or null
  int ld_aux;
  void* C;                       // bf16 or fp32 [M, ldc]
  int ldc;
  int c_f32;                     // 1 => C is fp32
  __nv_bfloat16* C2;             // optional second output: pre-activation (only with act == GELU)
  int ldc2;
  int act;                       // PFN_EPI_*
  int accumulate;                // 1 => atomically add into fp32 C (split-K / grad accumulation)
  int k_splits;
  int kb_per_split;              // k-blocks (of 64) per split
  int tiles_m, tiles_n;
  int tma_store;                 // 1 => bf16 C (and C2) leave through the smem staging buffer + TMA store
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = 64 + kNumEpiWarps * 32;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStagesMax = BLOCK_N == 256 ? 4 : 6;      // fp32 / atomic outputs: no staging buffer, deeper ring
  static constexpr int kStagesStaged = BLOCK_N == 256 ? 3 : 5;   // bf16 outputs: one stage gives way to the staging buffer
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kStageOutBytes = kBlockM * BLOCK_N * 2;   // bf16 output tile (two column halves, one per epilogue group)
  static constexpr int kRingBytes = kStagesMax * (kABytes + kBBytes);
  static constexpr int kStagedBytes = kStagesStaged * (kABytes + kBBytes) + kStageOutBytes;
  static constexpr int kDataBytes = kRingBytes > kStagedBytes ? kRingBytes : kStagedBytes;
  static constexpr int kSmemBytes = kDataBytes + 256 + 1024;
};

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const GemmTcParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  const int STAGES = p.tma_store ? Cfg::kStagesStaged : Cfg::kStagesMax;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::kABytes;
  uint8_t* sOut = smem + STAGES * (Cfg::kABytes + Cfg::kBBytes);      // only used when p.tma_store
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kDataBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStagesMax;
  uint64_t* tfull_bar = empty_bar + Cfg::kStagesMax;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    if (p.tma_store) { tc::tma_prefetch_desc(&tmC); if (p.C2 != nullptr) tc::tma_prefetch_desc(&tmC2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tfull_bar[s], 1);
      tc::mbar_init(&tempty_bar[s], kNumEpiWarps);
    }
    tc::mbar_fence_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb_total = (p.K + kBlockK - 1) / kBlockK;
  const int tiles = p.tiles_m * p.tiles_n;
  const int total_work = tiles * p.k_splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w / tiles;
        const int tile = w - split * tiles;
        const int m0 = (tile / p.tiles_n) * kBlockM;
        const int n0 = (tile % p.tiles_n) * BLOCK_N;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = sA + stage * Cfg::kABytes;
          uint8_t* b_dst = sB + stage * Cfg::kBBytes;
          const int k0 = kb * kBlockK;
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full_bar[stage], Cfg::kABytes + Cfg::kBBytes);
            if constexpr (!A_MN) {
              tc::tma_load_2d(a_dst, &tmA, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < kBlockM / 64; ++c)
                tc::tma_load_2d(a_dst + c * 8192, &tmA, &full_bar[stage], m0 + c * 64, k0);
            }
            if constexpr (!B_MN) {
              tc::tma_load_2d(b_dst, &tmB, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BLOCK_N / 64; ++c)
                tc::tma_load_2d(b_dst + c * 8192, &tmB, &full_bar[stage], n0 + c * 64, k0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kBlockM, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
        const int split = w / tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, num_kb_total);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        tc::mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(sA + stage * Cfg::kABytes);
          const uint32_t b_addr = tc::smem_u32(sB + stage * Cfg::kBBytes);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // K-major: advance 16 elements (32 B) inside the 128 B swizzle row.
              // MN-major: advance 16 k-rows (16 * 128 B); LBO = stride between 64-wide M/N chunks.
              const uint64_t a_desc = A_MN ? tc::umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                           : tc::umma_smem_desc(a_addr + k * 32, 16, 1024);
              const uint64_t b_desc = B_MN ? tc::umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                           : tc::umma_smem_desc(b_addr + k * 32, 16, 1024);
              tc::umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            tc::umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) tc::umma_commit(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Two groups of four warps; group `half` owns the left / right BLOCK_N/2 columns of the tile and its own staging
    // buffer ([128 rows x 64 cols] bf16 chunks in the TMA 128-byte swizzle).
    const int ew = warp - 2;
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // column half
    constexpr int COLS_PER_GROUP = BLOCK_N / 2;
    constexpr int OUT_CHUNKS = COLS_PER_GROUP / 64;
    uint8_t* stg = sOut + half * (Cfg::kStageOutBytes / 2);
    const bool issuer = (ew & 3) == 0 && lane == 0;       // the one thread per group that owns the bulk-store groups
    const int nbar = 1 + half;                             // named barrier of this group (0 is __syncthreads)
    const bool two_pass = p.tma_store && p.act == PFN_EPI_GELU && p.C2 != nullptr;
    bool store_pending = false;
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int split = w / tiles;
      const int tile = w - split * tiles;
      const int m0 = (tile / p.tiles_n) * kBlockM;
      const int n0 = (tile % p.tiles_n) * BLOCK_N;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      tc::mbar_wait(&tfull_bar[as], aphase);
      tc::tc_fence_after();
      const int trow = q * 32 + lane;
      const int row = m0 + trow;
      const bool row_ok = row < p.M;
      const bool add_bias = p.bias != nullptr && split == 0;
      const int npass = two_pass ? 2 : 1;
#pragma unroll 1
      for (int pass = 0; pass < npass; ++pass) {
        // pass 0 of a two-pass tile writes the pre-activation (C2), the last pass writes C
        const bool write_pre_only = two_pass && pass == 0;
        if (p.tma_store && store_pending) {
          // the previous bulk store must have finished READING the staging buffer before it is overwritten
          if (issuer) tc::tma_store_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(nbar) : "memory");
          store_pending = false;
        }
#pragma unroll 1
        for (int cc = 0; cc < COLS_PER_GROUP; cc += 32) {
          const int col0 = n0 + half * COLS_PER_GROUP + cc;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(as * BLOCK_N + half * COLS_PER_GROUP + cc);
          tc::tmem_ld_32x32b_x32(taddr, v);
          tc::tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          const bool full_chunk = (col0 + 32 <= p.N);
          if (add_bias) {
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
                f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) f[i] += __ldg(p.bias + col0 + i);
            }
          }
          if (p.act == PFN_EPI_GELU && !write_pre_only) {
            if (p.C2 != nullptr && row_ok && !p.tma_store) {
              __nv_bfloat16* dst = p.C2 + static_cast<size_t>(row) * p.ldc2 + col0;
              if (full_chunk) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  uint4 pk;
                  pk.x = tc::pack_bf16x2(f[i], f[i + 1]); pk.y = tc::pack_bf16x2(f[i + 2], f[i + 3]);
                  pk.z = tc::pack_bf16x2(f[i + 4], f[i + 5]); pk.w = tc::pack_bf16x2(f[i + 6], f[i + 7]);
                  *reinterpret_cast<uint4*>(dst + i) = pk;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) dst[i] = __float2bfloat16_rn(f[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
          }
          if (p.aux != nullptr && row_ok && !write_pre_only && (p.act == PFN_EPI_GELU_BWD || split == 0)) {
            const __nv_bfloat16* src = p.aux + static_cast<size_t>(row) * p.ld_aux + col0;
            float a[32];
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                const uint4 pk = *reinterpret_cast<const uint4*>(src + i);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 t = __bfloat1622float2(h[j]);
                  a[i + 2 * j] = t.x; a[i + 2 * j + 1] = t.y;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) a[i] = (col0 + i < p.N) ? __bfloat162float(src[i]) : 0.f;
            }
            if (p.act == PFN_EPI_GELU_BWD) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] *= gelu_erf_grad(a[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] += a[i];
            }
          }
          if (p.tma_store) {
            // staging write: 64-column chunk (cc / 64), 16-byte units (cc % 64) / 8 .. +3 of row `trow`, XOR-swizzled
            uint8_t* rowp = stg + (cc >> 6) * 16384 + trow * 128;
            const int u0 = (cc & 63) >> 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 pk;
              pk.x = tc::pack_bf16x2(f[8 * i], f[8 * i + 1]); pk.y = tc::pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
              pk.z = tc::pack_bf16x2(f[8 * i + 4], f[8 * i + 5]); pk.w = tc::pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
              *reinterpret_cast<uint4*>(rowp + (((u0 + i) ^ (trow & 7)) << 4)) = pk;
            }
          } else if (row_ok) {
            if (p.c_f32) {
              float* dst = reinterpret_cast<float*>(p.C) + static_cast<size_t>(row) * p.ldc + col0;
              if (p.accumulate) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) atomicAdd(dst + i, f[i]);
              } else if (full_chunk) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) dst[i] = f[i];
              }
            } else {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<size_t>(row) * p.ldc + col0;
              if (full_chunk) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  uint4 pk;
                  pk.x = tc::pack_bf16x2(f[i], f[i + 1]); pk.y = tc::pack_bf16x2(f[i + 2], f[i + 3]);
                  pk.z = tc::pack_bf16x2(f[i + 4], f[i + 5]); pk.w = tc::pack_bf16x2(f[i + 6], f[i + 7]);
                  *reinterpret_cast<uint4*>(dst + i) = pk;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) dst[i] = __float2bfloat16_rn(f[i]);
              }
            }
          }
        }
        if (pass == npass - 1) {
          // accumulator stage drained: hand it back to the MMA warp
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tempty_bar[as]);
        }
        if (p.tma_store) {
          tc::fence_proxy_async_smem();                       // generic-proxy staging writes -> visible to the TMA
          asm volatile("bar.sync %0, 128;" ::"r"(nbar) : "memory");
          if (issuer) {
            const CUtensorMap* tm = write_pre_only ? &tmC2 : &tmC;
#pragma unroll
            for (int c = 0; c < OUT_CHUNKS; ++c) {
              const int cbase = n0 + half * COLS_PER_GROUP + c * 64;
              if (cbase < p.N) tc::tma_store_2d(tm, stg + c * 16384, cbase, m0);   // rows >= M / cols >= N are clipped
            }
            tc::tma_store_commit();
          }
          store_pending = true;
        }
      }
    }
    if (p.tma_store && issuer) tc::tma_store_wait<0>();        // all bulk stores complete before the CTA exits
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN>
static int launch_gemm_tc(const pfn_gemm_desc* d, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[2];
    uint32_t box[2];
    if (!A_MN) { dims[0] = d->K; dims[1] = d->M; box[0] = 64; box[1] = kBlockM; }
    else       { dims[0] = d->M; dims[1] = d->K; box[0] = 64; box[1] = 64; }
    strides[0] = 0; strides[1] = static_cast<uint64_t>(d->lda) * 2;
    if (int rc = make_tensor_map_bf16(&tmA, d->A, 2, dims, strides, box, true)) return rc;
    if (!B_MN) { dims[0] = d->K; dims[1] = d->N; box[0] = 64; box[1] = BLOCK_N; }
    else       { dims[0] = d->N; dims[1] = d->K; box[0] = 64; box[1] = 64; }
    strides[1] = static_cast<uint64_t>(d->ldb) * 2;
    if (int rc = make_tensor_map_bf16(&tmB, d->B, 2, dims, strides, box, true)) return rc;
  }
  CUtensorMap tmC, tmC2;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmC2, 0, sizeof(tmC2));
  // bf16 outputs without accumulation leave through smem staging + TMA store (needs 16 B aligned rows)
  const bool tma_store = d->c_dtype == PFN_BF16 && !d->accumulate && (d->k_splits <= 1) && d->ldc % 8 == 0 &&
                         (d->C2 == nullptr || d->ldc2 % 8 == 0);
  if (tma_store) {
    uint64_t dims[2] = {static_cast<uint64_t>(d->N), static_cast<uint64_t>(d->M)};
    uint64_t strides[2] = {0, static_cast<uint64_t>(d->ldc) * 2};
    uint32_t box[2] = {64, kBlockM};
    if (int rc = make_tensor_map_bf16(&tmC, d->C, 2, dims, strides, box, true)) return rc;
    if (d->C2 != nullptr) {
      strides[1] = static_cast<uint64_t>(d->ldc2) * 2;
      if (int rc = make_tensor_map_bf16(&tmC2, d->C2, 2, dims, strides, box, true)) return rc;
    }
  }
  GemmTcParams p;
  p.tma_store = tma_store ? 1 : 0;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.bias = d->bias;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(d->aux);
  p.ld_aux = d->ld_aux;
  p.C = d->C; p.ldc = d->ldc; p.c_f32 = d->c_dtype == PFN_F32;
  p.C2 = reinterpret_cast<__nv_bfloat16*>(d->C2); p.ldc2 = d->ldc2;
  p.act = d->epilogue;
  p.tiles_m = (d->M + kBlockM - 1) / kBlockM;
  p.tiles_n = (d->N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (d->K + kBlockK - 1) / kBlockK;
  int splits = d->k_splits <= 0 ? 1 : d->k_splits;
  if (splits > num_kb) splits = num_kb;
  const int per = (num_kb + splits - 1) / splits;
  splits = (num_kb + per - 1) / per;
  p.k_splits = splits;
  p.kb_per_split = per;
  p.accumulate = (d->accumulate || splits > 1) ? 1 : 0;
  PFN_CHECK_ARG(!p.accumulate || p.c_f32, "gemm_tc: accumulate / split-K requires an fp32 output");
  const int total = p.tiles_m * p.tiles_n * splits;
  const int grid = total < num_sms() ? total : num_sms();
  auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    PFN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, tmC2, p);
  PFN_LAUNCH_OK();
  return 0;
}

}  // namespace pfn

extern "C" int pfn_gemm_bf16_tc(const pfn_gemm_desc* d, void* stream) {
  using namespace pfn;
  PFN_CHECK_ARG(d != nullptr, "gemm_tc: null descriptor");
  PFN_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm_tc: empty problem %d x %d x %d", d->M, d->N, d->K);
  PFN_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0, "gemm_tc: lda/ldb must be multiples of 8 elements (got %d, %d)",
                d->lda, d->ldb);
  PFN_CHECK_ARG(d->epilogue >= 0 && d->epilogue <= PFN_EPI_GELU_BWD, "gemm_tc: bad epilogue %d", d->epilogue);
  PFN_CHECK_ARG(d->epilogue != PFN_EPI_GELU_BWD || d->aux != nullptr, "gemm_tc: GELU' epilogue needs aux = pre-activation");
  const bool vec_ok = (d->c_dtype == PFN_F32 ? d->ldc % 4 == 0 : d->ldc % 8 == 0) &&
                      (reinterpret_cast<uintptr_t>(d->C) & 15) == 0;
  PFN_CHECK_ARG(vec_ok, "gemm_tc: C must be 16-byte aligned with ldc multiple of 16 bytes");
  PFN_CHECK_ARG(d->aux == nullptr || (d->ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0),
                "gemm_tc: aux must be 16-byte aligned with ld multiple of 8");
  PFN_CHECK_ARG(d->C2 == nullptr || (d->ldc2 % 8 == 0 && (reinterpret_cast<uintptr_t>(d->C2) & 15) == 0),
                "gemm_tc: C2 must be 16-byte aligned with ld multiple of 8");
  PFN_CHECK_ARG(d->bias == nullptr || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "gemm_tc: bias must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool wide = d->N > 128;
  const int key = (wide ? 4 : 0) | (d->a_mn_major ? 2 : 0) | (d->b_mn_major ? 1 : 0);
  switch (key) {
    case 0: return launch_gemm_tc<128, false, false>(d, s);
    case 1: return launch_gemm_tc<128, false, true>(d, s);
    case 2: return launch_gemm_tc<128, true, false>(d, s);
    case 3: return launch_gemm_tc<128, true, true>(d, s);
    case 4: return launch_gemm_tc<256, false, false>(d, s);
    case 5: return launch_gemm_tc<256, false, true>(d, s);
    case 6: return launch_gemm_tc<256, true, false>(d, s);
    default: return launch_gemm_tc<256, true, true>(d, s);
  }
}
