// GP prior sampler: one CTA per dataset builds the kernel matrix on the fly, factors it with a blocked
// left-looking Cholesky in true fp32 (FFMA) and applies the factor to z in the same pass:
//     K = os * k(x, x; ls) + (noise + jitter) I ,   L = chol(K) ,   y = L z
// Restates the sampling maths the reference obtains from gpytorch (priors/fast_gp.py:13-32,48-56:
// ScaleKernel(RBFKernel) + GaussianLikelihood noise, MultivariateNormal.sample -> cholesky root times randn)
// and, with per-dataset hyperparameters and Matern-nu kernels, priors/fast_gp_mix.py:24-55,88-99.
//
// Left-looking, panel width 32:  for panel p (columns c0..c0+31)
//     U[r, :]  = K[r, c0:c0+32] - L[r, 0:c0] L[c0:c0+32, 0:c0]^T          (128x32 tiles; the T^3/3 flops of the factorisation)
//     L11      = chol(U[c0:c0+32, :]),  W = L11^-1                        (one warp, rows / columns in registers, shuffles)
//     L[r, c0:c0+32] = U[r, :] W^T  for r > c0+31                         (a 128x32x32 product for all eight warps)
//     y[r]    += L[r, c0:c0+32] . z[c0:c0+32]
// Both products run on the tensor cores as 3xTF32 (`mma.sync.m16n8k8.tf32`): every fp32 operand is split into hi = tf32(a)
// and lo = tf32(a - hi) and the product is accumulated as lo*hi + hi*lo + hi*hi in fp32, which keeps ~21 bits of each
// operand -- the factor passes the same fp32 residual tests (||L L^T - K|| <= 2e-5 ||K||) as the FFMA version of round 1.
// Why (ncu --set full of the round-1 kernel, profiles/r2_ncu_gp_sample.md): 33 % of the warp samples sat at the barrier
// behind the panel solve -- one thread per row running 496 DEPENDENT FMAs while half the CTA idled -- and 45 % in the FFMA
// update loop at 54 % issue utilisation.  Multiplying by the explicit inverse of the 32x32 diagonal block (formed once per
// panel by the warp that factors it) turns the solve into the same fully parallel product as the update.
// Where the time goes now (ablation builds, 296 datasets of T = 1000 on one B200, profiles/r2_gp_sampler_ablation.md):
// 4.79 ms total = update products 1.9 (the legacy mma.sync path delivers 506 tf32 MAC/clk/SM, 1/8 of tcgen05 -- with the
// three-way split only 1.3x the FFMA rate, tools/ubench/mma_sync_rate.cu) + panel streaming 1.0 + diagonal block 0.9 +
// solve 0.44 + factor stores 0.26 + rest 0.3.  The next step is the update on tcgen05 kind::tf32 with hi/lo smem tiles.
// K is never materialised; the only HBM traffic is writing the factor once and re-reading finished panels.
// The factor is kept TRANSPOSED in `work` (work[b][c][r] = L[r][c]): panel re-reads become contiguous rows that stream
// through a 3-stage cp.async ring straight into the k-major smem tiles, and the panel write-back is coalesced.
#include "common.cuh"
#include "../../include/pfn_b200.h"

namespace pfn {

constexpr int NB = 32;        // panel width
constexpr int TR = 128;       // rows per update tile
constexpr int GP_MAX_F = 128;
constexpr int GP_STAGES = 3;
constexpr int LDR = TR + 8;   // smem row strides (floats) of the k-major ring tiles and of the U tile: +8 / +4 keep the
constexpr int LDC = NB + 8;   // mma.sync fragment loads (4 k-rows x 8 consecutive elements per instruction) conflict-free
constexpr int LDU = NB + 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float gp_kernel_value(float d2, float os, int kernel_type) {
  if (kernel_type == PFN_KERNEL_RBF) return os * expf(-0.5f * d2);
  const float r = sqrtf(d2);
  if (kernel_type == PFN_KERNEL_MATERN12) return os * expf(-r);
  if (kernel_type == PFN_KERNEL_MATERN32) {
    const float a = 1.7320508075688772f * r;
    return os * (1.0f + a) * expf(-a);
  }
  const float a = 2.23606797749979f * r;
  return os * (1.0f + a + (5.0f / 3.0f) * d2) * expf(-a);
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rest = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}
// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// acc[nt] (16 x 8 tiles nt = 0..3 of a 16 x 32 result) += A[16 x 8*KSTEPS] B[8*KSTEPS x 32] as 3xTF32.
//   A(row, k) = Abase[k * lda_k + row * lda_r]   rows = this warp's 16 rows;   B(k, n) = Bbase[k * ldb_k + n * ldb_n]
template <int KSTEPS>
__device__ __forceinline__ void mma_3xtf32_16x32(float (&acc)[4][4], const float* Abase, int lda_k, int lda_r, const float* Bbase,
                                                 int ldb_k, int ldb_n, int fg, int ft) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const int k0 = ks * 8;
    uint32_t ah[4], al[4];
    split_tf32(Abase[(k0 + ft) * lda_k + fg * lda_r], ah[0], al[0]);
    split_tf32(Abase[(k0 + ft) * lda_k + (fg + 8) * lda_r], ah[1], al[1]);
    split_tf32(Abase[(k0 + ft + 4) * lda_k + fg * lda_r], ah[2], al[2]);
    split_tf32(Abase[(k0 + ft + 4) * lda_k + (fg + 8) * lda_r], ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bh[2], bl[2];
      split_tf32(Bbase[(k0 + ft) * ldb_k + (nt * 8 + fg) * ldb_n], bh[0], bl[0]);
      split_tf32(Bbase[(k0 + ft + 4) * ldb_k + (nt * 8 + fg) * ldb_n], bh[1], bl[1]);
      mma_tf32(acc[nt], al, bh);       // small terms first
      mma_tf32(acc[nt], ah, bl);
      mma_tf32(acc[nt], ah, bh);
    }
  }
}

__global__ void __launch_bounds__(256, 2)
gp_sample_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ ls,
                 const float* __restrict__ os_arr, const float* __restrict__ noise_arr, float jitter, int kernel_type,
                 float* __restrict__ y, float* work, int* __restrict__ info, int T, int F, int ldw) {
  extern __shared__ __align__(16) float gp_dyn[];         // the cp.async ring lives in dynamic shared memory
  float (*sR)[NB][LDR] = reinterpret_cast<float (*)[NB][LDR]>(gp_dyn);                            // sR[s][k][r] = L[r0 + r, j0 + k]
  float (*sC)[NB][LDC] = reinterpret_cast<float (*)[NB][LDC]>(gp_dyn + GP_STAGES * NB * LDR);     // sC[s][k][c] = L[c0 + c, j0 + k]
  __shared__ float sU[TR][LDU];                     // updated tile, then the solved panel rows (row-major, padded)
  __shared__ __align__(16) float sL[NB][NB];        // L11 (row-major), unit rows past the matrix edge
  __shared__ float sW[NB][LDU];                     // W = L11^-1 (row-major, lower triangular)
  __shared__ float sz[NB];
  __shared__ float s_inv_ls[GP_MAX_F];
  __shared__ int s_info;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wrp = tid >> 5;  // warp w owns rows 16 w .. 16 w + 15 of a 128-row tile, all 32 panel columns
  const int fg = lane >> 2;  // mma.sync fragment coordinates: row within 8 / column ...
  const int ft = lane & 3;   // ... and k index / column pair
  const float* xb = x + static_cast<size_t>(b) * T * F;
  const float* zb = z + static_cast<size_t>(b) * T;
  float* yb = y + static_cast<size_t>(b) * T;
  float* Lt = work + static_cast<size_t>(b) * T * ldw;   // Lt[c * ldw + r] = L[r][c]
  const float os = os_arr[b];
  const float diag_add = noise_arr[b] + jitter;

  for (int f = tid; f < F; f += blockDim.x) s_inv_ls[f] = 1.0f / ls[static_cast<size_t>(b) * F + f];
  if (tid == 0) s_info = 0;
  for (int r = tid; r < T; r += blockDim.x) yb[r] = 0.f;
  __syncthreads();

  for (int c0 = 0; c0 < T; c0 += NB) {
    const int nb = min(NB, T - c0);
    if (tid < NB) sz[tid] = (c0 + tid < T) ? zb[c0 + tid] : 0.f;
    for (int r0 = c0; r0 < T; r0 += TR) {
      // ---------------------------------------------------------------- update: acc = L[r,0:c0] L[c,0:c0]^T
      // acc[nt][j]: C fragment of the 16 x 8 tile nt -> (row 16 w + fg + 8 (j >> 1), column 8 nt + 2 ft + (j & 1))
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      // 3-stage cp.async ring over the 32-column chunks of the finished panels (rows of Lt are contiguous in r)
      const int nch = c0 / NB;
      auto issue = [&](int ch) {
        const int st = ch % GP_STAGES;
        const int j0 = ch * NB;
        // sR: 32 k-rows x 128 floats = 1024 16-byte pieces; sC: 32 x 32 floats = 256 pieces; 5 per thread
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int piece = tid + q * 256;
          const int k = piece >> 5, r4 = (piece & 31) * 4;
          const int r = r0 + r4;
          cp_async16(&sR[st][k][r4], Lt + static_cast<size_t>(j0 + k) * ldw + r, r < ldw);
        }
        {
          const int k = tid >> 3, c4 = (tid & 7) * 4;
          const int c = c0 + c4;
          cp_async16(&sC[st][k][c4], Lt + static_cast<size_t>(j0 + k) * ldw + c, c < ldw);
        }
      };
#pragma unroll
      for (int sidx = 0; sidx < GP_STAGES - 1; ++sidx) {
        if (sidx < nch) issue(sidx);
        cp_async_commit();
      }
      for (int ch = 0; ch < nch; ++ch) {
        cp_async_wait<GP_STAGES - 2>();
        __syncthreads();
        if (ch + GP_STAGES - 1 < nch) issue(ch + GP_STAGES - 1);
        cp_async_commit();
        const int st = ch % GP_STAGES;
        // A(row, k) = sR[st][k][16 w + row],  B(k, n) = sC[st][k][n]
        mma_3xtf32_16x32<NB / 8>(acc, &sR[st][0][wrp * 16], LDR, 1, &sC[st][0][0], LDC, 1, fg, ft);
      }
      cp_async_wait<0>();
      __syncthreads();     // every thread is done with the ring before the next tile refills it
      // ---------------------------------------------------------------- U = K - acc   (kernel built on the fly)
      {
        float d2[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) d2[i][j] = 0.f;
        for (int f = 0; f < F; ++f) {
          const float il = s_inv_ls[f];
          float xr[2], xc[4][2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = r0 + wrp * 16 + fg + 8 * h;
            xr[h] = r < T ? xb[static_cast<size_t>(r) * F + f] * il : 0.f;
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = c0 + nt * 8 + 2 * ft + e;
              xc[nt][e] = c < T ? xb[static_cast<size_t>(c) * F + f] * il : 0.f;
            }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float df = xr[j >> 1] - xc[nt][j & 1]; d2[nt][j] = fmaf(df, df, d2[nt][j]); }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tr_ = wrp * 16 + fg + 8 * (j >> 1), tc_ = nt * 8 + 2 * ft + (j & 1);
            const int r = r0 + tr_, c = c0 + tc_;
            float u = 0.f;
            if (r < T && c < T) {
              u = gp_kernel_value(d2[nt][j], os, kernel_type);
              if (r == c) u = os + diag_add;   // k(x,x) = 1 for every supported kernel
              u -= acc[nt][j];
            }
            sU[tr_][tc_] = u;
          }
        }
      }
      __syncthreads();
      // ---------------------------------------------------------------- diagonal block: chol + inverse in registers (warp 0)
      if (r0 == c0) {
        if (tid < 32) {
          float row[NB];
#pragma unroll
          for (int c = 0; c < NB; ++c) row[c] = sU[lane][c];
          if (lane >= nb) {
#pragma unroll
            for (int c = 0; c < NB; ++c) row[c] = (c == lane) ? 1.0f : 0.f;
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            float djj = __shfl_sync(0xffffffffu, row[j], j);
            if (!(djj > 0.f)) {
              if (lane == 0 && s_info == 0) s_info = c0 + j + 1;
              djj = 1.0f;
            }
            const float ljj = sqrtf(djj);
            const float lij = (lane > j) ? row[j] / ljj : (lane == j ? ljj : 0.f);
            row[j] = lij;
#pragma unroll
            for (int c = j + 1; c < NB; ++c) {
              const float lcj = __shfl_sync(0xffffffffu, lij, c);
              if (lane >= c) row[c] = fmaf(-lij, lcj, row[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < NB; ++c) sL[lane][c] = (c <= lane) ? row[c] : 0.f;
          // W = L11^-1 by forward substitution, lane = column j of W:  W[i][j] = (delta_ij - sum_{j<=k<i} L[i][k] W[k][j]) / L[i][i]
          // (row i of L11 is broadcast from the lane that holds it)
          float wcol[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            float acc_w = (i == lane) ? 1.0f : 0.f;
#pragma unroll
            for (int k = 0; k < i; ++k) {
              const float lik = __shfl_sync(0xffffffffu, row[k], i);
              acc_w = fmaf(-lik, wcol[k], acc_w);
            }
            const float lii = __shfl_sync(0xffffffffu, row[i], i);
            wcol[i] = (i >= lane) ? acc_w / lii : 0.f;
          }
#pragma unroll
          for (int i = 0; i < NB; ++i) sW[i][lane] = wcol[i];
        }
        __syncthreads();
      }
      // ---------------------------------------------------------------- rows below the diagonal block: V = U W^T (all warps)
      //   V[r][c] = sum_k U[r][k] W[c][k]:   A(row, k) = sU[16 w + row][k],   B(k, n) = sW[n][k]
      {
        float v[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i][j] = 0.f;
        mma_3xtf32_16x32<NB / 8>(v, &sU[wrp * 16][0], 1, LDU, &sW[0][0], 1, LDU, fg, ft);
        __syncthreads();                 // every warp has read its U rows before they are overwritten with V
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) sU[wrp * 16 + fg + 8 * (j >> 1)][nt * 8 + 2 * ft + (j & 1)] = v[nt][j];
      }
      __syncthreads();
      // ---------------------------------------------------------------- write the panel rows, y += L[r, panel] . z[panel]
      if (tid < TR) {
        const int r = r0 + tid;
        if (r < T) {
          const bool in_diag = (r0 == c0 && tid < NB);
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            const float vc = in_diag ? sL[tid][c] : sU[tid][c];
            dot = fmaf(vc, sz[c], dot);
            if (c0 + c < T) Lt[static_cast<size_t>(c0 + c) * ldw + r] = vc;     // lanes = consecutive r: coalesced
          }
          yb[r] += dot;
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) info[b] = s_info;
}

}  // namespace pfn

using namespace pfn;

extern "C" int pfn_gp_sample(const float* x, const float* z, const float* ls, const float* os, const float* noise,
                             float jitter, int kernel_type, float* y, float* work, int* info, int Bn, int T, int F,
                             void* stream) {
  PFN_CHECK_ARG(Bn > 0 && T > 0 && F > 0, "gp_sample: bad shape Bn=%d T=%d F=%d", Bn, T, F);
  PFN_CHECK_ARG(F <= GP_MAX_F, "gp_sample: F=%d exceeds %d", F, GP_MAX_F);
  PFN_CHECK_ARG(kernel_type >= PFN_KERNEL_RBF && kernel_type <= PFN_KERNEL_MATERN52, "gp_sample: bad kernel type %d", kernel_type);
  PFN_CHECK_ARG((reinterpret_cast<uintptr_t>(work) & 15) == 0, "gp_sample: work buffer must be 16-byte aligned");
  const int ldw = (T + 3) & ~3;
  constexpr int kDynSmem = GP_STAGES * (NB * LDR + NB * LDC) * static_cast<int>(sizeof(float));
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    PFN_CUDA_OK(cudaFuncSetAttribute(gp_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem));
  }
  gp_sample_kernel<<<Bn, 256, kDynSmem, reinterpret_cast<cudaStream_t>(stream)>>>(x, z, ls, os, noise, jitter, kernel_type, y,
                                                                         work, info, T, F, ldw);
  PFN_LAUNCH_OK();
  return 0;
}
