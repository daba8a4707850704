/*
 * pfn_b200.h — C ABI of libpfn_b200.so, the sm_100a kernel library behind the PFN training hot path
 * (prior sample -> masked-attention transformer fwd+bwd -> BarDistribution NLL).
 *
 * The reference (automl/TransformersCanDoBayesianInference @ 9c20031) has no FFI of its own: its hot path is
 * Python calling torch.nn / gpytorch.  Each entry point below names the reference call it replaces
 * (file:line into the reference tree, or `torch:` for the library code the reference reaches).
 *
 * Conventions
 *   - all pointers are DEVICE pointers into caller-owned buffers (PyTorch CUDA tensors); nothing is allocated
 *     or retained by the library; `stream` is a cudaStream_t passed as void*; no call synchronises.
 *   - activations are sequence-first like the reference: token row = t*B + b, row-major [T*B, cols] with an
 *     explicit leading dimension (elements).
 *   - return 0 on success; non-zero on failure with a message in pfn_last_error() (thread-local).  No C++
 *     exception crosses this boundary.
 *   - dtype codes: PFN_F32 / PFN_BF16.  bf16 kernels accumulate and keep all statistics in fp32.
 */
#ifndef PFN_B200_H_
#define PFN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFN_B200_VERSION 1

enum { PFN_F32 = 0, PFN_BF16 = 1 };
enum { PFN_EPI_NONE = 0, PFN_EPI_GELU = 1, PFN_EPI_GELU_BWD = 2, PFN_EPI_ROWDOT = 3, PFN_EPI_MUL = 4 };
enum { PFN_KERNEL_RBF = 0, PFN_KERNEL_MATERN12 = 1, PFN_KERNEL_MATERN32 = 2, PFN_KERNEL_MATERN52 = 3 };

const char* pfn_last_error(void);
int pfn_version(void);
int pfn_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM:  C[M,N] (+)= epi( sum_k A(m,k) B(n,k) + bias[n] )  (+ aux[m,n] residual)
 *   K-major operand : element (i,k) at base + i*ld + k;   MN-major operand : element (i,k) at base + k*ld + i
 *   epilogue GELU      : C = gelu_erf(acc+bias), optional C2 = acc+bias (pre-activation, saved for backward)
 *   epilogue GELU_BWD  : C = acc * gelu_erf'(aux)
 *   epilogue MUL       : C = acc * aux   (tcgen05 path only).  With GELU's c2_gelu_grad = 1 the forward stores gelu'(pre)
 *                        in C2 instead of the pre-activation, and the backward dgrad is this plain product: the ~14
 *                        instructions per element of gelu' leave the backward epilogue, whose cost is instruction issue.
 *   epilogue ROWDOT    : C = acc (+bias), and rowdot_out[m * ceil(N / rowdot_width) + n / rowdot_width] += sum over the
 *                        column group of C[m,n] * aux[m,n]   (fp32 atomics; aux is NOT added to C).  Used to produce
 *                        delta = rowsum(dO * O) per (token, head) in the out-projection dgrad (tcgen05 path only).
 *   accumulate / k_splits>1 : atomic fp32 accumulation into C (weight gradients)
 * Replaces: nn.Linear / in_proj / out_proj / linear1 / linear2 / decoder GEMMs and their autograd backward
 *   (reference transformer.py:17-18,23,84-85; torch:nn/functional.py:6478; torch:nn/modules/transformer.py:980-982).
 * pfn_gemm_bf16_tc : tcgen05 + TMA + TMEM path (bf16 operands; lda/ldb multiples of 8; 16-byte aligned bases).
 * pfn_gemm_simt    : fp32-FMA path for fp32 parity mode and shapes the tensor-core path does not take.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pfn_gemm_desc {
  int M, N, K;
  const void* A; int lda; int a_mn_major;
  const void* B; int ldb; int b_mn_major;
  void* C; int ldc; int c_dtype;
  const float* bias;
  const void* aux; int ld_aux;
  void* C2; int ldc2;
  int epilogue;
  int accumulate;
  int k_splits;
  int ab_dtype;            /* dtype of A, B, aux, C2 (simt path; the tc path is bf16 only) */
  float* rowdot_out;       /* epilogue ROWDOT: [M, ceil(N / rowdot_width)] fp32, accumulated (zero it first) */
  int rowdot_width;        /* columns per group (the head dimension); must be a multiple of 128 */
  int c2_gelu_grad;        /* epilogue GELU with C2, tcgen05 path only: 1 = C2 receives gelu'(acc+bias) instead of acc+bias */
} pfn_gemm_desc;

int pfn_gemm_bf16_tc(const pfn_gemm_desc* d, void* stream);
int pfn_gemm_simt(const pfn_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked multi-head attention under the single_eval_pos mask (mask never materialised):
 *   keys(i) = {0..sep-1}  U  ({i} if i >= sep)          (reference transformer.py:35-41 generate_D_q_matrix)
 *   o_i = sum_j softmax_j(q_i.k_j * scale) v_j           (torch:nn/functional.py:6632-6690)
 * qkv : [T*B, 3*H*dh] packed in-projection output (q | k | v), head h = columns h*dh..(h+1)*dh of each third.
 * out : [T*B, H*dh];  lse : [B*H, T] fp32 natural-log-sum-exp of the scaled scores (saved for backward).
 * Backward writes dqkv [T*B, 3*H*dh] completely (no accumulation).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pfn_attn_desc {
  int T, B, H, dh, sep;
  int dtype;
  float scale;
  const void* qkv; int ld_qkv;
  void* out; int ld_out;
  float* lse;
  const void* dout; int ld_dout;
  void* dqkv; int ld_dqkv;
  float* delta;            /* [B*H, T] fp32 scratch for backward: rowsum(dO * O) */
  int batch_major;         /* 0: token row = t*B + b (reference layout); 1: token row = b*T + t (tcgen05 kernels only) */
  /* dropout on the attention probabilities (torch:nn/functional.py multi_head_attention_forward `dropout_p`; reference
   * train.py:22 default 0.2): drop_thr = round(256 p) in [0,255], 0 = off; the keep bit of (row i, key j) of head (b,h) is
   * pfn_dropout_keep_mask's bit for (row = (b*H + h)*T + i, col = j) under drop_seed. */
  uint32_t drop_seed;
  int drop_thr;
  /* backward, tcgen05 kernels only, optional: dq_colsum[H*dh] += column sums of dQ (the q third of the in-projection bias
   * gradient), accumulated from the staged dQ tiles so that dqkv need not be re-read.  (The k third is zero in exact
   * arithmetic -- every row of dS sums to zero -- and the v third equals colsum(dO) = colsum(dz) W_out; see engine.py.) */
  float* dq_colsum;
  /* backward, tcgen05 kernels only: 1 = `delta` already holds rowsum(dO * O) in TOKEN-major layout [T*B, H] (produced by the
   * ROWDOT epilogue of the out-projection dgrad GEMM); the kernels then skip their own delta pass.  0 = `delta` is [B*H, T]
   * scratch that the backward fills itself. */
  int delta_token_major;
} pfn_attn_desc;

int pfn_attention_fwd_simt(const pfn_attn_desc* d, void* stream);
int pfn_attention_bwd_simt(const pfn_attn_desc* d, void* stream);
int pfn_attention_fwd_tc(const pfn_attn_desc* d, void* stream);
int pfn_attention_bwd_tc(const pfn_attn_desc* d, void* stream);
/* debug: clock64 event log of CTA 0 of subsequent tcgen05 attention launches ([3][cap][4] int64; which: 0 fwd, 1 dq, 2 dkv; null = off) */
int pfn_debug_attention_trace(long long* buf, int cap, int which);

/* ------------------------------------------------------------------------------------------------
 * Embedding stage (reference transformer.py:68-74):
 *   out[t,b,:] = x[t,b,:] Wx^T + bx + (t < sep ? y[t,b] wy + by : 0)
 * x [T*B, F] fp32, y [T*B] fp32, Wx [E,F], bx [E], wy [E], by [E] fp32; out [T*B, E] (out_dtype).
 * Backward accumulates (+=) into dWx, dbx, dwy, dby (fp32).
 * ---------------------------------------------------------------------------------------------- */
int pfn_embed_fwd(const float* x, const float* y, const float* Wx, const float* bx, const float* wy, const float* by,
                  void* out, int out_dtype, int T, int B, int F, int E, int sep, void* stream);
int pfn_embed_bwd(const void* dout, int dtype, const float* x, const float* y, float* dWx, float* dbx, float* dwy,
                  float* dby, int T, int B, int F, int E, int sep, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim, eps inside the sqrt, biased variance (torch:nn/modules/transformer.py:951-956
 * norm1/norm2, eps 1e-5).  The residual add is done by the producing GEMM's epilogue, so z = x + sublayer(x).
 *   fwd : h = (z - mean) * rstd * gamma + beta ; saves mean, rstd (fp32 per row)
 *   bwd : dz = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dh*gamma ; dgamma += sum dh*xhat ; dbeta += sum dh
 *         optional colsum_out[E] += sum_rows dz   (bias gradient of the GEMM that produced z)
 * ---------------------------------------------------------------------------------------------- */
int pfn_layernorm_fwd(const void* z, int ldz, const float* gamma, const float* beta, void* h, int ldh, float* mean,
                      float* rstd, int rows, int E, float eps, int dtype, void* stream);
int pfn_layernorm_bwd(const void* dh, int lddh, const void* z, int ldz, const float* mean, const float* rstd,
                      const float* gamma, void* dz, int lddz, float* dgamma, float* dbeta, float* colsum_out, int rows,
                      int E, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise dropout (+ residual) with a regenerated counter-based mask (csrc/dropout.cuh):
 *   out[r,c] = (keep(seed,r,c) ? x[r,c] * 256/(256-thr) : 0) + (residual ? residual[r,c] : 0)
 * in place when out == x.  Replaces torch:nn/modules/transformer.py:961-982 dropout1 / dropout / dropout2 of the encoder
 * layer in forward, and is applied to the incoming gradient with the same (seed, thr) in backward.  cols % 8 == 0.
 * pfn_dropout_keep_mask writes the keep bits (1/0) of a rows x cols site as bytes: the hook that lets a test's oracle
 * consume exactly the mask the kernels use (attention site: row = (b*H + h)*T + i, col = key j).
 * ---------------------------------------------------------------------------------------------- */
int pfn_dropout(const void* x, int ldx, const void* residual, int ldr, void* out, int ldo, int rows, int cols, int dtype,
                uint32_t seed, int thr, void* stream);
int pfn_dropout_keep_mask(uint8_t* out, int rows, int cols, uint32_t seed, int thr, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step of the training inner loop (reference train.py:94-97):
 *     torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm);  optimizer.step()      [torch.optim.Adam]
 * over ALL parameter tensors in two launches.  `table` (DEVICE memory, n_tensors entries) names, per tensor, the fp32
 * parameter, its gradient, the two Adam moments and -- optionally -- a bf16 copy of the parameter that is rewritten with the
 * updated value (the operand the next step's GEMMs read).  `chunk_start` (DEVICE, n_tensors + 1 ints) holds the prefix sums
 * of ceil(n / pfn_adam_chunk_elems()) per tensor; n_chunks = chunk_start[n_tensors].  `step` is the 1-based step count of
 * the bias corrections; max_grad_norm <= 0 skips the clipping; weight_decay is torch.optim.Adam's L2 term.
 * norm_sq (DEVICE, 1 float) receives the squared total gradient norm BEFORE clipping.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pfn_adam_tensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  void* p_bf16;            /* NULL = none */
  long long n;
} pfn_adam_tensor;
int pfn_adam_chunk_elems(void);
int pfn_adam_step(const pfn_adam_tensor* table, const int* chunk_start, int n_tensors, int n_chunks, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float max_grad_norm, int step, float* norm_sq, void* stream);

/* column sums: out[n] += sum_m X[m,n]   (bias gradients; torch autograd of addmm bias) */
int pfn_colsum(const void* X, int ld, int dtype, float* out, int rows, int N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bar-distribution NLL (reference bar_distribution.py:19-33 BarDistribution, :83-108 FullSupport).
 *   idx  = searchsorted_left(borders, y) - 1 with the two edge fix-ups (:19-23)             -> int64, bit-exact
 *   nll  = logsumexp(z) - z[idx] + log(width[idx])  (+ half-normal tails when full_support)
 *   rows with idx outside [0, n_bars) are counted in *oob_count (the reference asserts, :27); FullSupport clamps.
 * bwd : dlogits[r,:] = g[r] * (softmax(z[r,:]) - onehot(idx[r]))
 * ---------------------------------------------------------------------------------------------- */
int pfn_bar_nll_fwd(const void* logits, int ld, int dtype, const float* y, const float* borders, int n_bars,
                    int full_support, float* nll, int64_t* idx, float* lse, int* oob_count, int rows, void* stream);
int pfn_bar_nll_bwd(const void* logits, int ld, int dtype, const int64_t* idx, const float* lse, const float* g,
                    void* dlogits, int ld_d, int d_dtype, int n_bars, int n_cols_pad, int rows, void* stream);
int pfn_bar_bucket_idx(const float* y, const float* borders, int n_bars, int64_t* idx, int rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GP prior sample (reference priors/fast_gp.py:36-58, priors/fast_gp_mix.py:58-134):
 *   K_b = os_b * k(x_b, x_b; ls_b) + noise_b * I ;  L_b = chol(K_b) ;  y_b = L_b z_b
 * x [Bn, T, F] fp32, z [Bn, T] fp32, ls [Bn, F], os [Bn], noise [Bn] fp32, y [Bn, T] fp32,
 * work [Bn, T, ldw] fp32 scratch, ldw = T rounded up to 4; on return work[b][c][r] = L_b[r][c] (the factor, TRANSPOSED;
 * entries with r < c are unspecified), info [Bn] int (0 ok, k>0: pivot k not positive).
 * jitter is added to every diagonal (gpytorch psd_safe_cholesky retry semantics are driven by the host).
 * ---------------------------------------------------------------------------------------------- */
int pfn_gp_sample(const float* x, const float* z, const float* ls, const float* os, const float* noise, float jitter,
                  int kernel_type, float* y, float* work, int* info, int Bn, int T, int F, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PFN_B200_H_ */
