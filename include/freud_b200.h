/* freud_b200 -- C ABI of the B200-native SAE training / feature-search hot path.
 *
 * Drop-in boundary for ksadov/FREUD.  The reference has no FFI: its hot path is torch ATen ops called
 * from Python (the modules under src/models, src/utils/activations.py, src/scripts/train_sae.py:421-453).  Each entry
 * point below replaces the ATen sequence cited next to it; the Python host mirror in freud_b200/ binds
 * them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; the caller (PyTorch) owns every buffer,
 *     including workspaces; the library never frees or keeps a pointer after the call returns
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous
 *   - return 0 on success; non-zero -> freud_last_error() (thread-local) describes the failure;
 *     no exceptions cross the ABI
 *   - row-major, innermost dimension contiguous; N = B*T tokens, d = activation size, n = dictionary size
 *   - precision: FREUD_BF16 = bf16 tensor-core operands, fp32 accumulate (reference under autocast);
 *                FREUD_FP32 = fp32 results (encoder GEMM as 3-pass split-TF32 on the tensor cores)
 */
#ifndef FREUD_B200_H_
#define FREUD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FREUD_BF16 = 0, FREUD_FP32 = 1 };

const char* freud_last_error(void);
int freud_version(void);

/* ------------------------------------------------------------------ TopK SAE forward */

/* x - b_dec (topkautoencoder.py:74) written as the encoder GEMM's A operand, fused with the batch-axis
 * total variance sum((x - x.mean(0))^2) of topkautoencoder.py:104 (accumulated into *tv, which the call
 * zeroes first).  x is [B,T,d] in fp32 (x_dtype 0), fp16 (1) or bf16 (2) -- activation stores collected on CUDA hold
 * fp16; the widening is exact and happens in this pass.  FREUD_BF16: xc_hi = bf16 [N,d], xc_lo unused.
 * FREUD_FP32: xc_hi / xc_lo = fp32 [N,d] holding the tf32 high / low parts.
 * colmean (optional, [T,d] fp32) receives x.mean(0); data-parallel ranks exchange it to form the variance of
 * the concatenated batch. */
int freud_topk_prep_x(const void* x, int x_dtype, const float* b_dec, void* xc_hi, void* xc_lo, double* tv,
                      float* colmean,
                      int64_t B, int64_t T, int64_t d, int precision, void* stream);

/* Weight operand preparation (the implicit autocast cast of nn.Linear / matmul operands):
 * FREUD_BF16: hi = bf16 copy; FREUD_FP32: hi / lo = tf32 split (fp32 storage). */
int freud_split_operand(const float* w, void* hi, void* lo, int64_t numel, int precision, void* stream);

/* Fused encoder: relu((x - b_dec) @ W_enc.T + b_enc) followed by top-32 per token, without materialising
 * the [N,n] pre-activations (TopKAutoEncoder.pre_acts + select_topk, topkautoencoder.py:72-85, k == 32).
 * tcgen05/TMEM GEMM fed by TMA; selection order: value descending, index ascending (set semantics of
 * torch.topk(sorted=False)).  top_vals fp32 [N,32], top_idx int32 [N,32].  Requires n >= 64, d % 8 == 0.
 * One CTA scans all column tiles of a 128-token row block.  When N is not a multiple of 148 row blocks, the row
 * blocks of the last wave are cut into column ranges scanned by separate CTAs and merged afterwards; this needs
 * `workspace` of freud_topk_encode_workspace(N, n) bytes (may be 0).  With workspace == NULL (or too small) the
 * last wave simply runs with idle SMs; results are identical.  * hist (optional, int32 [n], zeroed by the caller): receives the number of emitted entries per feature -- the
 * histogram freud_csc_build starts from (pass it there as `offsets` with counts_ready = 1), counted with one atomic
 * per entry while the rows are written instead of in a pass of its own. */
int freud_topk_encode_workspace(int64_t N, int64_t n, int64_t* bytes);
int freud_topk_encode(const void* xc_hi, const void* xc_lo, const void* w_hi, const void* w_lo,
                      const float* b_enc, float* top_vals, int32_t* top_idx,
                      int64_t N, int64_t d, int64_t n, int precision, void* workspace, int64_t workspace_bytes, int32_t* hist,
                      void* stream);

/* Diagnostic (FREUD_ENC_STATS=1 selects an instrumented build of the encoder kernel): clock64 cycle counters summed over
 * the scanner / compactor warps of every launch since the last reset -- out[0] scanner lifetime, [1] scanner waiting for
 * an accumulator tile, [2] scanner waiting for a free candidate buffer, [3] hand-overs, [4] compactor lifetime,
 * [5] compactor waiting for a hand-over, [6] compactions, [7] compactor warps, [8] tiles scanned.  Host pointer;
 * synchronises the device.  Zeros when the instrumented build never ran. */
int freud_topk_encode_stats(unsigned long long* out, int reset);

/* out[M,N] = act(A[M,K] @ B[N,K]^T + bias[N]) on the tensor cores; act = relu if relu != 0.
 * (pre_acts materialised for the AuxK / multi-TopK branches, topkautoencoder.py:72-77,121,135; and the
 * L1 SAE's x @ W + b and c @ W.T, l1autoencoder.py:74,84.)  Operands prepared as for freud_topk_encode.  * When out has a padded pitch (ldo > N, ldo % 4 == 0) the up to three padding columns that share a 16-byte granule
 * with column N-1 may be written as zeros (the TMA store clips in 16-byte granules); columns beyond stay untouched. */
int freud_gemm_nt(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const float* bias,
                  float* out, int64_t M, int64_t N, int64_t K, int64_t ldo, int relu, int precision,
                  void* stream);

/* Exact per-row top-k of a materialised fp32 matrix (torch.topk(k, sorted=False), topkautoencoder.py:81,
 * 121,135).  If col_mask != NULL, columns with col_mask[j] == 0 are treated as -inf (the
 * torch.where(dead_mask[None], pre_acts, -inf) of :118).  Output order: value desc, index asc. */
int freud_row_topk(const float* latents, const uint8_t* col_mask, float* vals, int32_t* idx,
                   int64_t rows, int64_t n, int64_t k, void* stream);

/* AuxK support (topkautoencoder.py:118-121): the dead latents form a column subset, so their pre-activations
 * come from a GEMM against the compacted encoder rows instead of a masked [N,n] matrix.
 *   freud_gather_rows: dst[r,:] = src[rows[r],:]  (row_bytes % 4 == 0; weights, biases)
 *   freud_index_map  : out[i] = table[in[i]]      (subset-local top-k indices -> dictionary indices) */
int freud_gather_rows(const void* src, const int32_t* rows, void* dst, int64_t n_rows, int64_t row_bytes,
                      void* stream);
int freud_index_map(const int32_t* table, const int32_t* in, int32_t* out, int64_t count, void* stream);

/* Dense AuxK branch on the compacted dead-latent subset (bf16 mode).  k_aux = d/2 selected latents per token make
 * the row-sparse kernels 12x the main path's work, while a dense GEMM over the S dead latents is cheap:
 *   freud_row_topk_mask   : out[r,j] = latents[r,j] if j is in row r's top-k else 0 (bf16, row pitch ld >= n;
 *                           latents fp32 with row pitch ld_in >= n; nonneg != 0 promises
 *                           latents >= 0 (post-ReLU), which selects the streaming warp-per-row kernel)
 *   freud_scatter_add_rows: dst[rows_idx[r],:] += src[r,:] (subset gradients back into the full matrices) */
int freud_row_topk_mask(const float* latents, void* out_bf16, int64_t rows, int64_t n, int64_t k, int64_t ld_in,
                        int64_t ld, int nonneg, void* stream);
/* colsum[j] = sum_r x[r,j] over a bf16 matrix [rows, ld] (first n columns): the bias gradient of the dense AuxK
 * branch (db_enc[dead] = sum_t dpre[t,:]). */
int freud_col_sum_bf16(const void* x_bf16, float* colsum, int64_t rows, int64_t n, int64_t ld, void* stream);
int freud_scatter_add_rows(const float* src, const int32_t* rows_idx, float* dst, int64_t n_rows, int64_t row_elems,
                           void* stream);
/* Products with operands stored "the other way round" (read through MN-major tensor-core descriptors; nothing is
 * transposed in memory).  bf16 operands, fp32 accumulation / output.
 *   freud_gemm_tn_splitk: out[M,N] = A^T B, A stored [K, lda >= M], B stored [K, ldb >= N] (K = tokens: the
 *                         weight-gradient shape, e.g. dW_dec[dead] = A^T g_hat, autograd of topkautoencoder.py:17-18,
 *                         or the L1 SAE's X^T dZ, l1autoencoder.py:74,78); `splits` partials in workspace
 *                         [splits,M,N], summed by freud_sum_splits.
 *   freud_gemm_nn       : out[M,ldo] = act(A B + bias), A stored [M, lda >= K], B stored [K, ldb >= N]. */
/* out_bf16[M, ld16] = (act_bf16[M, ld16] > 0) ? scale * (A B^T) + shift : 0, A [M, lda >= K], B [N,K] bf16 K-major,
 * (scale, shift) = affine[0..1] on the device or (1, 0) if NULL: the activation gradient of the dense AuxK branch with
 * the ReLU / top-k mask applied in the GEMM epilogue (topkautoencoder.py:118-123). */
int freud_gemm_nt_mask(const void* a, const void* b, const void* act_bf16, void* out_bf16, const float* affine,
                       int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ld16, void* stream);
/* L1 SAE forward fused into the two tensor-core GEMMs (bf16 operands; l1autoencoder.py:69-95):
 *   freud_l1_encode_fused: c = relu(x W + b) stored as bf16 [M, ld16] (zero padded), sums[0] += sum(c) (= the L1
 *                          penalty's numerator, c >= 0); `latent` (fp32 [M,N]) optional.  x bf16 [M,K=d], wt = W^T
 *                          bf16 [N=n, K=d].
 *   freud_l1_decode_fused: x_hat = c W^T compared with target x [M,N=d] fp32: sums[0] += sum over x != -1 of e^2,
 *                          sums[1] += count(x != -1), sums[2] += sum e^2  (mse_loss with ignored_index=-1, :29-36, and
 *                          the return_mse value, :94); e * [x != -1] stored as bf16 [M, ld16] (the UNSCALED gradient of
 *                          the reconstruction loss w.r.t. x_hat); `x_hat` (fp32 [M,N]) optional.  c bf16 [M, lda],
 *                          w bf16 [N=d, ldb] with K = n valid columns.
 *   freud_gemm_nt_mask with affine = device (scale, shift): dz = (c > 0) ? scale * (dxhat W) + shift : 0, the
 *                          gradient through ReLU with the L1 term added (SURVEY.md M3'). */
int freud_l1_encode_fused(const void* x_bf16, const void* wt_bf16, const float* bias, void* c_bf16, float* latent,
                          double* sums, int64_t M, int64_t N, int64_t K, int64_t ld16, void* stream);
int freud_l1_decode_fused(const void* c_bf16, const void* w_bf16, const float* target, void* resid_bf16, float* x_hat,
                          double* sums, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ld16,
                          void* stream);
int freud_gemm_tn_splitk(const void* a, const void* b, float* workspace, int64_t M, int64_t N, int64_t K, int64_t lda,
                         int64_t ldb, int64_t splits, void* stream);
int freud_gemm_nn(const void* a, const void* b, const float* bias, float* out, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldo, int relu, void* stream);

int freud_sum_splits(const float* parts, float* out, int64_t splits, int64_t numel, void* stream);

/* Sparse decode + residual (eager_decode + decode, topkautoencoder.py:15-18,87-91,101):
 *   sae_out[t,:] = sum_j top_vals[t,j] * W_dec[top_idx[t,j],:] + b_dec
 * W_dec is fp32 (w_is_bf16 == 0) or a bf16 copy.  Optional outputs (NULL to skip):
 *   resid      [N,d]  = sae_out - target  (bf16 if resid_is_bf16 else fp32)
 *   sse        double += sum resid^2      (caller zeroes)
 *   colsum     [d] fp32 += sum_t resid    (caller zeroes)                                        */
int freud_topk_decode(const float* top_vals, const int32_t* top_idx, const void* W_dec, int w_is_bf16,
                      const float* b_dec, const float* target, float* sae_out, void* resid,
                      int resid_is_bf16, double* sse, float* colsum,
                      int64_t N, int64_t d, int64_t k, void* stream);

/* dacts[t,j] = sum_c g[t,c] * W_dec[top_idx[t,j], c]   (gradient of decode w.r.t. the selected
 * activations, autograd of topkautoencoder.py:17-18).  g is bf16 or fp32 [N,d]. */
int freud_topk_dacts(const void* g, int g_is_bf16, const int32_t* top_idx, const void* W_dec, int w_is_bf16,
                     float* dacts, int64_t N, int64_t d, int64_t k, void* stream);

/* Fused twin of the two calls above for the bf16 fast path (k == 32, d in {384, 512, 768, 1024, 1280}): a token's 32
 * decoder rows are fetched once (one TMA bulk copy per row into shared memory) and serve both the reconstruction
 * (topkautoencoder.py:15-18,87-91,101) and dacts[t,j] = <bf16(sae_out[t] - target[t]), W_dec[top_idx[t,j]]> (autograd
 * of :17-18).  All outputs required; sse / colsum are accumulated into (caller zeroes).  Entries with index -1
 * (another dictionary shard's) contribute nothing and get dacts 0. */
int freud_topk_decode_dacts_supported(int64_t d, int64_t k);
int freud_topk_decode_dacts(const float* top_vals, const int32_t* top_idx, const void* W_dec_bf16, const float* b_dec,
                            const void* target, int target_dtype /* 0 fp32, 1 fp16, 2 bf16 */, float* sae_out,
                            void* resid_bf16, double* sse, float* colsum, float* dacts, int64_t N, int64_t d, int64_t k,
                            void* stream);

/* fp32 mode: top_vals[t,j] <- relu((x[t] - b_dec) . W_enc[top_idx[t,j]] + b_enc[top_idx[t,j]]) with an fp32 FMA
 * chain (topkautoencoder.py:72-77 restricted to the selected latents): the tensor-core product selects, this pass
 * restores the selected values to fp32-GEMM accuracy.  Entries with index -1 are left untouched. */
int freud_topk_refine(const float* x, const float* b_dec, const float* W_enc, const float* b_enc,
                      const int32_t* top_idx, float* top_vals, int64_t N, int64_t d, int64_t k, void* stream);

/* Elementwise AuxK / multi-TopK gradient seeds (autograd of topkautoencoder.py:126-138):
 *   out = alpha * a + beta * b   (b may be NULL), alpha/beta read from device scalars coef[0], coef[1];
 * written as bf16 or fp32. */
int freud_axpby(const float* a, const float* b, const float* coef, void* out, int out_is_bf16,
                int64_t numel, void* stream);

/* ------------------------------------------------------------------ feature-sharded mode
 * Encoder / decoder rows are split across ranks; each rank runs freud_topk_encode on its shard.
 *   freud_shard_merge   : vals/idx [G,N,32] (all-gathered per-shard top-32, shard-local indices, each list sorted
 *                         value desc / index asc) -> global top-32 with dictionary indices idx + g*n_local
 *   freud_shard_localize: keep the winners in [lo, lo+n_local): local index or -1, value or 0.  Entries with
 *                         index -1 are skipped by freud_topk_decode / freud_topk_dacts / freud_csc_build.
 *   freud_residual      : after the partial reconstructions are all-reduced: resid = sae_out - target, SSE and
 *                         column sums (the tail of freud_topk_decode). */
int freud_shard_merge(const float* vals, const int32_t* idx, float* out_vals, int32_t* out_idx, int64_t N,
                      int64_t G, int64_t n_local, void* stream);
int freud_shard_localize(const float* vals, const int32_t* gidx, float* lvals, int32_t* lidx, int64_t count,
                         int64_t lo, int64_t n_local, void* stream);
int freud_residual(const float* sae_out, const float* target, void* resid, int resid_is_bf16, double* sse,
                   float* colsum, int64_t N, int64_t d, void* stream);

/* ------------------------------------------------------------------ TopK SAE backward */

/* Feature-major (CSC) index of the selected entries: offsets[f]..offsets[f+1] lists the flat positions
 * p = t*k + j with top_idx[p] == f.  offsets is int32 [n+1]; entries int32 [N*k]; cursor int32 [n+1]
 * workspace (fill cursors, then the queue of long lists for the sort pass and its counter).  Every list of up to 4096
 * entries is token-ordered, so the gradient sums are run-to-run deterministic.  Also the did_fire bookkeeping of train_sae.py:442 (offsets[f+1] > offsets[f]). */
int freud_csc_build(const int32_t* top_idx, int64_t N, int64_t k, int64_t n, int32_t* offsets,
                    int32_t* entries, int32_t* cursor, int counts_ready, void* stream);

/* Row-sparse weight gradients of one decode (autograd of :17-18 and of nn.Linear, :75):
 *   dW_dec[f,:] (+)= s_dec * sum_{p in list(f)} top_vals[p] * g[t(p),:]
 *   dpre[p]       = top_vals[p] > 0 ? s_enc * dacts[p] : 0
 *   dW_enc[f,:] (+)= sum_p dpre[p] * xc[t(p),:]        db_enc[f] (+)= sum_p dpre[p]
 * freud_csc_meta first lays the per-entry terms out in list order: meta int32 [3 * n_entries] = token t(p),
 * top_vals[p] * s_dec and dpre[p] planes (s_dec = scales[0], s_enc = scales[1], device floats; n_entries = N*k).
 * freud_topk_sparse_grads then walks the lists: up to 192 entries by one warp per 32-slice column slab, longer
 * ones in chunks of 1024 entries (one CTA each; chunk_off is an int32 [n+1] workspace).  g / xc: bf16 [N,d], or
 * fp32 with xc = x - b_dec recomputed from x and b_dec when xc_is_bf16 == 0.  Every row of dW_dec / dW_enc /
 * db_enc is written (accumulate == 0) or added to (accumulate != 0); no memset is needed beforehand.
 * A contiguous feature range [f0, f1) can be processed alone by passing offsets + f0, the row / element pointers
 * of f0 and n = f1 - f0 (the lists keep their absolute positions). */
int freud_csc_meta(const int32_t* offsets, const int32_t* entries, const float* top_vals, const float* dacts,
                   const float* scales, int32_t* meta, int64_t n_entries, int64_t n, int64_t k, void* stream);
int freud_topk_sparse_grads(const int32_t* offsets, const int32_t* meta, const void* g, int g_is_bf16,
                            const void* xc, int xc_is_bf16, const float* b_dec, float* dW_dec, float* dW_enc,
                            float* db_enc, int32_t* chunk_off, int64_t n_entries, int64_t n, int64_t d, int64_t k,
                            int accumulate, void* stream);

/* db_dec[c] (+)= s * colsum[c] - sum_f db_enc_part[f] * W_enc[f,c]   (s = scales[0]; either term may be
 * skipped with a NULL pointer).  Autograd of `x - b_dec` (:74) and `+ b_dec` (:91).  W_enc fp32, or its bf16 copy
 * (bf16 mode: under autocast the reference back-propagates through the bf16 matmul operand). */
int freud_topk_bdec_grad(const float* colsum, const float* scales, const float* db_enc, const void* W_enc,
                         int w_is_bf16, float* db_dec, int64_t n, int64_t d, int accumulate, void* stream);

/* Loss scalars (topkautoencoder.py:104-106,126-132,138,150), all on device:
 *   tv' = tv == 0 ? 1 : tv;  out[0] = fvu = sse/tv';  out[1] = mse = sse/numel;
 *   out[2] = out[3] = 2/tv' (the (s_dec, s_enc) pair freud_topk_sparse_grads reads);  out[4] = tv' */
int freud_topk_loss_scalars(const double* sse, const double* tv, float* out, int64_t numel, void* stream);

/* Dead-latent bookkeeping (train_sae.py:442-446): frames[f] = fired(f) ? 0 : frames[f] + n_tokens,
 * fired(f) = offsets[f+1] > offsets[f]. */
int freud_dead_latent_update(const int32_t* offsets, int64_t* frames, int64_t n, int64_t n_tokens, void* stream);

/* W_dec /= ||W_dec[i,:]|| + eps   (set_decoder_norm_to_unit_norm, topkautoencoder.py:153-159) */
int freud_rownorm_project(float* W, int64_t rows, int64_t cols, float eps, void* stream);
/* G -= (G . W)_row * W           (remove_gradient_parallel_to_decoder_directions, :161-175) */
int freud_remove_parallel_grad(float* G, const float* W, int64_t rows, int64_t cols, void* stream);

/* ------------------------------------------------------------------ L1 SAE */

/* W[:,j] /= max(||W[:,j]||, 1e-12) in place (F.normalize(dim=0), l1autoencoder.py:71-73); also writes the
 * transposed copy Wt [n,d] used as the K-major GEMM operand.  W is decoder.weight [d,n]. */
int freud_l1_colnorm(float* W, float* Wt, int64_t d, int64_t n, void* stream);

/* L1 SAE loss values and gradient scales from the accumulated sums acc = [sum|c|, masked sse, unmasked count, sse]
 * (l1autoencoder.py:85-86,29-36): out[0..4] = l1_loss, reconstruction_loss (= recon_alpha * masked mse), mse,
 * 2 * recon_alpha / count, 1 / n_glob.  n_glob = tokens of the (global) batch, d = activation size. */
int freud_l1_loss_scalars(const double* acc, double n_glob, double d, double recon_alpha, float* out, void* stream);

/* Loss pieces of L1AutoEncoder.forward (l1autoencoder.py:85-86,94, mse_loss :29-36):
 *   acc[0] += sum |c|,  acc[1] += sum_{x != -1} (x_hat-x)^2,  acc[2] += #{x != -1},  acc[3] += sum (x_hat-x)^2
 * and, if dxhat != NULL, dxhat = (x != -1) * (x_hat - x)  (scaled later).  acc: 4 doubles, caller zeroes. */
int freud_l1_loss_reduce(const float* latent, const float* x_hat, const float* x, float* dxhat, double* acc,
                         int64_t N, int64_t d, int64_t n, void* stream);

/* dz = (c > 0) * (s_recon * dc_recon + s_l1) in place on dc_recon [N,n]; db[j] += sum_t dz (caller zeroes). */
int freud_l1_dz(float* dc, const float* latent, const float* scales, float* db, int64_t N, int64_t n,
                void* stream);

/* dW[d,n] = scales[0] * X^T dz + scales[1] * dxhat^T c: the two tied-weight accumulating GEMMs of SURVEY.md M3'
 * (autograd of l1autoencoder.py:74,84) as one split-K kernel over the token axis. */
int freud_l1_weight_grad(const float* x, const float* dz, const float* dxhat, const float* latent,
                         const float* scales, float* dW, int64_t N, int64_t d, int64_t n, void* stream);

/* ------------------------------------------------------------------ optimiser (train_sae.py:449-450) */

#define FREUD_MAX_TENSORS 8
typedef struct {
  int32_t count;
  float* param[FREUD_MAX_TENSORS];
  float* grad[FREUD_MAX_TENSORS];
  float* exp_avg[FREUD_MAX_TENSORS];
  float* exp_avg_sq[FREUD_MAX_TENSORS];
  void* bf16_shadow[FREUD_MAX_TENSORS]; /* optional bf16 copy refreshed with the new parameter, or NULL */
  int64_t numel[FREUD_MAX_TENSORS];
} freud_tensor_list;

/* sumsq[0] = sum over all tensors of grad^2 (zeroed first): the inner norms of clip_grad_norm_
 * (torch/nn/utils/clip_grad.py:50). */
int freud_grad_sumsq(const freud_tensor_list* host_list, double* sumsq, void* stream);
/* grad *= clamp(max_norm / (sqrt(sumsq) + 1e-6), max=1)  (clip_grad.py:121,165-169); total norm -> *norm_out */
int freud_clip_grads(const freud_tensor_list* host_list, const double* sumsq, float max_norm, float* norm_out,
                     void* stream);
/* Adam (torch/optim/adam.py:347; betas (0.9,0.999) defaults passed explicitly).  `step` is the 1-based
 * step count.  If sumsq != NULL the clip coefficient is applied to the gradient on the fly (fused
 * clip + Adam; the stored gradients are left unscaled). */
int freud_adam_step(const freud_tensor_list* host_list, double lr, double beta1, double beta2, double eps,
                    int64_t step, const double* sumsq, float max_norm, void* stream);
/* RAdam with non-decoupled weight decay (torch/optim/radam.py:256-361). */
int freud_radam_step(const freud_tensor_list* host_list, double lr, double beta1, double beta2, double eps,
                     double weight_decay, int64_t step, const double* sumsq, float max_norm, void* stream);

/* ------------------------------------------------------------------ fused data-parallel optimiser step
 * (N ranks == the single-GPU step of train_sae.py:448-450 on the concatenated batch.)
 * The flat parameter space [W_enc | b_enc | W_dec | b_dec] (each padded to 32 elements) is cut into one contiguous
 * slice [lo, hi) per rank.  peer_* are HOST arrays of `world` device pointers, one per rank, to the ranks' flat
 * buffers in NVLink peer-mapped (symmetric) memory; mc_* the NVSwitch multicast address of the same buffer or NULL.
 *
 * freud_dp_reduce_scatter: this rank's gradient buffer [lo, hi) <- sum over ranks of peer_grads[r][lo, hi) (fixed
 *   rank order, or multimem.ld_reduce when mc_grad != NULL); the slice's sum of squares is accumulated in *partial
 *   (device double, zero on entry, re-zeroed on exit together with *done_ctr) and posted to slot `rank` of every
 *   peer's slot buffer peer_slots[r] (>= world doubles).  A cross-rank barrier must separate it from the ranks'
 *   backward kernels before and from freud_dp_adam_allgather after.
 * freud_dp_adam_allgather: clip coefficient from the `world` posted partials (if clip), Adam (torch/optim/adam.py:347)
 *   on [lo, hi) with exp_avg / exp_avg_sq holding ONLY the slice (index e - lo), fp32 master updated in
 *   peer_params[rank]; the updated values of a region go to every rank: is_weight regions as bf16 into
 *   peer_shadows[r] (what the tensor-core and gather kernels read), other regions as fp32 into peer_params[r].
 *   A cross-rank barrier must follow before any rank reads them. */
int freud_dp_reduce_scatter(void* const* peer_grads, const float* mc_grad, int64_t world, int64_t rank, int64_t lo,
                            int64_t hi, double* partial, unsigned int* done_ctr, void* const* peer_slots,
                            void* stream);
int freud_dp_adam_allgather(void* const* peer_params, void* const* peer_shadows, float* mc_param, void* mc_shadow,
                            const float* grad, float* exp_avg, float* exp_avg_sq, int64_t world, int64_t rank,
                            int64_t lo, int64_t hi, const int64_t* region_begin, const int64_t* region_end,
                            const int32_t* region_is_weight, int64_t n_regions, double lr, double beta1, double beta2,
                            double eps, int64_t step, const double* slots, float max_norm, int clip, void* stream);

/* ------------------------------------------------------------------ validation feature statistics
 * (src/scripts/train_sae.py:70-118,175-178)
 * freud_feature_absmax: out[f] = max |acts[i]| over entries with idx[i] == f, 0 if f never appears
 *                       (topk_feature_extraction on one file's [T,k] encoding; idx int64 or int32).
 * freud_col_absmax    : out[j] = max_r |x[r,j]|   (the L1 SAE's per-file latent abs-max). */
int freud_feature_absmax(const float* acts, const void* idx, int idx_is_int64, int64_t count, float* out, int64_t n,
                         void* stream);
int freud_col_absmax(const float* x, int64_t rows, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------ feature search (utils/activations.py) */

/* Per-file statistics of one feature over dense activations acts[N_files, T, F] (fp32 or fp16):
 * for the first n_frames[i] frames (trim_activation, utils/activations.py:19-29) of column `feature`:
 *   vmax[i] = max, amax[i] = first argmax, vabs[i] = signed value at the first argmax of |a|  (:104-119).
 * n_frames[i] == 0 -> vmax = -inf, amax = -1.  If trace != NULL it receives the column [N_files, T] fp32. */
int freud_search_dense(const void* acts, int acts_is_fp16, const int32_t* n_frames, int64_t n_files, int64_t T,
                       int64_t F, int64_t feature, float* vmax, int32_t* amax, float* vabs, float* trace,
                       void* stream);
/* Same statistics for indexed (top-k) activations vals/idx [N_files, T, k] (activation_tensor_from_indexed,
 * utils/activations.py:41-57: value of the first slot whose index == feature, else 0).  idx is int64 or int32. */
int freud_search_indexed(const float* vals, const void* idx, int idx_is_int64, const int32_t* n_frames,
                         int64_t n_files, int64_t T, int64_t k, int64_t feature, float* vmax, int32_t* amax,
                         float* vabs, float* trace, void* stream);
/* The same three statistics for EVERY feature column in one pass over the dense store (utils/activations.py:94-130
 * loops over files per query; here one scan serves all later queries): tables [N_files, F], column `f` equal to what
 * freud_search_dense returns for feature f.  F % 4 == 0. */
int freud_search_table_dense(const void* acts, int acts_is_fp16, const int32_t* n_frames, int64_t n_files, int64_t T,
                             int64_t F, float* vmax_tab, int32_t* amax_tab, float* vabs_tab, void* stream);

/* Ranking of utils/activations.py:88-93,121-130: among files passing the min/max filter (applied to the
 * signed statistic), the n_top largest keys (vmax, or |vabs| when absolute != 0), ties broken by lower file
 * index (stable sort).  use_min/use_max select the optional bounds.  out_files int32 [n_top] (-1 padded),
 * out_count int32 [1]. */
int freud_search_topn(const float* vmax, const float* vabs, int64_t n_files, int absolute, int use_min,
                      double min_val, int use_max, double max_val, int64_t n_top, int32_t* out_files,
                      int32_t* out_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FREUD_B200_H_ */
