/* clipk — C ABI of the B200-native (sm_100a) patch-aligned contrastive scoring + loss kernels.
 *
 * The reference (lst627/CLIP-Embeds) has no FFI / plugin registry: its boundary for this path is Python
 * (`nn.Module.forward` + autograd).  This header is the boundary a maintainer binds instead of the ATen op
 * sequences cited on each entry point (paths relative to the reference root;
 * PACL = Patch-Aligned-Contrastive-Learning).  Python binding: clip_embeds_b200/_lib.py (ctypes), see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are row-major and contiguous
 *     unless a leading dimension is given; sizes are element counts;
 *   - the library never allocates, frees or synchronises: the caller owns inputs, outputs, saved statistics
 *     and workspaces (`*_workspace_bytes` queries); all work is enqueued on `stream` (a cudaStream_t);
 *   - return 0 on success, a negative CLIPK_ERR_* otherwise; clipk_last_error() gives a thread-local message;
 *   - requires an sm_100 device: any other device returns CLIPK_ERR_ARCH (there is no fallback path);
 *   - re-entrant, no global mutable state besides per-device attribute caches.
 */
#ifndef CLIPK_H_
#define CLIPK_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPK_OK 0
#define CLIPK_ERR_INVALID (-1)
#define CLIPK_ERR_CUDA (-2)
#define CLIPK_ERR_ARCH (-3)

#define CLIPK_BF16 0
#define CLIPK_F32 1
#define CLIPK_F16 2   /* IEEE half: the paired / eval-scorer path (the reference evaluates under fp16 autocast, eval_pacl.py:53) */

#define CLIPK_ACT_SIGMOID10 0 /* sigmoid(10*cos): reference parity, pacl.py:133 */
#define CLIPK_ACT_ONES 1      /* activations overwritten with ones: the checked-in forward(), pacl.py:141-142 */
/* softmax over the patches of 10 * cos (north_star (2); SURVEY Appendix A.1 lists it as the third activation, the
 * reference itself has no such branch).  The pooled feature is L2-normalised afterwards, so the softmax denominator
 * cancels: the kernels pool with a_p = exp(10 (s_p - 1)) (the -1 only keeps the weights in (0, 1]), whose Jacobian
 * is the diagonal 10 a_p; act_out of the paired forward holds these UN-normalised weights. */
#define CLIPK_ACT_SOFTMAX10 2

const char* clipk_last_error(void);
int clipk_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py reports the per-step delta) */
unsigned long long clipk_launch_count(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x), else CLIPK_ERR_ARCH. */
int clipk_check_device(void);
/* Diagnostic: with CLIPK_TRACE=1 in the environment every engine launch is bracketed by CUDA events; this call
 * synchronises the device, prints the per-kernel totals to stdout and clears the trace.  No-op otherwise. */
int clipk_trace_dump(void);

/* ---------------------------------------------------------------------------------------------------------
 * Engine building block (also the unit-test entry of the tcgen05 pipeline):
 *   C[b] (+)= alpha * A[b] * B[b]^T,   A: M x K, B: N x K, bf16 operands, fp32 accumulation in TMEM.
 *   a_mn / b_mn = 0: operand stored [rows][K] (K contiguous, leading dim ld = row stride)
 *               = 1: operand stored [K][rows] (rows contiguous, ld = stride between k)
 *   out_dtype CLIPK_F32 | CLIPK_BF16; accumulate != 0 adds into C (fp32 only).
 * Replaces: torch `@` / `einsum` / `bmm` on the path (cuBLAS), e.g. pacl.py:129, :460, :472, :499-500.
 */
int clipk_gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn, int64_t ldb,
                    int64_t strideB, void* C, int64_t ldc, int64_t strideC, int out_dtype, int M, int N, int K,
                    int batches, float alpha, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * PACL paired path (reference training / eval semantics; HBM-bound, one pass over V per direction).
 *   V [B / v_div, P, D], T [B, D] (dtype CLIPK_BF16 | CLIPK_F32 | CLIPK_F16; arithmetic is fp32 in every case).  Sample b uses image b / v_div
 *   (v_div = 1: paired batch; v_div = K: one image scored against K captions, eval_pacl.py:50-57).
 *   act_out [B,P] (nullable) = sigmoid(10 cos)            -> patch_alignment, pacl.py:120-133
 *   img_feat [B,D] = n(sum_p a_p V_p), txt_feat [B,D] = n(T) -> forward, pacl.py:140-145 (a := 1 for CLIPK_ACT_ONES)
 *   cosine [B] (nullable) = <img_feat, txt_feat>           -> diagonal of eval_pacl.py:56
 *   stats [B,2] = (||u||, ||t||), saved for backward.
 * Backward: d_img, d_txt [B,D] fp32 (d_txt nullable) -> dV [B,P,D], dT [B,D] in the input dtype.
 */
int clipk_pacl_paired_fwd(const void* V, const void* T, int dtype, int B, int v_div, int P, int D, int act,
                          float* act_out, float* img_feat, float* txt_feat, float* cosine, float* stats,
                          void* stream);
int clipk_pacl_paired_bwd(const void* V, const void* T, int dtype, int B, int P, int D, int act,
                          const float* img_feat, const float* stats, const float* d_img, const float* d_txt,
                          void* dV, void* dT, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * PACL all-pairs text-conditioned scores (north_star (1)-(3); reference: the eval call model(img_i, texts) of
 * eval_pacl.py:53-57, :303-309 looped over images; patch_alignment pacl.py:120-133; pooling pacl.py:143-145).
 *   V bf16 [Bi,P,D], T bf16 [Bt,D]  ->  scores fp32 [Bi,Bt] = c * < n(sum_p sigmoid(10 s_ikp) V_ip), n(t_k) >
 * Saved for backward (caller-owned): rnV [Bi,P], rnT [Bt], num [Bi,Bt], usq [Bi,Bt].  The [Bi,Bt,P] activations
 * only ever exist for `group` images per lane inside `workspace` (tcgen05 GEMMs with fused epilogues); image groups
 * are issued round-robin on `lanes` (1..4) internal streams forked from / joined to `stream` with events.
 * `pooled` (nullable, bf16 [Bi,Bt,D], caller-owned): when given, the forward also stores the un-normalised pooled
 * vectors u_ik and the backward reads them instead of recomputing activations + pooling (7 GEMM units per step
 * instead of 8; SURVEY 7.3-5 "store U as bf16").  Pass the same buffer (or NULL both times) to fwd and bwd.
 * Backward: dscores [Bi,Bt] -> dV bf16 [Bi,P,D], dT fp32 [Bt,D] (gradient w.r.t. the raw T).
 * workspace_bytes `mode`: bit 0 = backward, bit 1 = pooled.
 * rnv_given != 0: `rnV` already holds 1 / max(|V_ip|, 1e-12) (the projection head's output GEMM emits the squared row
 * norms, clipk_patch_proj_fwd `ysq`): the forward then skips its own pass over V.
 */
size_t clipk_pacl_allpairs_workspace_bytes(int Bi, int Bt, int P, int D, int group, int lanes, int mode);
int clipk_pacl_allpairs_fwd(const void* V, const void* T, int Bi, int Bt, int P, int D, int act, float c, float* rnV,
                            float* rnT, float* num, float* usq, float* scores, void* pooled, void* workspace,
                            size_t ws_bytes, int group, int lanes, int rnv_given, void* stream);
int clipk_pacl_allpairs_bwd(const void* V, const void* T, int Bi, int Bt, int P, int D, int act, float c,
                            const float* rnV, const float* rnT, const float* num, const float* usq,
                            const float* dscores, const void* pooled, void* dV, float* dT, void* workspace,
                            size_t ws_bytes, int group, int lanes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Cross-entropy over a materialised fp32 matrix L [M,N] (leading dim ld): F.cross_entropy at pacl.py:509-512,
 * loss.py:188-191.  label of row i = labels[i] (nullable) or i + label_offset; labels < 0 are ignored rows.
 *   clipk_ce_rows       : row_lse [M], row_loss [M] (= lse - L[i,label], 0 for ignored rows; nullable)
 *   clipk_ce_cols       : per-column (max, sum exp(L - max)) over the M local rows (merge across ranks, then log)
 *   clipk_ce_scores_grad: dL = w_row (softmax_row - onehot) + w_col (softmax_col - onehot), label = i + offset
 *   clipk_ce_rows_grad  : dL = row_w[i] (softmax_row - onehot) with explicit labels
 */
int clipk_ce_rows(const float* L, int M, int N, int64_t ld, const int64_t* labels, int64_t label_offset,
                  float* row_lse, float* row_loss, void* stream);
int clipk_ce_cols(const float* L, int M, int N, int64_t ld, float* col_max, float* col_sum, void* stream);
/* Sharded InfoNCE on a score matrix with ONE collective: every rank all-gathers a payload of 2N + 2 floats
 * (col_max [N], col_sum [N] from clipk_ce_cols, then clipk_ce_rowsums' two sums); clipk_ce_merge turns the W gathered
 * payloads [W][2N+2] into the merged column LSEs and the GLOBAL loss 1/2 (CE(L, I) + CE(L^T, I)) (pacl.py:509-512).
 *   clipk_ce_rowsums: out2[0] = sum_i row_loss_i, out2[1] = sum_i (row_lse_i - row_loss_i)  (the label logits)  */
int clipk_ce_rowsums(const float* row_lse, const float* row_loss, int M, float* out2, void* stream);
int clipk_ce_merge(const float* gathered, int W, int N, float* col_lse, float* loss, void* stream);
int clipk_ce_scores_grad(const float* L, int M, int N, int64_t ld, const float* row_lse, const float* col_lse,
                         int64_t label_offset, float w_row, float w_col, float* dL, void* stream);
int clipk_ce_rows_grad(const float* L, int M, int N, int64_t ld, const float* row_lse, const int64_t* labels,
                       int64_t label_offset, const float* row_w, float* dL, void* stream);

/* fp32 GEMM with generic strides (fp32 feature path: `logit_scale * x @ y.T`, pacl.py:499-500, loss.py:156-164):
 *   C[m,n] = alpha * sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] + beta * C[m*ldc + n] */
int clipk_sgemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                    int64_t ldc, int M, int N, int K, float alpha, float beta, void* stream);

/* The same product with fp32 accuracy on the bf16 tensor cores: every operand value is split into three bf16 terms
 * (h + m + l = 24 mantissa bits) and the six significant cross products are folded into one tcgen05 GEMM over
 * K' = 6 * roundup(K, 64) (fp32 accumulation in TMEM).  Same stride convention as clipk_sgemm_f32; `accumulate`
 * != 0 adds to C (beta = 1).  `workspace` (128-byte aligned) holds the split operands. */
size_t clipk_gemm_f32_split_workspace_bytes(int M, int N, int K);
int clipk_gemm_f32_split(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                         int64_t ldc, int M, int N, int K, float alpha, int accumulate, void* workspace,
                         size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * bf16 feature cross-entropy on the tcgen05 engine (NegCLIP ClipLoss, loss.py:137-193; PACL ClipLoss, pacl.py:489-514):
 *   logits = scale * X Y^T + bias, X bf16 [M,D], Y bf16 [N,D]; the [M,N] logits are never written.
 *   fwd: row_lse [M], row_loss [M].   bwd: row_w [M] (upstream / #valid rows; 0 for ignored rows)
 *        dX [M,D] / dY [N,D] fp32 (nullable; acc* != 0 accumulates into the buffer).
 */
size_t clipk_ce_feat_workspace_bytes(int M, int N);
/* backward workspace including the split-K slabs of the dX GEMM (a workspace sized by the query above still works:
 * the dX GEMM then runs unsplit) */
size_t clipk_ce_feat_bwd_workspace_bytes(int M, int N, int D);
/* `scale_dev` (nullable): when given, the logit scale is read from this device scalar instead of `scale` -- open_clip
 * passes logit_scale.exp() as a CUDA tensor (open_clip_train/train.py:107), and copying it to the host would stall the
 * stream once per call.
 * `slab_counts` (nullable, int32 [W] on the device) with `slab_n0`, `slab_rows`: the columns of Y from slab_n0 on are W
 * fixed-capacity slabs of slab_rows rows (each rank's hard-negative captions, all-gathered at capacity); only the first
 * slab_counts[r] rows of slab r exist, the rest are masked out of the softmax (the reference's host-side size exchange,
 * loss.py:78-86, becomes an in-band device array).  slab_n0 and slab_rows must be multiples of 32. */
int clipk_ce_feat_fwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias, const float* scale_dev,
                      const int* slab_counts, int slab_n0, int slab_rows, const int64_t* labels, int64_t label_offset, float* row_lse, float* row_loss, void* workspace,
                      size_t ws_bytes, void* stream);
int clipk_ce_feat_bwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias, const float* scale_dev,
                      const int* slab_counts, int slab_n0, int slab_rows, const int64_t* labels, int64_t label_offset, const float* row_lse, const float* row_w, float* dX,
                      int accX, float* dY, int accY, void* workspace, size_t ws_bytes, void* stream);

/* Symmetric cross-entropy from ONE logits GEMM (the reference computes logits_per_image and
 * logits_per_text with two GEMMs, open_clip/src/open_clip/loss.py:156-157, pacl.py:499-500):
 *   rows    : X [M,D] (this rank's images), label of row m = m + label_offset            -> row_lse, row_loss [M]
 *   columns : the first `ncol` rows of Y (the original captions of all ranks) also carry a CE over ALL image rows;
 *             col_sum [ncol] += sum_m exp(logit[m,j] - (scale / 2 + bias))   (zeroed here; partial over THIS rank's rows:
 *             all-reduce it over the ranks, then col_lse[j] = scale / 2 + bias + log(col_sum[j]))
 * Backward: dL = row_w[m] (softmax_row - onehot) + (exp(logit + col_bias[j]) - col_w[lab] onehot), with
 *   col_bias[j] = log(col_w[j]) - col_lse[j]; dX [M,D] and dY [N,D] (this rank's partial: sum it over the ranks) in fp32
 *   (grads_bf16 = 0) or bf16 (grads_bf16 = 1, M <= 4096).  bf16 features, CTA-pair tcgen05 engine, logits never written;
 *   workspace as clipk_ce_feat_bwd_workspace_bytes(M, N, D).  scale_dev / slab_* as above.
 *   phase (M <= 4096): 0 = everything; 1 = dL + dY only, 2 = dX only from the dL that phase 1 left in the workspace -- the
 *   caller starts the reduce-scatter of dY between the two calls so that it overlaps the dX GEMM. */
int clipk_ce_sym_fwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias, const float* scale_dev,
                     const int* slab_counts, int slab_n0, int slab_rows, int64_t label_offset, int ncol, float* row_lse,
                     float* row_loss, float* col_sum, void* workspace, size_t ws_bytes, void* stream);
int clipk_ce_sym_bwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias, const float* scale_dev,
                     const int* slab_counts, int slab_n0, int slab_rows, int64_t label_offset, int ncol,
                     const float* row_lse, const float* row_w, const float* col_bias, const float* col_w, void* dX,
                     void* dY, int grads_bf16, int phase, void* workspace, size_t ws_bytes, void* stream);

/* The same backward with bf16 gradients written straight from the GEMM epilogues (dX, dY bf16, both required, no
 * accumulation); supported for 128 < M <= 4096 and N > 128 (CLIPK_ERR_INVALID otherwise: use the fp32 entry). */
int clipk_ce_feat_bwd_bf16(const void* X, const void* Y, int M, int N, int D, float scale, float bias,
                           const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows,
                           const int64_t* labels, int64_t label_offset, const float* row_lse, const float* row_w,
                           void* dX, void* dY, void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * SPARC token-to-patch alignment (sparc.forward, PACL/model/pacl.py:453-478):
 *   S = L V^T (raw), min-max over patches, threshold sigma, row-normalise, G = W V; outputs l_hat = n(L),
 *   g_hat = n(G) (fp32 [B,T,D]) and their norms [B,T] (saved for backward).  V bf16 [B,P,D], L bf16 [B,T,D], T <= 128.
 *   pooled (nullable, fp32 [B,D]) = mean over patches of V (SparcLoss's global image feature, pacl.py:561): produced by the
 *   same pass (fused kernel: an all-ones row of the pooling GEMM).
 * Backward recomputes S / W inside `workspace`; d_g_hat, d_l_hat fp32 [B,T,D]; g_add (nullable) fp32 [B,D] is added
 * to every patch row of dV (the broadcast gradient of mean_p V from SparcLoss's global term, pacl.py:561);
 * dV [B,P,D] bf16 (dv_bf16 != 0) or fp32; dL fp32 [B,T,D].
 */
size_t clipk_sparc_workspace_bytes(int B, int T, int P, int D, int backward);
int clipk_sparc_align_fwd(const void* V, const void* L, int B, int T, int P, int D, float sigma, float* l_hat,
                          float* g_hat, float* lnorm, float* gnorm, float* pooled, void* workspace, size_t ws_bytes,
                          void* stream);
int clipk_sparc_align_bwd(const void* V, const void* L, int B, int T, int P, int D, float sigma, const float* l_hat,
                          const float* g_hat, const float* lnorm, const float* gnorm, const float* d_g_hat,
                          const float* d_l_hat, const float* g_add, void* dV, int dv_bf16, float* dL, void* workspace,
                          size_t ws_bytes, void* stream);

/* SparcLoss local term, both directions (masked_pairwise_contrastive_loss, pacl.py:522-556; a = g_hat, b = l_hat
 * fp32 [B,T,D], mask fp32 [B,T]).  fwd: loss_sum [B]; total local loss = sum_b loss_sum / (2 * sum(mask)).
 * bwd: wgt = DEVICE scalar (upstream * 0.5 / sum(mask)); d_a, d_b fp32 [B,T,D]; loss_sum_scratch [B]. */
size_t clipk_sparc_local_workspace_bytes(int B, int T, int D);
int clipk_sparc_local_fwd(const float* a, const float* b, const float* mask, int B, int T, int D, float scale,
                          float* loss_sum, void* workspace, size_t ws_bytes, void* stream);
int clipk_sparc_local_bwd(const float* a, const float* b, const float* mask, int B, int T, int D, float scale,
                          const float* wgt, float* d_a, float* d_b, float* loss_sum_scratch, void* workspace,
                          size_t ws_bytes, void* stream);

/* torch.mean(X, dim=1) for X [B,R,D] (bf16 | fp32) -> fp32 [B,D] (pacl.py:443-447, :561-562) and F.normalize rows. */
int clipk_mean_dim1(const void* X, int dtype, int B, int R, int D, float* out, void* stream);
int clipk_normalize_rows_fwd(const float* X, int64_t rows, int D, float* out, float* norm, void* stream);
int clipk_normalize_rows_bwd(const float* xh, const float* g, const float* norm, int64_t rows, int D, float* dx,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Projection heads (the producer of the patch tensor the scorer reads):
 *   visual_projection = LayerNorm -> Dropout -> Patch_Projection   PACL/model/pacl.py:35-48, :70-74
 *   text_projection   = LayerNorm -> Dropout -> Linear             PACL/model/pacl.py:75-79
 * Rows are tokens.  x: bf16 | fp32 [rows, D] (dtype enum); xn bf16; gamma, beta, mean, rstd fp32.
 *   clipk_ln_fwd : xn = (x - mean) rstd gamma + beta, saves mean / rstd (biased variance, eps inside the sqrt)
 *   clipk_ln_bwd : dgamma, dbeta [D] from dxn (bf16); dx (nullable; dtype of x) = gradient wrt the LayerNorm input
 *   Dropout (pacl.py:72,77) is fused into both: drop_p > 0 zeroes each element of xn with probability drop_p (rounded
 *   to a multiple of 2^-16) and scales the others by 1 / (1 - drop_p); the decision of element i depends only on
 *   (seed, i).  keep_bits [rows * D / 8] bytes receives one keep bit per element (bit j of byte v = element 8 v + j) and
 *   is handed back to clipk_ln_bwd, which masks / scales dxn before the LayerNorm Jacobian.  drop_p = 0: identity.
 * Patch_Projection: y = W1 xn + b1 + W3 gelu(W2 xn + b2) + b3 (erf GELU).  Weights bf16 [out, in], biases fp32,
 * b13 = b1 + b3.  fwd writes Gp = gelu'(z), H = gelu(z) (saved for the backward) and Y (all bf16 [R, Dout]); bwd
 * consumes dY (bf16) and writes dxn (nullable, bf16 [R, Din]) and fp32 weight / bias gradients (db13 = db1 = db3).
 * ysq (nullable, fp32 [R]): squared L2 norm of every (bf16-rounded) output row, accumulated in the epilogue of the output
 * GEMM -- the row norms the PACL scorer needs of its patch tensor (pacl.py:122) without another pass over Y.
 * clipk_linear_*: y = x W^T + b on bf16 rows, dx (nullable) bf16, dW / db fp32. */
/* apply_rope (PACL/model/pacl.py:147-181) on token rows [B*S, D]: pairs (x[2j], x[2j+1]) rotated by
 * the angle of (position = row % S, j), written de-interleaved (first halves, then second halves).  sin_t / cos_t: fp32
 * [S, D/2] tables built by the caller exactly as the reference builds them.  inverse != 0: transposed rotation. */
int clipk_rope(const void* x, int dtype_in, int64_t rows, int S, int D, const float* sin_t, const float* cos_t, void* y,
               int dtype_out, int inverse, void* stream);
int clipk_ln_fwd(const void* x, int dtype, int64_t rows, int D, const float* gamma, const float* beta, float eps,
                 void* xn, float* mean, float* rstd, float drop_p, uint64_t seed, void* keep_bits, void* stream);
size_t clipk_ln_bwd_workspace_bytes(int64_t rows, int D);
int clipk_ln_bwd(const void* x, int dtype, int64_t rows, int D, const float* gamma, const float* mean,
                 const float* rstd, const void* dxn, float* dgamma, float* dbeta, void* dx, float drop_p,
                 const void* keep_bits, void* workspace, size_t ws_bytes, void* stream);
int clipk_patch_proj_fwd(const void* xn, int64_t R, int Din, int Dout, const void* W1, const void* W2, const void* W3,
                         const float* b13, const float* b2, void* Gp, void* H, void* Y, float* ysq, void* stream);
size_t clipk_patch_proj_bwd_workspace_bytes(int64_t R, int Din, int Dout);
int clipk_patch_proj_bwd(const void* xn, const void* Gp, const void* H, const void* dY, int64_t R, int Din, int Dout,
                         const void* W1, const void* W2, const void* W3, void* dxn, float* dW1, float* dW2, float* dW3,
                         float* db13, float* db2, void* workspace, size_t ws_bytes, void* stream);
int clipk_linear_fwd(const void* x, int64_t R, int Din, int Dout, const void* W, const float* bias, void* y,
                     void* stream);
size_t clipk_linear_bwd_workspace_bytes(int64_t R, int Din, int Dout);
int clipk_linear_bwd(const void* x, const void* dY, int64_t R, int Din, int Dout, const void* W, void* dx, float* dW,
                     float* db, void* workspace, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Batched eval protocol: the accuracy accounting of PACL/eval_pacl.py on the device, fed by the
 * one-launch scorer (clipk_pacl_paired_fwd with v_div = K).
 *   clipk_eval_correct : correct[i] = score[i,0] > score[i,k] for all k > 0           (eval_pacl.py:53-57, :130-134)
 *   clipk_eval_whatsup : individual / pair / set counts over (set_id, rel_id) keys      (eval_pacl.py:59-104)
 *                        rel_id: 0 left, 1 right, 2 on, 3 under, 4 in-front, 5 behind; a later item with the same
 *                        key overwrites an earlier one; counts[8] = indiv lr/ou/fb, pair lr/ou/fb, sets, items
 *   clipk_eval_mmvp    : pairs of (2 images x 2 texts): pred / pair / single counts per category (eval_pacl.py:303-335) */
int clipk_eval_correct(const float* scores, int items, int K, int* correct, void* stream);
int clipk_eval_whatsup(const int* correct, const int* set_id, const int* rel_id, int items, int nsets, int* winner_ws,
                       int* counts, void* stream);
int clipk_eval_mmvp(const float* s_img1, const float* s_img2, const int* gt, int pairs, int pairs_per_cat, int ncat,
                    int* pred, int* counts, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Retrieval ranks fused into the logits GEMM (get_clip_metrics,
 * open_clip/src/open_clip_train/train.py:360-377): X bf16 [M,D], Y bf16 [N,D]; rank_row[i] = #{j != i : <x_i,y_j> >
 * <x_i,y_i>}, rank_col[j] = #{i != j : <x_i,y_j> > <x_j,y_j>} (0 = ground truth first); the logits are never written.
 * clipk_rank_stats: (sum of ranks, #rank<1, #rank<5, #rank<10) as uint64[4] on the device. */
int clipk_retrieval_ranks(const void* X, const void* Y, int M, int N, int D, float* diag_ws, int* rank_row,
                          int* rank_col, void* stream);
int clipk_rank_stats(const int* ranks, int n, unsigned long long* stats4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPK_H_ */
