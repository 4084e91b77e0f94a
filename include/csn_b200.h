/*
 * csn_b200 — C ABI of the B200-native cross-shape-attention hot path.
 *
 * The reference (marios2019/CSN) is pure PyTorch and has no FFI/plugin layer: its boundary is the
 * nn.Module surface (MID-FC/csa_models.py:37-432, MinkowskiNet/models/attention.py:9-113,
 * MinkowskiNet/models/hrnet.py:359-490).  This header is what a maintainer binds (ctypes, see
 * INTEGRATION.md) to replace the ATen calls underneath those modules.  Conventions:
 *   - plain pointers and sizes only (no torch types); all pointers are DEVICE pointers unless a
 *     parameter is documented as host memory;
 *   - caller allocates every output and workspace; nothing is allocated inside;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *   - return value 0 = ok, non-zero = error; the message is available from csn_last_error();
 *   - re-entrant, no hidden global state except a cached driver entry point and the error string
 *     (thread-local).
 */
#ifndef CSN_B200_H_
#define CSN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types */
enum { CSN_F32 = 0, CSN_F16 = 1, CSN_BF16 = 2 };
/* operand majors for csn_gemm */
enum { CSN_MAJOR_K = 0, CSN_MAJOR_MN = 1 };

/* last error message of the calling thread ("" if none) */
const char* csn_last_error(void);
/* library/ABI version, bumped on any signature change */
int csn_abi_version(void);

/* Dropout epoch (one word per device, 0 at start): every launch's mask seed is drop_seed + epoch * 0x9E3779B9.  A
 * training step captured as a CUDA graph freezes its kernel arguments, seeds included; calling this on the same
 * stream before each replay makes the replay draw fresh masks (csn_b200.graphs.GraphedStep.replay(epoch=...)). */
int csn_set_drop_epoch(uint32_t epoch, void* stream);
/* number of kernels this library has launched since load (all streams); used by bench.py */
int64_t csn_launch_count(void);

/*
 * Strided, batched tensor-core contraction  D[b][m][n] = alpha * sum_k A[b][m][k] * B[b][n][k]
 * (tcgen05.mma kind::f16, fp32 accumulation in TMEM, operands fed by TMA).
 * Replaces every torch.matmul / nn.Linear on the path: csa_models.py:103-105 (projections), :139
 * (q k^T), :142 (attn v), :115 (fc), and their autograd transposes.
 *
 * An operand is a 2-D view of 16-bit elements.  major = CSN_MAJOR_K : view[mn][k], k contiguous;
 * major = CSN_MAJOR_MN : view[k][mn], mn contiguous (a transposed operand, consumed without a copy).
 * `inner`/`outer` are the extents of the contiguous / strided dimension of the whole view, `ld` the
 * stride of the strided dimension in elements (ld*2 bytes must be a multiple of 16, ptr 16-byte
 * aligned).  Batch index (b0,b1,b2,b3) selects the sub-problem: its mn origin is sum_i b_i*mn_off[i],
 * its k origin sum_i b_i*k_off[i].  Reads outside [0,inner)x[0,outer) return zero.
 */
typedef struct csn_mat {
  const void* ptr;
  int32_t dtype; /* CSN_F16 or CSN_BF16; A and B must agree */
  int32_t major;
  int64_t inner, outer, ld;
  int64_t mn_off[4];
  int64_t k_off[4];
} csn_mat;

typedef struct csn_out {
  void* ptr;
  int32_t dtype;      /* CSN_F32 / CSN_F16 / CSN_BF16 */
  int32_t transposed; /* 0: D[m*ld + n], 1: D[n*ld + m] */
  int64_t ld;
  int64_t off[4];     /* element offset per batch index */
  int32_t accumulate; /* 1: atomically add into an fp32 D (required when split_k > 1) */
  int32_t reserved;
} csn_out;

int csn_gemm(const csn_mat* A, const csn_mat* B, const csn_out* D, int32_t M, int32_t N, int32_t K,
             const int32_t nb[4], float alpha, int32_t split_k, void* stream);

/* Two contractions of the same shape (M, N > 128, K, batch extents) in ONE launch, their tiles interleaved batch by
 * batch: problem 0 = (A0 K-major, B0 MN-major) -> D0, problem 1 = (A1 MN-major, B1 MN-major) -> D1, both outputs 16-bit
 * row-major views of one buffer (D1->ptr inside D0's rows, same ld).  For the attention backward, dQ = dS K and
 * dK = dS^T Q (autograd of csa_models.py:139) read the same dS tile: issued back to back, the second read hits L2
 * instead of HBM. */
int csn_gemm_dual(const csn_mat* A0, const csn_mat* B0, const csn_out* D0, const csn_mat* A1, const csn_mat* B1,
                  const csn_out* D1, int32_t M, int32_t N, int32_t K, const int32_t nb[4], float alpha, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Shape-compatibility retrieval measure and top-K neighbour selection
 * (reference: CrossShapeAt.get_retrieval_measure / get_knn_graph, MID-FC/csa_models.py:244-280 and
 * :360-404; HRNetSimCSN.cosine_similarity, MinkowskiNet/models/hrnet.py:472-490; topk with
 * self-exclusion, MinkowskiNet/lib/csn_utils.py:91-96).
 * ------------------------------------------------------------------------------------------- */

/* out[r] = in[r] / max(|in[r]|_2, eps) as 16-bit rows (F.normalize, csa_models.py:253,255; eps = 0
 * gives the MinkowskiNet variant hrnet.py:475-480).  in: fp32 [rows][D], D in {128,256,384,512}. */
int csn_normalize_rows(const float* in, void* out, int64_t rows, int32_t D, float eps,
                       int32_t out_dtype, void* stream);

/* Work tables for csn_knn_scores (int32, device memory):
 *   items[i] = {q_row0, n_valid, list_begin, list_count, out_off, 0}: a tile of up to 128 query rows
 *              starting at row q_row0 of feat_q, of which n_valid belong to the query shape, scored
 *              against candidates cands[list_begin .. list_begin+list_count);
 *   cands[j] = {row0, len}: a candidate shape occupying rows [row0, row0+len) of feat_c.
 * partial[out_off + pos] = sum over the tile's valid rows p of max_q <feat_q[p], feat_c[row0+q]>.
 * feat_q / feat_c: unit-norm 16-bit rows of width 256 (csn_normalize_rows). The N_q x N_c cosine
 * matrix never leaves TMEM (the reference materialises it: csa_models.py:256). */
int csn_knn_scores(const void* feat_q, int64_t rows_q, const void* feat_c, int64_t rows_c,
                   int32_t dtype, const int32_t* items, int32_t n_items, const int32_t* cands,
                   float* partial, void* stream);

/* Exact variant for the pairs near the top-K boundary: operands are split into hi = fp16(x) and
 * lo = fp16((x - hi) * 2^11) (csn_normalize_rows_split), three MMAs per k-step reproduce the fp32
 * cosines of csa_models.py:256 to ~1e-7. Same work tables as csn_knn_scores. */
int csn_normalize_rows_split(const float* in, void* out_hi, void* out_lo, int64_t rows, int32_t D, float eps,
                             void* stream);
int csn_knn_scores_exact(const void* q_hi, const void* q_lo, int64_t rows_q, const void* c_hi, const void* c_lo,
                         int64_t rows_c, const int32_t* items, int32_t n_items, const int32_t* cands,
                         float* partial, void* stream);

/* scores[q*ld + c] = (sum_{t<ntiles} partial[(q*ntiles+t)*n_cand + c]) / n_rows, summed in a fixed
 * order (csa_models.py:257 `.max(-1)[0].mean(-1)`). */
int csn_knn_reduce(const float* partial, float* scores, int32_t n_q, int32_t n_cand, int32_t ntiles,
                   int32_t n_rows, int64_t ld_scores, void* stream);

/* Per-row top-k (k <= 64), values and int64 indices sorted descending; equal scores: lower index
 * first (`retrieval_measure.topk(K+1, -1)`, csa_models.py:278,401). */
int csn_topk_rows(const float* scores, int64_t ld, int32_t n_rows, int32_t n_cols, int32_t k,
                  float* out_val, int64_t* out_idx, void* stream);

/* Top-K boundary band on the device (no host round trip between csn_knn_scores and the final indices): for every
 * query row of `scores` the k-th largest value is found and every candidate within `margin` of it (or above) is
 * appended, in ascending candidate order, to that query's slice of the csn_knn_scores_exact work tables:
 *   band_idx[q*n_cols + i] = candidate index, cands[(q*n_cols + i)] = {row0, len} of that candidate, counts[q] = band
 *   size, items[item0[q] + t] = {q_row0[q] + 128 t, valid rows, q*n_cols, counts[q], (item0[q] + t)*n_cols, 0} for the
 *   row tiles t of query q (item0 = exclusive prefix sum of the queries' tile counts, built once per store).
 * csn_knn_scores_exact then runs over sum_q ceil(len_q/128) items with a partial buffer of that many rows of n_cols
 * floats, and csn_knn_band_patch sums the tiles of every re-scored pair (fixed order, fp64) into scores. */
int csn_knn_band_select(const float* scores, int64_t ld, int32_t n_q, int32_t n_cols, int32_t k, float margin,
                        const int32_t* q_row0, const int32_t* q_len, const int32_t* item0, const int32_t* c_row0,
                        const int32_t* c_len, int32_t* band_idx, int32_t* counts, int32_t* cands, int32_t* items,
                        void* stream);
int csn_knn_band_patch(const float* partial, const int32_t* band_idx, const int32_t* counts, const int32_t* item0,
                       const int32_t* q_len, int32_t n_q, int32_t n_cols, float* scores, int64_t ld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused attention core  O = softmax(Q K^T / sqrt(d)) V  (flash-style: the score matrix only exists as
 * 128x128 fp32 tiles in TMEM).  Replaces ScaledDotProductAttention.forward (csa_models.py:138-144,
 * MinkowskiNet/models/attention.py:69-75) for every (block, head, 128-row query tile) listed in the
 * work table `items` (int32 x 10 per item, device memory):
 *   {q_row0, q_valid, kv_row0, kv_len, o_row0, col0, lse_off, flags, v_row0, 0}
 * the tile's queries are rows [q_row0, q_row0+128) of Q (q_valid of them real), its keys/values rows
 * [kv_row0, kv_row0+kv_len) of K / V, all restricted to columns [col0, col0+d_head); the result goes to
 * rows [o_row0, ..) / columns [col0, ..) of O (16-bit) and lse[lse_off + r] = log sum_j exp(s_rj).
 * (V rows start at v_row0, normally == kv_row0).
 * The whole 128-row tile is written (TMA stores), rows >= q_valid as zeros: O lives in a padded layout
 * with o_rows rows in total.
 * Q, K, V: 16-bit row-major views of `width` columns with leading dimensions ldq / ldk / ldv.
 * d_head in {64, 256}.  O_lo (optional, same layout as O) receives the rounding residual of O,
 * (o - round16(o)) * 2^11 (fp16) or * 2^8 (bf16), which csn_attn_delta uses to form
 * delta = rowsum(dO o O) to ~22 bits (a 16-bit O alone costs 1e-3 on dQ/dK when all value rows share
 * a large mean, as post-ReLU features do).
 * paired != 0 asserts that items 2m and 2m+1 stream exactly the same K/V tiles (same kv_row0, kv_len,
 * col0, v_row0): they then run as a 2-CTA cluster whose TMA loads are multicast to both CTAs.
 * ------------------------------------------------------------------------------------------- */
/* Training-mode dropout (nn.Dropout on the probabilities, csa_models.py:141 / attention.py:72; on the fc output,
 * csa_models.py:115 / attention.py:51): drop_p in [0, 1), 0 = off.  The mask is a pure function of (drop_seed, row id,
 * column) — csrc/attn_common.cuh — regenerated by every kernel that needs it, never stored; pass the same seed and p to
 * csn_attn_fwd, csn_attn_bwd_dv and csn_attn_bwd_dq of one step (row id = index into lse, column = key index inside
 * the chunk; the dV work items carry the tile's first key in their last field), and to csn_gemm_res_ln /
 * csn_add_ln_fwd and csn_ln_bwd for the fc output (row id = row of Z, column = channel).  lse is always the
 * log-sum-exp of the UNdropped scores. */
int csn_attn_fwd(const void* Q, const void* K, const void* V, int64_t q_rows, int64_t kv_rows, int64_t width,
                 int64_t ldq, int64_t ldk, int64_t ldv, int32_t d_head, int32_t dtype, const int32_t* items,
                 int32_t n_items, void* O, int64_t o_rows, int64_t ldo, float* lse, void* O_lo, int32_t paired,
                 uint32_t drop_seed, float drop_p, void* stream);

/* Key-stationary attention backward for d_head = 64: dK = dS^T Q and dV = P^T dO of one 128-key tile per item in
 * ONE pass, P and dS recomputed per tile from K, V, Q, dO, lse and delta and never written to memory (autograd of
 * ScaledDotProductAttention.forward, MinkowskiNet/models/attention.py:69-75, MID-FC/csa_models.py:138-144, which
 * keep the (B, h, Lq, Lk) attention matrix).  With csn_attn_bwd_dq(dS = NULL) this is the whole attention
 * backward in O(L) memory.  items: n_items x 12 int32 {k_row0, k_valid, q_row0, q_len, o_row0, col0, stat_off,
 * do_row0, key0, 0, 0, 0}: resident key rows [k_row0, +128) of the K / V views (k_valid of them exist), streamed
 * query rows [q_row0, q_row0 + q_len) of the Q view and [do_row0, ...) of dO, lse / delta of streamed query n at
 * [stat_off + n], outputs at rows [o_row0, +128) x columns [col0, +64) of dK and dV (rows >= k_valid: zeros).
 * key0 = index of the first resident key inside its chunk (dropout mask column, see the dropout paragraph).
 * lse / delta hold n_stats values (a multiple of 4); stat_scratch (2 * n_stats floats, 16B aligned) receives their
 * pre-scaled copies (lse * log2 e | delta / sqrt d), which the kernel's producer warp bulk-copies tile by tile. */
int csn_attn_bwd_dkv(const void* K, const void* V, const void* Q, const void* dO, int64_t kv_rows, int64_t q_rows,
                     int64_t do_rows, int64_t width, int64_t ldk, int64_t ldv, int64_t ldq, int64_t lddo,
                     int32_t d_head, int32_t dtype, const int32_t* items, int32_t n_items, void* dK, void* dV,
                     int64_t out_rows, int64_t ldout, const float* lse, const float* delta, int64_t n_stats,
                     float* stat_scratch, uint32_t drop_seed, float drop_p, void* stream);

/* Backward of the attention core (autograd of csa_models.py:139-142), three pieces:
 *  csn_attn_delta : delta[(blk*h+head)*rows_pad + r] = sum_c dO[blk*rows_pad+r][head*d+c] * (O + O_lo/2^11)[..]
 *  csn_attn_bwd_dv: dV = P^T dO, P^T rebuilt from K, Q and the forward lse. Same kernel and item layout
 *                   as csn_attn_fwd with the roles swapped: the resident 128-row tile holds KEY rows
 *                   (q_row0/q_valid), kv_row0/kv_len address the QUERY rows (K-major operand), v_row0 the
 *                   dO rows, lse_off the lse of the first query row, o_row0/col0 the dV tile.
 *  csn_attn_bwd_dq: per 128-row query tile: S = Q K^T, dP = dO V^T, dS = P o (dP - delta) / sqrt(d),
 *                   dQ += dS K; dS is also written (16-bit, [ds_row0 + r][ds_col0 + key]) so that
 *                   dK = dS^T Q runs as one csn_gemm.  dQ == NULL: only dS is produced (S|dP are then
 *                   double-buffered in TMEM) and dQ = dS K is left to csn_gemm as well.  dS == NULL (d_head 64,
 *                   dQ != NULL): nothing but dQ is written, dS stays in TMEM as the A operand of dQ += dS K
 *                   (dK and dV then come from csn_attn_bwd_dkv).  items: int32 x 12
 *   {q_row0, q_valid, kv_row0, kv_len, o_row0 (rows of dO and dQ), col0, stat_off, ds_row0, ds_col0, flags, 0, 0}. */
int csn_attn_delta(const void* dO, const void* O, const void* O_lo, float* delta, int64_t rows, int32_t rows_pad, int32_t n_head,
                   int32_t d_head, int64_t ld, int32_t dtype, void* stream);
int csn_attn_bwd_dv(const void* Kres, const void* Qstr, const void* dO, int64_t k_rows, int64_t q_rows,
                    int64_t do_rows, int64_t width, int64_t ldk, int64_t ldq, int64_t lddo, int32_t d_head, int32_t dtype,
                    const int32_t* items, int32_t n_items, void* dV, int64_t dv_rows, int64_t lddv, const float* lse,
                    int32_t paired, uint32_t drop_seed, float drop_p, void* stream);
int csn_attn_bwd_dq(const void* Q, const void* dO, const void* K, const void* V, int64_t q_rows, int64_t do_rows,
                    int64_t kv_rows, int64_t width, int64_t ldq, int64_t lddo, int64_t ldk, int64_t ldv,
                    int32_t d_head, int32_t dtype, const int32_t* items, int32_t n_items, void* dQ, int64_t lddq,
                    void* dS, int64_t ds_rows, int64_t ldds, const float* lse, const float* delta, int32_t paired,
                    uint32_t drop_seed, float drop_p, void* stream);

/* Fused segmentation epilogue: logits = W f (the bias-free 1x1 `logit` conv, csa_models.py:201), masked
 * cross-entropy (csa_training.py:94-108: mean over the points whose label != ignore_index) and its backward in one
 * pass over the channel-major activation feat (element (b, k, n) at feat[b*b_stride + k*ch_stride + n], 256 channels):
 *   *n_valid (zero-initialised device int) <- number of unmasked points;  loss_part[b*ceil(n_points/128) + i] <-
 *   partial sums of -log p[label] (the loss is their sum / n_valid);  dlogits (optional, [B][C][n_points]) <-
 *   (softmax - onehot) / n_valid, zero at masked points;  dfeat (optional, feat's layout) <- W^T dlogits;
 *   both multiplied by *grad_scale when that optional device scalar (the upstream gradient of the loss) is given.
 * W: [n_classes][256] fp32, n_classes <= 64; labels int64 [B][n_points]. */
int csn_seg_loss(const float* feat, int64_t b_stride, int64_t ch_stride, int32_t n_batch, int32_t n_points,
                 const float* W, int32_t n_classes, const int64_t* labels, int32_t ignore_index, int32_t* n_valid,
                 float* loss_part, float* dlogits, float* dfeat, const float* grad_scale, void* stream);

/* Fused CSA head, forward AND backward in one pass over the attention blocks' pre-LayerNorm rows Z
 * (row r of block j at Z[(j*rows_pad + r)*256], statistics mean/rstd per row, LayerNorm affine gamma/beta):
 *   y[b][r] = sum_k w[b*n_k+k] * LayerNorm(Z[blk[b*n_k+k]][r])      the weighted sum, csa_models.py:232-238
 *   logits = W y (csa_models.py:201);  masked cross-entropy and accuracy (csa_training.py:94-108);  per-class IoU
 *   counters (csa_training.py:110-134);  and, when dOutT != NULL, the backward of all of it:
 *   dOutT[b*rows_pad + r] = W^T dlogits (row-major padded rows, zero in pad rows / masked points: what csn_ln_bwd
 *   reads through src_idx),  *amax = max|dOutT|,  dcomp[b*n_k+k] += <dOutT[b], LayerNorm(Z[blk])> (fp64: it is
 *   consumed through differences of nearly equal numbers),
 *   dW[c][ch] = sum dlogits[c] y[ch]  (dW_part: [csn_csa_head_grid()][n_classes][256] scratch, reduced in fixed order).
 * Zero-initialised by the caller: n_valid (1 int), stats (3*n_classes + 2 ints: #pred==c, #label==c, #both, then
 * #correct, #labels outside [0, n_classes)), amax (1 float), dcomp, loss_part.  loss_part: csn_csa_head_grid() floats
 * (one partial sum per persistent CTA) whose sum divided by *n_valid is the loss.  y_out (optional): the combined features as padded rows.  labels: int64,
 * (b, point n) at labels[b*lab_stride + n]; n_k <= 6; n_classes <= 64.
 * Replaces csn_combine_fwd + the ATen conv/log-softmax/nll kernels and their backwards + csn_pack_rows(dOut) +
 * csn_block_dot.  loss (optional, 1 float): receives sum(loss_part) / max(*n_valid, 1). */
int csn_csa_head_grid(int32_t n_b, int32_t rows_pad);
int csn_csa_head(const float* Z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 const int32_t* blk, const float* w, int32_t n_b, int32_t n_k, const float* W, int32_t n_classes,
                 const int64_t* labels, int64_t lab_stride, int32_t ignore_index, int32_t* n_valid, int32_t n_points,
                 int32_t chunk, int32_t chunk_pad, int32_t rows_pad, float* loss_part, float* dOutT, float* amax,
                 double* dcomp, float* dW_part, float* dW, int32_t* stats, float* y_out, float* loss, void* stream);

/* Shape-compatibility glue (csa_models.py:222-230), forward and backward (csrc/compat.cu), fp64 inside:
 *   u_q[b] = normalize(Wq pooled[b*K1] + bq),  u_k[r] = normalize(Wk y_stack[r] + bk) with the stacked rows
 *   y_stack[r] = pooled[(r % B)*K1 + r / B] (rows ordered [k][b], :213,220) and the (B, K1, 256) VIEW of :227:
 *   comp[b][k] = softmax_k(u_q[b] . u_k[b*K1 + k]).  pooled: fp32 [B*K1][256], slot (b, k) at row b*K1 + k.
 * csn_compat_bwd: dcomp (fp64 [B*K1]) times the optional device scalar *gscale ->  d weights / biases (fp32), dpool
 * (fp32 [B*K1][256], gradient of the pooled descriptors) and *dpool_amax = max |dpool| (optional, zero-initialised).
 * u_q [B][256], u_k [B*K1][256], n_q [B], n_k [B*K1], comp64 [B*K1], dlin_q / dlin_k: fp64 scratch kept by the caller. */
int csn_compat_fwd(const float* pooled, const float* Wq, const float* bq, const float* Wk, const float* bk, int32_t B,
                   int32_t K1, double* u_q, double* u_k, double* n_q, double* n_k, double* comp64, float* comp,
                   void* stream);
/* cw[j] = comp[cw_index[j]] * *gscale (0 for cw_index[j] < 0): the per-block fan-out weights csn_ln_bwd takes as src_w;
 * *amax_out = *amax_in * |*gscale| + *dpool_amax * inv_points: the bound of the upstream gradient it scales by.
 * gscale / dpool_amax may be NULL (1 / 0). */
int csn_compat_fanout(const float* comp, const int32_t* cw_index, int32_t n_blocks, const float* gscale,
                      const float* amax_in, const float* dpool_amax, float inv_points, float* cw, float* amax_out,
                      void* stream);
int csn_compat_bwd(const float* pooled, const float* Wq, const float* Wk, const double* u_q, const double* u_k,
                   const double* n_q, const double* n_k, const double* comp64, const double* dcomp, const float* gscale,
                   int32_t B, int32_t K1, double* dlin_q, double* dlin_k, float* dWq, float* dbq, float* dWk, float* dbk,
                   float* dpool, float* dpool_amax, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HBM-bound pieces of the CSA/SSA layer.  Row buffers use "padded" coordinates: a shape's block has
 * rows_pad = n_chunks*chunk_pad rows; chunk c (points [c*chunk, (c+1)*chunk)) occupies rows
 * [c*chunk_pad, c*chunk_pad + chunk); pad rows are zero.  MID-FC: chunk = 500 (csa_models.py:83-90),
 * chunk_pad = 512.
 * ------------------------------------------------------------------------------------------- */

/* Channel-major fp32 features (the reference layout (B,256,N,1), csa_models.py:92-94) -> padded
 * row-major rows, 16-bit (dst16) and optionally fp32 (dst32, may be NULL).  Source element
 * (i0,i1,c,n) is src[i0*src_s0 + i1*src_s1 + c*ch_stride + n]; destination slot is
 * dst_slot0 + i0*dst_s0 + i1*dst_s1.  dst16 may be NULL (fp32 transpose only); amax (optional, device,
 * zero-initialised) receives max |src|; chunk_sum (optional, device, zero-initialised,
 * [(slot*n_chunks + chunk)][256]) receives the channel sums over the valid points of every chunk. */
int csn_pack_rows(const float* src, void* dst16, float* dst32, int64_t ch_stride, int32_t n0,
                  int64_t src_s0, int32_t n1, int64_t src_s1, int64_t dst_slot0, int64_t dst_s0,
                  int64_t dst_s1, int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad,
                  int32_t dtype, float* amax, float* chunk_sum, void* stream);

/* The same packing from a 16-bit channel-major source (src_dtype = CSN_F16 / CSN_BF16; strides in elements): for
 * callers that keep their feature collection as a 16-bit pinned host cache, which halves the host-to-device bytes
 * of a step (MID-FC/features_data_loader.py:124-140 reads fp32 .npy files per step). */
int csn_pack_rows_src16(const void* src, int32_t src_dtype, void* dst16, float* dst32, int64_t ch_stride, int32_t n0,
                        int64_t src_s0, int32_t n1, int64_t src_s1, int64_t dst_slot0, int64_t dst_s0, int64_t dst_s1,
                        int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, int32_t dtype,
                        float* amax, float* chunk_sum, void* stream);

/* P = softmax over the first cols_valid columns of each fp32 row (F.softmax(dim=-1),
 * csa_models.py:141), 16-bit, zero in pad columns and in pad rows (row % group_rows >= rows_valid). */
int csn_softmax_fwd(const float* S, void* P, int64_t rows, int32_t cols_pad, int32_t cols_valid,
                    int32_t group_rows, int32_t rows_valid, int32_t dtype, void* stream);
/* dS = P o (dP - rowsum(P o dP)) * scale   (autograd of csa_models.py:139-141). */
int csn_softmax_bwd(const void* P, const float* dP, void* dS, int64_t rows, int32_t cols_pad,
                    int32_t cols_valid, int32_t group_rows, int32_t rows_valid, float scale,
                    int32_t dtype, void* stream);

/* z = Z + residual (in place); Y = LayerNorm(z)*gamma + beta with eps inside the sqrt and biased
 * variance (csa_models.py:116-118, :57); mean / rstd saved per row; colsum (optional, [rows/64][256])
 * receives per-64-row partial sums of Y over valid rows, which csn_colsum_reduce adds in a fixed order
 * into the mean over points of csa_models.py:212,219 (deterministic, no atomics).
 * Residual row of (block, r) is R[(res_block[block]*block_rows + r)*256]. */
int csn_add_ln_fwd(float* Z, const float* R, const int32_t* res_block, float* Y, void* Y16, float* mean,
                   float* rstd, const float* gamma, const float* beta, float* colsum, int64_t rows,
                   int32_t block_rows, int32_t group_rows, int32_t rows_valid, float eps, int32_t dtype,
                   uint32_t drop_seed, float drop_p, void* stream);
int csn_colsum_reduce(const float* part, float* out, int32_t n_blocks, int32_t parts, float scale, void* stream);

/* The same contraction for plain problems (K-major A [M x K] and B [N x K], one batch, no split-K) on CTA PAIRS:
 * one tcgen05.mma.cta_group::2 of M = 256 per two SMs, each CTA holding its 128 rows of A and half of every B
 * slice (a 6-deep operand ring instead of 4, half the B operand read per SM).  lda / ldb / ldd in elements. */
int csn_gemm_pair(const void* A, const void* B, void* D, int32_t M, int32_t N, int32_t K, int64_t lda, int64_t ldb,
                  int64_t ldd, int32_t dtype, int32_t out_dtype, float alpha, void* stream);

/* The output projection fused with the residual add and the LayerNorm statistics (csa_models.py:115-118 in one
 * pass): Z[m][0..256) = alpha * (A B^T)[m] + residual(m), mean[m] / rstd[m] = LayerNorm statistics of that row
 * (biased variance, eps inside the square root). Z (fp32, row-major, leading dimension ldz) is the only
 * activation written: consumers re-normalise it on the fly (csn_combine_fwd, csn_block_dot, csn_ln_bwd).
 * A: [M x K] K-major 16-bit, B: [256 x K] K-major (nn.Linear weight layout).
 * Residual: row m = block*block_rows + g*group_rows + j (j < rows_valid, else a pad row: Z = 0, mean = rstd = 0)
 * is point n = g*rows_valid + j of the block's source shape, read from the reference's own channel-major layout
 * (csa_models.py:92-94,214-216) through TMA: each of the two residual tensors res0 / res1 (the query tensor and
 * the neighbour tensor) is viewed as a 2-D fp32 matrix [res*_rows][res_ld] whose row r holds one channel of one
 * shape; the [256][n_points] matrix of `block` starts at row res_row[block] of tensor res_sel[block].  Points
 * >= n_points are ignored.  The fp32 row-major copy of the inputs that csn_add_ln_fwd needs is never made.
 * block_rows and group_rows must be multiples of 32; res_ld*4 bytes a multiple of 16.
 * zbias (optional, fp32 [M / group_rows][256]): a row vector per row group (chunk) added to z before the statistics —
 * the fc image of the per-chunk value mean when V is centred (csn_gemm_colbias). */
int csn_gemm_res_ln(const csn_mat* A, const csn_mat* B, float* Z, int64_t ldz, int32_t M, int32_t K, float alpha,
                    const float* res0, int64_t res0_rows, const float* res1, int64_t res1_rows,
                    const int32_t* res_sel, const int32_t* res_row, int64_t res_ld, int32_t n_points,
                    int32_t block_rows, int32_t group_rows, int32_t rows_valid, float eps, float* mean,
                    float* rstd, const float* zbias, uint32_t drop_seed, float drop_p, void* stream);
/* csn_gemm (one batch, no split-K, row-major non-accumulating D) with a per-row-group column bias subtracted in the
 * epilogue, in fp32 before the output is rounded:
 *   D[m][n] = alpha * (A B^T)[m][n] - bias[(m / group_rows)*bias_ld + n - col0]   for n >= col0, m % group_rows < rows_valid.
 * Used for the Q|K|V projection with V centred on its per-chunk key mean c = mean_chunk(X) Wv^T: softmax rows sum to
 * one, so attn (V - c) + c == attn V (csa_models.py:142) exactly, while the 16-bit V - c, the 16-bit attention output
 * and delta = rowsum(dO o O) lose ~10x less to rounding (post-ReLU features give V a large common mean). */
int csn_gemm_colbias(const csn_mat* A, const csn_mat* B, const csn_out* D, int32_t M, int32_t N, int32_t K, float alpha,
                     const float* bias, int64_t bias_ld, int32_t col0, int32_t group_rows, int32_t rows_valid,
                     void* stream);
/* Small fp32 contraction on CUDA cores (fixed summation order): D[m][n] (+)= alpha * sum_k opA(m,k) * opB(n,k), with
 * opA(m,k) = transA ? A[ra(k)*lda + m] : A[ra(m)*lda + k], ra(i) = a_rows ? a_rows[i] : i (same for B).  For the
 * few-hundred-row pieces of the layer that must stay fp32: the per-chunk value means and their fc image, and the
 * matching correction of d fc.weight. */
int csn_sgemm_small(const float* A, int64_t lda, const int32_t* a_rows, int32_t transA, const float* B, int64_t ldb,
                    const int32_t* b_rows, int32_t transB, float* D, int64_t ldd, int32_t M, int32_t N, int32_t K,
                    float alpha, int32_t accumulate, void* stream);
/* dO = alpha * A B^T (16-bit, [M x N], N = n_head*d_head) fused with the attention backward's
 * delta[(blk*n_head + head)*rows_pad + r] = sum_c dO[m][head*d + c] * (O[m][head*d + c] + O_lo[..]/2^11 (fp16) or /2^8 (bf16)),
 * m = blk*rows_pad + r: the row-wise dot product that csn_attn_delta computes in a separate pass over dO, O and
 * O_lo (autograd of csa_models.py:115,142).  O / O_lo: 16-bit row-major [M x N] with leading dimension ldo, O_lo
 * may be NULL.  A: K-major [M x K]; B: [N x K] in either major; d_head 64 or 256. */
int csn_gemm_delta(const csn_mat* A, const csn_mat* B, void* dO, int64_t lddo, int32_t M, int32_t N, int32_t K,
                   float alpha, const void* O, const void* O_lo, int64_t ldo, float* delta, int32_t rows_pad,
                   int32_t n_head, int32_t d_head, void* stream);
/* colsum[(row/64)][c] = sum over the valid rows of the 64-row group of (Z - mean)*rstd*gamma + beta: the partials
 * csn_colsum_reduce turns into the pooled means, for rows whose statistics came from csn_gemm_res_ln. */
int csn_ln_colsum(const float* Z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  float* colsum, int64_t rows, int32_t block_rows, int32_t group_rows, int32_t rows_valid, void* stream);
/* dZ may be NULL (only the 16-bit copy is written).  bcast (optional, [n][256]) adds the row
 * bcast_scale * bcast[bcast_idx[block]] to every valid row of a block before the backward formula: the
 * gradient of the pooled mean (csa_models.py:212,219) without materialising it.
 * src_idx/src_w (optional, per block): the upstream gradient of block j is src_w[j] * dY[src_idx[j]] (dY then
 * holds one block of rows per SOURCE, e.g. the transposed output gradient of each batch item) or zero if
 * src_idx[j] < 0 — the compatibility-weighted fan-out of csa_models.py:232-238 is never materialised.
 * amax (optional, device): dY is multiplied by the power of two 2^floor(log2(128 / *amax)) on load so
 * that 16-bit gradient intermediates stay in the normal fp16 range; the caller divides the results.
 * chunk_gsum (optional, device, zero-initialised, [rows / group_rows][256]): column sums of dZ per row group. */
int csn_ln_bwd(const float* dY, const float* Z, const float* mean, const float* rstd, const float* gamma,
               float* dZ, void* dZ16, float* dgamma, float* dbeta, int64_t rows, int32_t block_rows,
               int32_t group_rows, int32_t rows_valid, int32_t dtype, const float* amax, const float* bcast,
               const int32_t* bcast_idx, float bcast_scale, const int32_t* src_idx, const float* src_w,
               float* chunk_gsum, uint32_t drop_seed, float drop_p, void* stream);

/* x[i] *= 1 / 2^floor(log2(128 / *amax)), i < n: undoes csn_ln_bwd's power-of-two scaling on the flat buffer that holds
 * every parameter gradient of a step. */
int csn_grad_unscale(float* x, int64_t n, const float* amax, void* stream);

/* Concatenated rows -> zero-padded slots of n_pad rows, 256 columns: out[s][r] = x[offsets[s] + r] for
 * r < offsets[s+1] - offsets[s], 0 beyond; out32 (fp32) and / or out16 (dtype f16 / bf16).  The layout change in
 * front of the ragged attention batch (the reference calls MultiHeadAttention once per shape, hrnet.py:370-417).
 * amax (optional, zero-initialised): receives max |x| (the loss scale of the backward pass). */
int csn_ragged_pad(const float* x, const int64_t* offsets, int32_t n_slots, int32_t n_pad, float* out32, void* out16,
                   int32_t dtype, float* amax, void* stream);

/* dst[s] = (dst[s] + sum over j < n_src with dst_block[j] == s of src[j]) * u, blocks of block_elems fp32 values
 * (at most 256 sources per destination); u = csn_grad_unscale's factor when amax != NULL, else 1.  The residual path
 * of the attention backward: `q = self.layer_norm(q + residual)` (MID-FC/csa_models.py:113, attention.py:53) sends
 * d z of every attention block to the input gradient of its query shape.  Deterministic (gather, fixed order). */
int csn_block_add(const float* src, const int32_t* dst_block, int32_t n_src, float* dst, int32_t n_dst, int64_t block_elems,
                  const float* amax, void* stream);

/* out[s][c] = mean over rows offsets[s] <= r < offsets[s+1] of x[r][c] (x row-major, n_cols wide; empty segment -> 0):
 * the per-shape global descriptors `feats.mean(dim=0)` of MinkowskiNet/models/hrnet.py:378,388 for a ragged batch of
 * shapes in one launch.  Deterministic (fixed summation order). */
int csn_segment_mean(const float* x, const int64_t* offsets, int32_t n_seg, int32_t n_cols, float* out, void* stream);

/* out[b][c][n] = sum_k w[b*n_k+k] * Y[blk[b*n_k+k]][padrow(n)][c]: the compatibility-weighted sum
 * of the self- and cross-attention outputs written back channel-major (csa_models.py:232-240);
 * rows16 (optional) receives a 16-bit row-major copy. With n_k = 1, w = 1 it is the plain
 * row-major -> channel-major transpose of csa_models.py:206.  When ln_mean != NULL, `Y` holds the
 * PRE-LayerNorm rows z and y = (z - mean)*rstd*gamma + beta is formed on the fly (the LayerNorm output is
 * then never written to HBM). */
int csn_combine_fwd(const float* Y, const int32_t* blk, const float* w, float* out, void* rows16,
                    int32_t n_b, int32_t n_k, int64_t out_b_stride, int64_t out_ch_stride,
                    int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, int32_t dtype,
                    const float* ln_mean, const float* ln_rstd, const float* ln_gamma, const float* ln_beta,
                    void* stream);
/* out[out_idx[j]] += sum over the valid rows r of block j of <G[src_idx[j]][r], y_j[r]> (y_j = Z rows of
 * block j, re-normalised on the fly when ln_mean != NULL): the gradient of the compatibility weights,
 * d comp[b,k] = <dOut[b]^T, MHA(x, x_k)> (autograd of csa_models.py:233,238), from the transposed output
 * gradient G. */
int csn_block_dot(const float* G, const float* Z, const int32_t* src_idx, const int32_t* out_idx, float* out,
                  int64_t rows, int32_t block_rows, int32_t group_rows, int32_t rows_valid, const float* ln_mean,
                  const float* ln_rstd, const float* ln_gamma, const float* ln_beta, void* stream);
/* Backward of the above plus of the pooled means.  For block j:
 *   dY[j] = cw[j]*dOut[cb[j]]^T (if cb[j] >= 0) + pool_scale*dpool[pb[j]] (if pb[j] >= 0), valid rows;
 *   dcomp[cw_index[j]] += <dOut[cb[j]]^T, Y[j]>  (if cw_index[j] >= 0).
 * dY == NULL: only dcomp is computed; dcomp == NULL: only dY.  amax (optional, device, zero-initialised)
 * receives max |dY|. */
int csn_combine_bwd(const float* dOut, const float* Y, const float* dpool, const int32_t* cb,
                    const float* cw, const int32_t* cw_index, const int32_t* pb, float pool_scale,
                    float* dY, float* dcomp, int32_t n_blocks, int64_t out_b_stride,
                    int64_t out_ch_stride, int32_t n_points, int32_t chunk, int32_t chunk_pad,
                    int32_t rows_pad, float* amax, const float* ln_mean, const float* ln_rstd,
                    const float* ln_gamma, const float* ln_beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSN_B200_H_ */
