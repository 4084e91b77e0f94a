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
 * aligned).  Batch index (b0,b1,b2) selects the sub-problem: its mn origin is sum_i b_i*mn_off[i],
 * its k origin sum_i b_i*k_off[i].  Reads outside [0,inner)x[0,outer) return zero.
 */
typedef struct csn_mat {
  const void* ptr;
  int32_t dtype; /* CSN_F16 or CSN_BF16; A and B must agree */
  int32_t major;
  int64_t inner, outer, ld;
  int64_t mn_off[3];
  int64_t k_off[3];
} csn_mat;

typedef struct csn_out {
  void* ptr;
  int32_t dtype;      /* CSN_F32 / CSN_F16 / CSN_BF16 */
  int32_t transposed; /* 0: D[m*ld + n], 1: D[n*ld + m] */
  int64_t ld;
  int64_t off[3];     /* element offset per batch index */
  int32_t accumulate; /* 1: atomically add into an fp32 D (required when split_k > 1) */
  int32_t reserved;
} csn_out;

int csn_gemm(const csn_mat* A, const csn_mat* B, const csn_out* D, int32_t M, int32_t N, int32_t K,
             const int32_t nb[3], float alpha, int32_t split_k, void* stream);

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

/* scores[q*ld + c] = (sum_{t<ntiles} partial[(q*ntiles+t)*n_cand + c]) / n_rows, summed in a fixed
 * order (csa_models.py:257 `.max(-1)[0].mean(-1)`). */
int csn_knn_reduce(const float* partial, float* scores, int32_t n_q, int32_t n_cand, int32_t ntiles,
                   int32_t n_rows, int64_t ld_scores, void* stream);

/* Per-row top-k (k <= 8), values and int64 indices sorted descending; equal scores: lower index
 * first (`retrieval_measure.topk(K+1, -1)`, csa_models.py:278,401). */
int csn_topk_rows(const float* scores, int64_t ld, int32_t n_rows, int32_t n_cols, int32_t k,
                  float* out_val, int64_t* out_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSN_B200_H_ */
