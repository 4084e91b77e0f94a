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

#ifdef __cplusplus
}
#endif
#endif /* CSN_B200_H_ */
