// Small fp32 contractions of the CSA layer's glue (CUDA cores, fp32 FMA):
//   D[m][n] (+)= alpha * sum_k opA(m, k) * opB(n, k)
// with optional row gathers on the stored rows of A and B.  These are the pieces whose operands are a few hundred rows
// (per-chunk mean vectors, per-chunk gradient sums): far too small for a tensor-core pipeline, and they carry values
// that must not be rounded to 16 bits (V is centred on c = mean_chunk(X) Wv^T, and c Wo^T re-enters the pre-LayerNorm
// sum in fp32; MID-FC/csa_models.py:105,115 — Wv and fc are linear, so the mean passes through them exactly).
#include <stdint.h>

#include "host_util.h"

namespace csn {

struct SgemmArgs {
  const float* A; long long lda; const int* a_rows; int transA;   // transA: element (m, k) at A[row(k)*lda + m], else A[row(m)*lda + k]
  const float* B; long long ldb; const int* b_rows; int transB;   // transB: element (n, k) at B[row(k)*ldb + n], else B[row(n)*ldb + k]
  float* D; long long ldd;
  int M, N, K;
  float alpha; int accumulate;
  int k_per_split;   // grid.z splits of the contraction (partial sums are added atomically into D)
};

// 64 x 64 output tile per CTA, 256 threads, 4 x 4 outputs per thread (rows ty + 16 i, columns tx + 16 j), K staged 16
// at a time as As[k][m] / Bs[k][n] (padded rows): per k-step a thread reads 4 + 4 values for 16 FMAs.
__global__ void __launch_bounds__(256) sgemm_small_kernel(const SgemmArgs p) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int k_begin = blockIdx.z * p.k_per_split, k_end = min(p.K, k_begin + p.k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: 1024 elements per operand tile, 4 per thread; the contiguous global index runs along threadIdx
  auto load = [&](const float* X, long long ld, const int* rows, int trans, int mn0, int MN, float (*Xs)[65], int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int mn, k;
      if (trans) { mn = e & 63; k = e >> 6; }     // storage [k][mn]: mn contiguous
      else { k = e & 15; mn = e >> 4; }           // storage [mn][k]: k contiguous
      const int gmn = mn0 + mn, gk = k0 + k;
      float v = 0.f;
      if (gmn < MN && gk < k_end) {
        const int sr = trans ? gk : gmn;
        const long long row = rows ? __ldg(rows + sr) : sr;
        v = __ldg(X + row * ld + (trans ? gmn : gk));
      }
      Xs[k][mn] = v;
    }
  };
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    load(p.A, p.lda, p.a_rows, p.transA, m0, p.M, As, k0);
    load(p.B, p.ldb, p.b_rows, p.transB, n0, p.N, Bs, k0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= p.N) continue;
      float* d = p.D + (long long)m * p.ldd + n;
      const float v = p.alpha * acc[i][j];
      if (gridDim.z > 1) atomicAdd(d, v);
      else *d = p.accumulate ? *d + v : v;
    }
  }
}

}  // namespace csn

extern "C" int csn_sgemm_small(const float* A, int64_t lda, const int32_t* a_rows, int32_t transA, const float* B,
                               int64_t ldb, const int32_t* b_rows, int32_t transB, float* D, int64_t ldd, int32_t M,
                               int32_t N, int32_t K, float alpha, int32_t accumulate, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(A && B && D, "csn_sgemm_small: null pointer");
  CSN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "csn_sgemm_small: negative extent");
  if (M == 0 || N == 0) return 0;
  const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  // long contractions with few output tiles (d fc.weight correction: 16 tiles, K = blocks x chunks) are split over
  // grid.z, the partial sums added atomically — only when the caller accumulates into an initialised D anyway
  int split = 1;
  if (accumulate && tiles < 64 && K >= 256) split = (K + 127) / 128 < 16 ? (K + 127) / 128 : 16;
  const int kps = ((K + split - 1) / split + 15) / 16 * 16;
  split = (K + kps - 1) / kps;
  SgemmArgs a{A, lda, a_rows, transA, B, ldb, b_rows, transB, D, ldd, M, N, K, alpha, accumulate, kps};
  sgemm_small_kernel<<<dim3((N + 63) / 64, (M + 63) / 64, split), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  CSN_LAUNCH_OK("sgemm_small_kernel");
  return 0;
}
