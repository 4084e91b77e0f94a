// Small fp32 contractions of the CSA layer's glue (CUDA cores, fp32 FMA, fixed summation order):
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
};

// 32 x 32 output tile per CTA, 256 threads (4 outputs each: rows ty, ty+8, ty+16, ty+24), K staged 32 at a time.
__global__ void __launch_bounds__(256) sgemm_small_kernel(const SgemmArgs p) {
  __shared__ float As[32][33];   // [k][m]
  __shared__ float Bs[32][33];   // [k][n]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < p.K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = ty + 8 * i;   // strided index, tx = contiguous index
      // A
      {
        const int m = p.transA ? m0 + tx : m0 + a, k = p.transA ? k0 + a : k0 + tx;
        float v = 0.f;
        if (m < p.M && k < p.K) {
          const int sr = p.transA ? k : m;
          const long long row = p.a_rows ? __ldg(p.a_rows + sr) : sr;
          v = __ldg(p.A + row * p.lda + (p.transA ? m : k));
        }
        if (p.transA) As[a][tx] = v; else As[tx][a] = v;
      }
      {
        const int n = p.transB ? n0 + tx : n0 + a, k = p.transB ? k0 + a : k0 + tx;
        float v = 0.f;
        if (n < p.N && k < p.K) {
          const int sr = p.transB ? k : n;
          const long long row = p.b_rows ? __ldg(p.b_rows + sr) : sr;
          v = __ldg(p.B + row * p.ldb + (p.transB ? n : k));
        }
        if (p.transB) Bs[a][tx] = v; else Bs[tx][a] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = Bs[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += As[k][ty + 8 * i] * b;
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n >= p.N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i;
    if (m >= p.M) continue;
    float* d = p.D + (long long)m * p.ldd + n;
    *d = p.accumulate ? *d + p.alpha * acc[i] : p.alpha * acc[i];
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
  SgemmArgs a{A, lda, a_rows, transA, B, ldb, b_rows, transB, D, ldd, M, N, K, alpha, accumulate};
  sgemm_small_kernel<<<dim3((N + 31) / 32, (M + 31) / 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  CSN_LAUNCH_OK("sgemm_small_kernel");
  return 0;
}
