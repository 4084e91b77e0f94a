// Shape-compatibility glue of the CSA layer (MID-FC/csa_models.py:222-230), forward and backward, as five small kernels
// instead of ~45 ATen launches per step:
//   u_q[b]   = normalize(W_q y_q[b] + b_q)            y_q[b] = pooled SSA descriptor of query b          (:222-223)
//   u_k[r]   = normalize(W_k y_stack[r] + b_k)        y_stack rows ordered [k = 0: b = 0..B-1; k = 1: ...]  (:213,220,224-226)
//   comp[b]  = softmax_k( u_q[b] . u_k[b*(K+1) + k] )  — the (B, K+1, D) VIEW of the stacked rows (:227-230): entry (b, k)
//              of the view is stack row r = b*(K+1)+k, i.e. neighbour index r / B of batch item r % B.
// pooled is stored slot-major: slot s = b*(K+1) + k  ->  stack row r holds pooled[(r % B)*(K+1) + r / B].
// Everything is carried in fp64: the backward of this softmax subtracts d comp values that agree in their first 3-5
// digits, and the tensors are tiny (B(K+1) x 256).  F.normalize: v / max(|v|, 1e-12).
#include <stdint.h>

#include "host_util.h"
#include "ptx.cuh"

namespace csn {

constexpr int CD = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int stack_slot(int r, int B, int K1) { return (r % B) * K1 + r / B; }

// grid = S + B CTAs: CTA r < S -> u_k[r]; CTA S + b -> u_q[b].  256 threads; warp w computes outputs [32w, 32w+32),
// lanes stride the 256 inputs (coalesced weight rows), shuffle reduction.
__global__ void __launch_bounds__(256) compat_linear_kernel(const float* __restrict__ pooled, const float* __restrict__ Wq,
                                                           const float* __restrict__ bq, const float* __restrict__ Wk,
                                                           const float* __restrict__ bk, int B, int K1, double* __restrict__ u_q,
                                                           double* __restrict__ u_k, double* __restrict__ n_q, double* __restrict__ n_k) {
  __shared__ double y[CD];
  __shared__ double lin[CD];
  __shared__ double red[8];
  const int S = B * K1, r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_q = r >= S;
  const int slot = is_q ? (r - S) * K1 : stack_slot(r, B, K1);
  const float* W = is_q ? Wq : Wk;
  const float* bias = is_q ? bq : bk;
  y[tid] = (double)__ldg(pooled + (long long)slot * CD + tid);
  __syncthreads();
  // the 256-long dot products run in fp32 (operands are fp32; the cancellation this file guards against sits in the
  // softmax backward, not here); eight outputs in flight per warp: independent load / FMA / shuffle chains
  float yv[CD / 32];
#pragma unroll
  for (int i = 0; i < CD / 32; ++i) yv[i] = (float)y[lane + 32 * i];
  for (int o0 = warp * 32; o0 < warp * 32 + 32; o0 += 8) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* w = W + (long long)(o0 + j) * CD;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < CD / 32; ++i) a = fmaf(__ldg(w + lane + 32 * i), yv[i], a);
      s[j] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    if (lane < 8) {
      float v = s[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) v = lane == j ? s[j] : v;
      lin[o0 + lane] = (double)v + (double)__ldg(bias + o0 + lane);
    }
  }
  __syncthreads();
  double sq = warp_sum_d(lin[tid] * lin[tid]);
  if (lane == 0) red[warp] = sq;
  __syncthreads();
  double nrm = 0.0;
#pragma unroll
  for (int w2 = 0; w2 < 8; ++w2) nrm += red[w2];
  nrm = sqrt(nrm);
  const double den = nrm > 1e-12 ? nrm : 1e-12;
  double* u = is_q ? u_q + (long long)(r - S) * CD : u_k + (long long)r * CD;
  u[tid] = lin[tid] / den;
  if (tid == 0) (is_q ? n_q[r - S] : n_k[r]) = den;
}

// one CTA of 256 threads: warp per (b, k) dot product, then the softmax over k by the first B threads
__global__ void __launch_bounds__(256) compat_softmax_kernel(const double* __restrict__ u_q, const double* __restrict__ u_k, int B,
                                                            int K1, double* __restrict__ comp64, float* __restrict__ comp) {
  extern __shared__ double logit[];   // [B*K1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < B * K1; e += 8) {
    const int b = e / K1;
    const double* a = u_q + (long long)b * CD;
    const double* c = u_k + (long long)e * CD;      // view entry (b, k) = stack row b*K1 + k = e
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CD / 32; ++i) s += a[lane + 32 * i] * c[lane + 32 * i];
    s = warp_sum_d(s);
    if (lane == 0) logit[e] = s;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    double mx = -1e300;
    for (int k = 0; k < K1; ++k) mx = fmax(mx, logit[b * K1 + k]);
    double se = 0.0;
    for (int k = 0; k < K1; ++k) se += exp(logit[b * K1 + k] - mx);
    for (int k = 0; k < K1; ++k) {
      const double c = exp(logit[b * K1 + k] - mx) / se;
      comp64[b * K1 + k] = c;
      comp[b * K1 + k] = (float)c;
    }
  }
}

// grid = S + B CTAs (same roles as compat_linear_kernel): d lin = (du - u (u . du)) / norm with
//   du_q[b] = sum_k dlogit[b][k] u_k[b*K1+k],  du_k[r = b*K1+k] = dlogit[b][k] u_q[b],
//   dlogit[b][k] = comp[b][k] (dcomp[b][k] - sum_j comp[b][j] dcomp[b][j]) * gscale
__global__ void __launch_bounds__(256) compat_bwd_lin_kernel(const double* __restrict__ u_q, const double* __restrict__ u_k,
                                                            const double* __restrict__ n_q, const double* __restrict__ n_k,
                                                            const double* __restrict__ comp64, const double* __restrict__ dcomp,
                                                            const float* __restrict__ gscale, int B, int K1,
                                                            double* __restrict__ dlin_q, double* __restrict__ dlin_k) {
  __shared__ double red[8];
  const int S = B * K1, r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_q = r >= S;
  const int b = is_q ? r - S : r / K1;
  const double gs = gscale ? (double)__ldg(gscale) : 1.0;
  double dot = 0.0;
  for (int j = 0; j < K1; ++j) dot += comp64[b * K1 + j] * dcomp[b * K1 + j];
  double du;
  const double* u;
  double den;
  if (is_q) {
    du = 0.0;
    for (int k = 0; k < K1; ++k) du += comp64[b * K1 + k] * (dcomp[b * K1 + k] - dot) * gs * u_k[(long long)(b * K1 + k) * CD + tid];
    u = u_q + (long long)b * CD;
    den = n_q[b];
  } else {
    du = comp64[r] * (dcomp[r] - dot) * gs * u_q[(long long)b * CD + tid];
    u = u_k + (long long)r * CD;
    den = n_k[r];
  }
  const double ut = u[tid];
  double s = warp_sum_d(ut * du);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  double ud = 0.0;
#pragma unroll
  for (int w2 = 0; w2 < 8; ++w2) ud += red[w2];
  // (|lin| <= 1e-12 -> the clamp is active and normalize is a plain division by 1e-12: d lin = du / 1e-12)
  const double dl = den > 1e-12 ? (du - ut * ud) / den : du / den;
  (is_q ? dlin_q + (long long)b * CD : dlin_k + (long long)r * CD)[tid] = dl;
}

// grid = 256 CTAs (output channel o) x 256 threads (input channel i): dWq[o][i], dWk[o][i], db
__global__ void __launch_bounds__(256) compat_bwd_w_kernel(const float* __restrict__ pooled, const double* __restrict__ dlin_q,
                                                          const double* __restrict__ dlin_k, int B, int K1,
                                                          float* __restrict__ dWq, float* __restrict__ dbq,
                                                          float* __restrict__ dWk, float* __restrict__ dbk) {
  const int o = blockIdx.x, i = threadIdx.x, S = B * K1;
  float sq[2] = {0.f, 0.f}, sk[4] = {0.f, 0.f, 0.f, 0.f}, bq = 0.f, bk = 0.f;   // fp32: plain outer-product sums
  for (int b = 0; b < B; ++b) {
    const float d = (float)dlin_q[(long long)b * CD + o];
    sq[b & 1] = fmaf(d, __ldg(pooled + (long long)(b * K1) * CD + i), sq[b & 1]);
    bq += d;
  }
#pragma unroll 4
  for (int r = 0; r < S; ++r) {
    const float d = (float)dlin_k[(long long)r * CD + o];
    sk[r & 3] = fmaf(d, __ldg(pooled + (long long)stack_slot(r, B, K1) * CD + i), sk[r & 3]);
    bk += d;
  }
  dWq[(long long)o * CD + i] = sq[0] + sq[1];
  dWk[(long long)o * CD + i] = (sk[0] + sk[1]) + (sk[2] + sk[3]);
  if (i == 0) { dbq[o] = bq; dbk[o] = bk; }
}

// grid = S CTAs (pooled slot) x 256 threads (input channel i): dpool[slot][i] = sum_o Wk[o][i] dlin_k[r(slot)][o]
// (+ sum_o Wq[o][i] dlin_q[b][o] for the query slots k = 0); also max |dpool| for the gradient scaling
__global__ void __launch_bounds__(256) compat_bwd_pool_kernel(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                             const double* __restrict__ dlin_q, const double* __restrict__ dlin_k,
                                                             int B, int K1, float* __restrict__ dpool, float* __restrict__ dpool_amax) {
  // grid = (S slots, 4 column quarters); thread = (input channel i = 64*blockIdx.y + tid % 64, quarter tid / 64 of the
  // 256 output channels): 64 (128 for query slots) fp32 FMAs per thread in four independent chains, SMEM reduction
  __shared__ float dk[CD];
  __shared__ float dq[CD];
  __shared__ float part[4][64];
  const int slot = blockIdx.x, tid = threadIdx.x, iq = tid & 63, qt = tid >> 6, i = blockIdx.y * 64 + iq;
  const int b = slot / K1, k = slot - b * K1;
  const int r = k * B + b;                       // the stack row that holds this slot
  dk[tid] = (float)dlin_k[(long long)r * CD + tid];
  dq[tid] = k == 0 ? (float)dlin_q[(long long)b * CD + tid] : 0.f;
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int o = qt * 64; o < qt * 64 + 64; o += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fmaf(__ldg(Wk + (long long)(o + j) * CD + i), dk[o + j], acc[j]);
  }
  if (k == 0) {
#pragma unroll 4
    for (int o = qt * 64; o < qt * 64 + 64; o += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(__ldg(Wq + (long long)(o + j) * CD + i), dq[o + j], acc[j]);
    }
  }
  part[qt][iq] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (tid < 64) {
    const float s = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    dpool[(long long)slot * CD + blockIdx.y * 64 + tid] = s;
    if (dpool_amax) {
      const float m = warp_max(fabsf(s));
      if ((tid & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(dpool_amax), __float_as_int(m));
    }
  }
}

// the per-block fan-out weights and the bound of the upstream gradient that csn_ln_bwd needs, in one launch:
//   cw[j] = comp[cw_index[j]] * gscale   (0 when cw_index[j] < 0),   amax_out = amax_in * |gscale| + dpool_amax / n_points
__global__ void compat_fanout_kernel(const float* __restrict__ comp, const int* __restrict__ cw_index, int n_blocks,
                                     const float* __restrict__ gscale, const float* __restrict__ amax_in,
                                     const float* __restrict__ dpool_amax, float inv_points, float* __restrict__ cw,
                                     float* __restrict__ amax_out) {
  const float gs = gscale ? __ldg(gscale) : 1.f;
  for (int j = threadIdx.x; j < n_blocks; j += blockDim.x) {
    const int ci = __ldg(cw_index + j);
    cw[j] = ci >= 0 ? __ldg(comp + ci) * gs : 0.f;
  }
  if (threadIdx.x == 0) amax_out[0] = __ldg(amax_in) * fabsf(gs) + (dpool_amax ? __ldg(dpool_amax) * inv_points : 0.f);
}

}  // namespace csn

extern "C" int csn_compat_fanout(const float* comp, const int32_t* cw_index, int32_t n_blocks, const float* gscale,
                                 const float* amax_in, const float* dpool_amax, float inv_points, float* cw, float* amax_out,
                                 void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(comp && cw_index && amax_in && cw && amax_out, "csn_compat_fanout: null pointer");
  compat_fanout_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(comp, cw_index, n_blocks, gscale, amax_in, dpool_amax,
                                                                              inv_points, cw, amax_out);
  CSN_LAUNCH_OK("compat_fanout_kernel");
  return 0;
}

extern "C" int csn_compat_fwd(const float* pooled, const float* Wq, const float* bq, const float* Wk, const float* bk,
                              int32_t B, int32_t K1, double* u_q, double* u_k, double* n_q, double* n_k, double* comp64,
                              float* comp, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(pooled && Wq && bq && Wk && bk && u_q && u_k && n_q && n_k && comp64 && comp, "csn_compat_fwd: null pointer");
  CSN_CHECK_ARG(B >= 1 && K1 >= 1 && B * K1 <= 4096, "csn_compat_fwd: bad batch / neighbour counts (%d, %d)", B, K1);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  compat_linear_kernel<<<B * K1 + B, 256, 0, s>>>(pooled, Wq, bq, Wk, bk, B, K1, u_q, u_k, n_q, n_k);
  CSN_LAUNCH_OK("compat_linear_kernel");
  compat_softmax_kernel<<<1, 256, B * K1 * sizeof(double), s>>>(u_q, u_k, B, K1, comp64, comp);
  CSN_LAUNCH_OK("compat_softmax_kernel");
  return 0;
}

extern "C" int csn_compat_bwd(const float* pooled, const float* Wq, const float* Wk, const double* u_q, const double* u_k,
                              const double* n_q, const double* n_k, const double* comp64, const double* dcomp,
                              const float* gscale, int32_t B, int32_t K1, double* dlin_q, double* dlin_k, float* dWq,
                              float* dbq, float* dWk, float* dbk, float* dpool, float* dpool_amax, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(pooled && Wq && Wk && u_q && u_k && n_q && n_k && comp64 && dcomp && dlin_q && dlin_k && dWq && dbq && dWk &&
                dbk && dpool, "csn_compat_bwd: null pointer");
  CSN_CHECK_ARG(B >= 1 && K1 >= 1, "csn_compat_bwd: bad batch / neighbour counts");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  compat_bwd_lin_kernel<<<B * K1 + B, 256, 0, s>>>(u_q, u_k, n_q, n_k, comp64, dcomp, gscale, B, K1, dlin_q, dlin_k);
  CSN_LAUNCH_OK("compat_bwd_lin_kernel");
  compat_bwd_w_kernel<<<CD, 256, 0, s>>>(pooled, dlin_q, dlin_k, B, K1, dWq, dbq, dWk, dbk);
  CSN_LAUNCH_OK("compat_bwd_w_kernel");
  compat_bwd_pool_kernel<<<dim3(B * K1, 4), 256, 0, s>>>(Wq, Wk, dlin_q, dlin_k, B, K1, dpool, dpool_amax);
  CSN_LAUNCH_OK("compat_bwd_pool_kernel");
  return 0;
}
