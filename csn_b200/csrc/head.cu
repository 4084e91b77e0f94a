// Fused CSA head: everything between the attention blocks' pre-LayerNorm rows and the loss,
// forward AND backward, in one pass over Z:
//   y[b][r]      = sum_k comp[b,k] * LayerNorm(Z[blk(b,k)][r])          (csa_models.py:232-238, the weighted sum)
//   logits       = W y                                                   (the bias-free 1x1 `logit` conv, :201)
//   loss, accu   = masked cross-entropy / accuracy                       (csa_training.py:94-108: label > 0 only)
//   intsc, union = per-class IoU counters                                (csa_training.py:110-134)
//   dlogits      = (softmax - onehot) / n_valid ;  dW = dlogits^T y ;  dy = W^T dlogits
//   dOutT[b][r]  = dy, row-major padded rows: exactly what csn_ln_bwd consumes (no transposes anywhere)
//   dcomp[b,k]   = sum_r < dy[b][r], LayerNorm(Z[blk(b,k)][r]) >          (gradient of the compatibility softmax)
// It replaces csn_combine_fwd + the ATen conv / log-softmax / nll kernels and their backwards + csn_pack_rows(dOut)
// + csn_block_dot: the (B,256,N) output activation, the logits and the output gradient never reach HBM in
// channel-major form, Z is read once from HBM (the second read for dcomp hits L2: same CTA, same tile).
//
// CTA = 256 threads, persistent over tiles of 32 padded rows.  Phases per tile:
//   P1 (warp per row)  : y rows -> SMEM tile
//   P2 (8 threads/row) : logits, softmax, loss, argmax, dlogits -> SMEM ; IoU counters
//   P3a (thread = channel): dW += dlogits^T y (register accumulators, flushed once per CTA)
//   P3b (8 threads/row): dy = W^T dlogits -> the SMEM tile (y is dead)
//   P4 (warp per row)  : dOutT rows, max|dy|, dcomp dots against the re-normalised Z rows
#include <stdint.h>

#include "host_util.h"
#include "ptx.cuh"

namespace csn {

constexpr int HD_DM = 256;
constexpr int HD_ROWS = 32;       // rows per tile
constexpr int HD_LD = 260;        // SMEM row stride in floats (1040 B: 16-byte aligned, bank offset 4 per row)
constexpr int HD_MAXK = 6;        // K + 1 <= 6 (reference: K <= 5)

struct HeadArgs {
  const float* Z; const float* mean; const float* rstd; const float* gamma; const float* beta;
  const int* blk;            // [n_b][n_k]: attention block of (b, k)
  const float* w;            // [n_b][n_k]: compatibility weights
  int n_k;
  const float* W; int C;     // logit weights [C][256]
  const long long* labels; long long lab_stride; int ignore_index;
  const int* n_valid;        // device scalar: number of unmasked points (csn_count_valid)
  int n_points, chunk, chunk_pad, rows_pad, n_tiles, tiles_per_b;
  float* loss_part;          // [n_tiles]: sum of -log p[label] over the tile's valid points
  float* dOutT;              // [n_b*rows_pad][256] or null (forward only)
  float* amax;               // optional: max |dOutT| (atomic max on the bit pattern; zero-initialised by the caller)
  float* dcomp;              // [n_b*n_k], atomically accumulated (zero-initialised by the caller) or null
  float* dW_part;            // [gridDim.x][C][256] per-CTA partial sums of dlogits^T y
  int* stats;                // [3*C + 2]: #pred==c, #label==c, #both, then #correct, #labels out of range
  float* y_out;              // optional [n_b*rows_pad][256]: the combined features, row-major padded rows
};

template <int CMAX, int NK>
__global__ void __launch_bounds__(256, 2) csa_head_kernel(const HeadArgs p) {
  extern __shared__ __align__(16) float hsm[];
  float* ys = hsm;                                  // [32][260]
  float* Wsm = ys + HD_ROWS * HD_LD;                // [CMAX][260]
  float* dls = Wsm + CMAX * HD_LD;                  // [32][CMAX]
  float* red = dls + HD_ROWS * CMAX;                // [8][8]
  float* gsm = red + 64;                            // gamma [256] | beta [256]
  int* hist = reinterpret_cast<int*>(gsm + 512);    // [3*CMAX + 2]
  constexpr int CPT = CMAX / 8;                     // classes per thread in P2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool grad = p.dOutT != nullptr;

  for (int i = tid; i < CMAX * HD_LD; i += 256) {
    const int c = i / HD_LD, k = i - c * HD_LD;
    Wsm[i] = (c < p.C && k < HD_DM) ? __ldg(p.W + c * HD_DM + k) : 0.f;
  }
  for (int i = tid; i < 3 * CMAX + 2; i += 256) hist[i] = 0;
  gsm[tid] = __ldg(p.gamma + tid);
  gsm[256 + tid] = __ldg(p.beta + tid);
  const float4* gs4 = reinterpret_cast<const float4*>(gsm);
  const float inv_nv = 1.f / (float)max(__ldg(p.n_valid), 1);
  float dw[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) dw[c] = 0.f;
  float amx = 0.f;
  __syncthreads();

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_b;
    const int r0 = (tile - b * p.tiles_per_b) * HD_ROWS;
    // per-(b,k) block index and weight (n_k <= 6 values each, L1-resident)
    int kb[NK]; float kw[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      kb[k] = (k < p.n_k) ? __ldg(p.blk + b * p.n_k + k) : 0;
      kw[k] = (k < p.n_k) ? __ldg(p.w + b * p.n_k + k) : 0.f;
    }
    float wsum = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) wsum += kw[k];

    // ------------------------------------------------------------------ P1: y rows
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const int rr = warp * 4 + i, r = r0 + rr;
      const int ii = r % p.chunk_pad;
      const bool rvalid = ii < p.chunk && (r / p.chunk_pad) * p.chunk + ii < p.n_points;
      float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tc = ta;
      if (rvalid) {
        float4 za[NK], zc[NK]; float mu[NK], rs[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (k < p.n_k) {
            const long long row = (long long)kb[k] * p.rows_pad + r;
            const float4* z4 = reinterpret_cast<const float4*>(p.Z + row * HD_DM);
            za[k] = __ldg(z4 + lane); zc[k] = __ldg(z4 + 32 + lane);
            mu[k] = __ldg(p.mean + row); rs[k] = __ldg(p.rstd + row);
          }
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (k < p.n_k) {
            const float a = kw[k] * rs[k], m = mu[k];
            ta.x += a * (za[k].x - m); ta.y += a * (za[k].y - m); ta.z += a * (za[k].z - m); ta.w += a * (za[k].w - m);
            tc.x += a * (zc[k].x - m); tc.y += a * (zc[k].y - m); tc.z += a * (zc[k].z - m); tc.w += a * (zc[k].w - m);
          }
        const float4 g0 = gs4[lane], g1 = gs4[32 + lane], b0 = gs4[64 + lane], b1 = gs4[96 + lane];
        ta.x = ta.x * g0.x + wsum * b0.x; ta.y = ta.y * g0.y + wsum * b0.y; ta.z = ta.z * g0.z + wsum * b0.z; ta.w = ta.w * g0.w + wsum * b0.w;
        tc.x = tc.x * g1.x + wsum * b1.x; tc.y = tc.y * g1.y + wsum * b1.y; tc.z = tc.z * g1.z + wsum * b1.z; tc.w = tc.w * g1.w + wsum * b1.w;
      }
      *reinterpret_cast<float4*>(ys + rr * HD_LD + lane * 4) = ta;
      *reinterpret_cast<float4*>(ys + rr * HD_LD + 128 + lane * 4) = tc;
      if (p.y_out) {
        float4* o4 = reinterpret_cast<float4*>(p.y_out + ((long long)b * p.rows_pad + r) * HD_DM);
        o4[lane] = ta; o4[32 + lane] = tc;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ P2: logits -> softmax -> loss / dlogits
    {
      const int rr = tid >> 3, cg = tid & 7, r = r0 + rr;
      const int ii = r % p.chunk_pad;
      const int n = (r / p.chunk_pad) * p.chunk + ii;
      const bool rvalid = ii < p.chunk && n < p.n_points;
      float acc[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
      const float4* y4 = reinterpret_cast<const float4*>(ys + rr * HD_LD);
#pragma unroll 4
      for (int k4 = 0; k4 < 64; ++k4) {
        const float4 yv = y4[k4];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + (cg + 8 * i) * HD_LD + k4 * 4);
          acc[i] += (yv.x * wv.x + yv.y * wv.y) + (yv.z * wv.z + yv.w * wv.w);
        }
      }
      long long lab = rvalid ? __ldg(p.labels + (long long)b * p.lab_stride + n) : (long long)p.ignore_index;
      const bool in_range = lab >= 0 && lab < p.C;
      const bool pvalid = rvalid && lab != p.ignore_index && in_range;
      float mx = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        const int c = cg + 8 * i;
        if (c < p.C && acc[i] > mx) { mx = acc[i]; bi = c; }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {   // the 8 threads of a row are consecutive lanes
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (om > mx || (om == mx && oi < bi)) { mx = om; bi = oi; }
      }
      float e[CPT], se = 0.f;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        e[i] = (cg + 8 * i < p.C) ? __expf(acc[i] - mx) : 0.f;
        se += e[i];
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      const float inv = 1.f / se;
      float lterm = 0.f;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        const int c = cg + 8 * i;
        const bool hit = pvalid && c == (int)lab;
        if (hit) lterm = -(acc[i] - mx - __logf(se));
        dls[rr * CMAX + c] = pvalid ? (e[i] * inv - (hit ? 1.f : 0.f)) * inv_nv : 0.f;
      }
      if (cg == 0) {
        if (pvalid) {
          atomicAdd(hist + bi, 1);
          atomicAdd(hist + CMAX + (int)lab, 1);
          if (bi == (int)lab) { atomicAdd(hist + 2 * CMAX + bi, 1); atomicAdd(hist + 3 * CMAX, 1); }
        } else if (rvalid && !in_range) {
          atomicAdd(hist + 3 * CMAX + 1, 1);
        }
      }
      lterm = warp_sum(lterm);
      if (lane == 0) red[warp] = lterm;
    }
    __syncthreads();
    if (tid == 0) p.loss_part[tile] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    if (!grad) { __syncthreads(); continue; }

    // ------------------------------------------------------------------ P3a: dW[c][tid] += sum_r dl[r][c] * y[r][tid]
#pragma unroll 4
    for (int rr = 0; rr < HD_ROWS; ++rr) {
      const float yv = ys[rr * HD_LD + tid];
#pragma unroll
      for (int q = 0; q < CMAX / 4; ++q) {
        const float4 d4 = *reinterpret_cast<const float4*>(dls + rr * CMAX + q * 4);
        dw[4 * q] += d4.x * yv; dw[4 * q + 1] += d4.y * yv; dw[4 * q + 2] += d4.z * yv; dw[4 * q + 3] += d4.w * yv;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ P3b: dy[r][ch] = sum_c dl[r][c] W[c][ch]  (into ys)
    {
      const int rr = tid >> 3, j = tid & 7;
      float4 a4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int c = 0; c < CMAX; ++c) {
        if (c >= p.C) break;
        const float d = dls[rr * CMAX + c];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + c * HD_LD + 32 * i + 4 * j);
          a4[i].x += d * wv.x; a4[i].y += d * wv.y; a4[i].z += d * wv.z; a4[i].w += d * wv.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(ys + rr * HD_LD + 32 * i + 4 * j) = a4[i];
    }
    __syncthreads();

    // ------------------------------------------------------------------ P4: dOutT rows, amax, dcomp
    {
      float dc[NK];
#pragma unroll
      for (int k = 0; k < NK; ++k) dc[k] = 0.f;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int rr = warp * 4 + i, r = r0 + rr;
        const int ii = r % p.chunk_pad;
        const bool rvalid = ii < p.chunk && (r / p.chunk_pad) * p.chunk + ii < p.n_points;
        const float4 da = *reinterpret_cast<const float4*>(ys + rr * HD_LD + lane * 4);
        const float4 dc4 = *reinterpret_cast<const float4*>(ys + rr * HD_LD + 128 + lane * 4);
        float4* o4 = reinterpret_cast<float4*>(p.dOutT + ((long long)b * p.rows_pad + r) * HD_DM);
        o4[lane] = da; o4[32 + lane] = dc4;   // zero rows for pads and masked points
        amx = fmaxf(amx, fmaxf(fmaxf(fabsf(da.x), fabsf(da.y)), fmaxf(fabsf(da.z), fabsf(da.w))));
        amx = fmaxf(amx, fmaxf(fmaxf(fabsf(dc4.x), fabsf(dc4.y)), fmaxf(fabsf(dc4.z), fabsf(dc4.w))));
        if (!rvalid || p.dcomp == nullptr) continue;
        const float4 g0 = gs4[lane], g1 = gs4[32 + lane], b0 = gs4[64 + lane], b1 = gs4[96 + lane];
        // dg = dy o gamma, s_b = <dy, beta>: <dy, LN(z_k)> = rs_k * (<dg, z_k> - mu_k * sum(dg)) + s_b
        const float4 ga = make_float4(da.x * g0.x, da.y * g0.y, da.z * g0.z, da.w * g0.w);
        const float4 gc = make_float4(dc4.x * g1.x, dc4.y * g1.y, dc4.z * g1.z, dc4.w * g1.w);
        const float sb = (da.x * b0.x + da.y * b0.y) + (da.z * b0.z + da.w * b0.w) + (dc4.x * b1.x + dc4.y * b1.y) + (dc4.z * b1.z + dc4.w * b1.w);
        float4 za[NK], zc[NK]; float mu[NK], rs[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (k < p.n_k) {
            const long long row = (long long)kb[k] * p.rows_pad + r;
            const float4* z4 = reinterpret_cast<const float4*>(p.Z + row * HD_DM);
            za[k] = __ldg(z4 + lane); zc[k] = __ldg(z4 + 32 + lane);
            mu[k] = __ldg(p.mean + row); rs[k] = __ldg(p.rstd + row);
          }
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (k < p.n_k) {
            const float m = mu[k];
            const float s = (ga.x * (za[k].x - m) + ga.y * (za[k].y - m)) + (ga.z * (za[k].z - m) + ga.w * (za[k].w - m)) +
                            (gc.x * (zc[k].x - m) + gc.y * (zc[k].y - m)) + (gc.z * (zc[k].z - m) + gc.w * (zc[k].w - m));
            dc[k] += rs[k] * s + sb;
          }
      }
      if (p.dcomp) {
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          const float s = warp_sum(dc[k]);
          if (lane == 0) red[warp * 8 + k] = s;
        }
      }
    }
    __syncthreads();
    if (p.dcomp && tid < p.n_k) {
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) s += red[w2 * 8 + tid];
      atomicAdd(p.dcomp + b * p.n_k + tid, s);
    }
    // (the next tile's P1 writes ys / red only after its own barrier sequence; P4's reads are complete here)
    __syncthreads();
  }

  // ---------------------------------------------------------------------- flush
  if (grad) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < p.C) p.dW_part[((long long)blockIdx.x * p.C + c) * HD_DM + tid] = dw[c];
    if (p.amax) {
      amx = warp_max(amx);
      if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(p.amax), __float_as_int(amx));
    }
  }
  __syncthreads();
  for (int i = tid; i < 3 * CMAX + 2; i += 256) {
    const int v = hist[i];
    if (v == 0) continue;
    if (i >= 3 * CMAX) atomicAdd(p.stats + 3 * p.C + (i - 3 * CMAX), v);
    else {
      const int which = i / CMAX, c = i - which * CMAX;
      if (c < p.C) atomicAdd(p.stats + which * p.C + c, v);
    }
  }
}

// dW[c][ch] = scale * sum_i part[i][c][ch]   (fixed order: deterministic)
__global__ void head_dw_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int n_part, int C) {
  const int c = blockIdx.x, ch = threadIdx.x;
  float s = 0.f;
  for (int i = 0; i < n_part; ++i) s += part[((long long)i * C + c) * HD_DM + ch];
  out[c * HD_DM + ch] = s;
}

__global__ void head_count_valid_kernel(const long long* __restrict__ labels, long long lab_stride, int n_b, int n_points,
                                        int ignore_index, int* __restrict__ count) {
  int c = 0;
  const long long total = (long long)n_b * n_points;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n_points, n = i - b * n_points;
    c += labels[b * lab_stride + n] != ignore_index;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

template <int CMAX, int NK>
static int launch_head_nk(const HeadArgs& a, int grid, cudaStream_t s) {
  auto kern = csa_head_kernel<CMAX, NK>;
  const int smem = (HD_ROWS * HD_LD + CMAX * HD_LD + HD_ROWS * CMAX + 64 + 512 + 3 * CMAX + 2) * 4;
  int dev = 0;
  CSN_CUDA_OK(cudaGetDevice(&dev));
  static bool configured[64] = {false};
  if (dev < 64 && !configured[dev]) {
    CSN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = true;
  }
  kern<<<grid, 256, smem, s>>>(a);
  CSN_LAUNCH_OK("csa_head_kernel");
  return 0;
}

template <int CMAX>
static int launch_head(const HeadArgs& a, int grid, cudaStream_t s) {
  if (a.n_k <= 1) return launch_head_nk<CMAX, 1>(a, grid, s);
  if (a.n_k <= 2) return launch_head_nk<CMAX, 2>(a, grid, s);
  if (a.n_k <= 4) return launch_head_nk<CMAX, 4>(a, grid, s);
  return launch_head_nk<CMAX, 6>(a, grid, s);
}

}  // namespace csn

extern "C" int csn_csa_head_grid(int32_t n_b, int32_t rows_pad) {
  const int n_tiles = n_b * (rows_pad / csn::HD_ROWS);
  const int cap = csn::num_sms() * 2;
  return n_tiles < cap ? n_tiles : cap;
}

extern "C" int csn_csa_head(const float* Z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            const int32_t* blk, const float* w, int32_t n_b, int32_t n_k, const float* W, int32_t n_classes,
                            const int64_t* labels, int64_t lab_stride, int32_t ignore_index, int32_t* n_valid,
                            int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, float* loss_part,
                            float* dOutT, float* amax, float* dcomp, float* dW_part, float* dW, int32_t* stats, float* y_out,
                            void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Z && mean && rstd && gamma && beta && blk && w && W && labels && n_valid && loss_part && stats,
                "csn_csa_head: null pointer");
  CSN_CHECK_ARG(n_k >= 1 && n_k <= HD_MAXK, "csn_csa_head: 1..%d attention blocks per query (got %d)", HD_MAXK, n_k);
  CSN_CHECK_ARG(n_classes >= 1 && n_classes <= 64, "csn_csa_head: 1..64 classes supported (got %d)", n_classes);
  CSN_CHECK_ARG(chunk_pad % HD_ROWS == 0 && rows_pad % chunk_pad == 0 && chunk <= chunk_pad, "csn_csa_head: bad padding");
  CSN_CHECK_ARG(!dOutT || (dW_part && dW), "csn_csa_head: the backward outputs need dW_part and dW");
  if (n_b == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  head_count_valid_kernel<<<148, 256, 0, s>>>(reinterpret_cast<const long long*>(labels), lab_stride, n_b, n_points, ignore_index, n_valid);
  CSN_LAUNCH_OK("head_count_valid_kernel");
  const int grid = csn_csa_head_grid(n_b, rows_pad);
  HeadArgs a{Z, mean, rstd, gamma, beta, blk, w, n_k, W, n_classes, reinterpret_cast<const long long*>(labels), lab_stride,
             ignore_index, n_valid, n_points, chunk, chunk_pad, rows_pad, n_b * (rows_pad / HD_ROWS), rows_pad / HD_ROWS,
             loss_part, dOutT, amax, dcomp, dW_part, stats, y_out};
  int rc;
  if (n_classes <= 16) rc = launch_head<16>(a, grid, s);
  else if (n_classes <= 32) rc = launch_head<32>(a, grid, s);
  else rc = launch_head<64>(a, grid, s);
  if (rc) return rc;
  if (dOutT) {
    head_dw_reduce_kernel<<<n_classes, HD_DM, 0, s>>>(dW_part, dW, grid, n_classes);
    CSN_LAUNCH_OK("head_dw_reduce_kernel");
  }
  return 0;
}
