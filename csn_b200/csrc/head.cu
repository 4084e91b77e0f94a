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
// CTA = 256 threads, ONE per SM, persistent over a contiguous range of tiles of R (16 or 8) padded rows.  The Z rows
// of a tile (R contiguous KB per attention block) and their statistics arrive by 1-D bulk copies (cp.async.bulk +
// mbarrier) into a 2-stage SMEM ring: the whole next tile is in flight while this one is processed, and the second
// use of the rows (dcomp) reads SMEM, not L2.  Phases per tile:
//   P1  (warp per row)       : y rows -> SMEM tile ys
//   P2a (warp = 32-channel slice of the contraction, thread = 2 rows x 4 classes): partial logits -> SMEM
//   P2b (warp per row, lane = class): logits, softmax, loss, argmax, dlogits -> SMEM ; IoU counters
//   P3a (thread = channel)   : dW += dlogits^T y (register accumulators, flushed once per CTA)
//   P3b (warp = 32 channels, thread = 4 rows x 4 channels): dy = W^T dlogits -> ys (y is dead)
//   P4  (warp per row)       : dOutT rows, max|dy|, dcomp dots against the re-normalised Z rows (from SMEM)
#include <stdint.h>

#include "host_util.h"
#include "ptx.cuh"

namespace csn {

constexpr int HD_DM = 256;
constexpr int HD_LD = 260;        // SMEM row stride in floats (1040 B: 16-byte aligned, bank offset 4 per row)
constexpr int HD_MAXK = 6;        // K + 1 <= 6 (reference: K <= 5)

struct HeadArgs {
  const float* Z; const float* mean; const float* rstd; const float* gamma; const float* beta;
  const int* blk;            // [n_b][n_k]: attention block of (b, k)
  const float* w;            // [n_b][n_k]: compatibility weights
  int n_k;
  const float* W; int C;     // logit weights [C][256]
  const long long* labels; long long lab_stride; int ignore_index;
  const int* n_valid;        // device scalar: number of unmasked points
  int n_points, chunk, chunk_pad, rows_pad, n_tiles, tiles_per_b;
  float* loss_part;          // [gridDim.x]: sum of -log p[label] over the CTA's valid points
  float* dOutT;              // [n_b*rows_pad][256] or null (forward only)
  float* amax;               // optional: max |dOutT| (atomic max on the bit pattern; zero-initialised by the caller)
  double* dcomp;             // [n_b*n_k] fp64, atomically accumulated (zero-initialised by the caller) or null
  float* dW_part;            // [gridDim.x][C][256] per-CTA partial sums of dlogits^T y
  int* stats;                // [3*C + 2]: #pred==c, #label==c, #both, then #correct, #labels out of range
  float* y_out;              // optional [n_b*rows_pad][256]: the combined features, row-major padded rows
};

template <int CMAX, int NK, int R>
struct HeadSmem {
  static constexpr int DLS = CMAX + 4;                       // dlogits row stride (floats)
  static constexpr int STAGE = NK * (R * HD_DM + 2 * R);     // floats per stage: Z rows, then mean, then rstd
  static constexpr int OFF_YS = 2 * STAGE;
  static constexpr int OFF_W = OFF_YS + R * HD_LD;
  static constexpr int OFF_PART = OFF_W + CMAX * HD_LD;
  static constexpr int OFF_DLS = OFF_PART + 8 * R * CMAX;
  static constexpr int OFF_RED = OFF_DLS + R * DLS;
  static constexpr int OFF_G = OFF_RED + 64;
  static constexpr int OFF_HIST = OFF_G + 512;
  static constexpr int OFF_BAR = (OFF_HIST + 3 * CMAX + 2 + 3) / 4 * 4;
  static constexpr int OFF_RED64 = OFF_BAR + 4;              // [8][8] doubles
  static constexpr int BYTES = (OFF_RED64 + 128) * 4;
};

template <int CMAX, int NK, int R>
__global__ void __launch_bounds__(256, (R == 8 && CMAX == 16 && NK <= 4) ? 2 : 1) csa_head_kernel(const HeadArgs p) {
  using SM = HeadSmem<CMAX, NK, R>;
  extern __shared__ __align__(128) float hsm[];
  float* ys = hsm + SM::OFF_YS;                       // [R][260]
  float* Wsm = hsm + SM::OFF_W;                       // [CMAX][260]
  float* part = hsm + SM::OFF_PART;                   // [8][R*CMAX]
  float* dls = hsm + SM::OFF_DLS;                     // [R][DLS]
  float* red = hsm + SM::OFF_RED;                     // [8][8]
  float* gsm = hsm + SM::OFF_G;                       // gamma [256] | beta [256]
  int* hist = reinterpret_cast<int*>(hsm + SM::OFF_HIST);
  const uint32_t bar0 = smem_u32(hsm + SM::OFF_BAR);
  double* red64 = reinterpret_cast<double*>(hsm + SM::OFF_RED64);
  constexpr int RPT = R / 8;                          // rows per thread in P2a / rows per warp in P1, P2b, P4
  constexpr int CQ = CMAX / 4;                        // classes per thread in P2a
  constexpr int CL = (CMAX + 31) / 32;                // classes per lane in P2b
  constexpr int RQ = R / 4;                           // rows per thread in P3b
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool grad = p.dOutT != nullptr;

  for (int i = tid; i < CMAX * HD_LD; i += 256) {
    const int c = i / HD_LD, k = i - c * HD_LD;
    Wsm[i] = (c < p.C && k < HD_DM) ? __ldg(p.W + c * HD_DM + k) : 0.f;
  }
  for (int i = tid; i < 3 * CMAX + 2; i += 256) hist[i] = 0;
  gsm[tid] = __ldg(p.gamma + tid);
  gsm[256 + tid] = __ldg(p.beta + tid);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  const float4* gs4 = reinterpret_cast<const float4*>(gsm);
  const float inv_nv = 1.f / (float)max(__ldg(p.n_valid), 1);
  float dw[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) dw[c] = 0.f;
  float amx = 0.f, lacc = 0.f;
  // d comp is consumed through differences of nearly equal numbers (softmax backward over the K+1 attention
  // outputs of one query): its partial sums are carried in fp64 from the per-row dot products onwards
  double dc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) dc[k] = 0.0;
  __syncthreads();

  const int t_begin = (int)((long long)p.n_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)p.n_tiles * (blockIdx.x + 1) / gridDim.x);
  // one thread feeds the ring: R KB of Z rows + R means + R rstds per attention block of the tile's query
  auto issue = [&](int tile, int stage) {
    const int b = tile / p.tiles_per_b;
    const int r0 = (tile - b * p.tiles_per_b) * R;
    float* st = hsm + stage * SM::STAGE;
    const uint32_t bar = bar0 + 8u * stage;
    mbar_arrive_expect_tx(bar, (uint32_t)p.n_k * (R * HD_DM + 2 * R) * 4u);
    for (int k = 0; k < p.n_k; ++k) {
      const long long row = (long long)__ldg(p.blk + b * p.n_k + k) * p.rows_pad + r0;
      bulk_load_1d(smem_u32(st + k * R * HD_DM), p.Z + row * HD_DM, R * HD_DM * 4, bar);
      bulk_load_1d(smem_u32(st + NK * R * HD_DM + k * R), p.mean + row, R * 4, bar);
      bulk_load_1d(smem_u32(st + NK * R * HD_DM + NK * R + k * R), p.rstd + row, R * 4, bar);
    }
  };
  // dcomp partial sums of the current query are kept in registers across its tiles
  auto flush_dcomp = [&](int b) {   // called by every thread (CTA-uniform)
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      double s = dc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red64[warp * 8 + k] = s;
      dc[k] = 0.0;
    }
    __syncthreads();
    if (tid < p.n_k) {
      double s = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < 8; ++w2) s += red64[w2 * 8 + tid];
      atomicAdd(p.dcomp + b * p.n_k + tid, s);
    }
    __syncthreads();
  };
  if (tid == 0 && t_begin < t_end) issue(t_begin, 0);
  int cur_b = -1;

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int it = tile - t_begin, stage = it & 1;
    const int b = tile / p.tiles_per_b;
    const int r0 = (tile - b * p.tiles_per_b) * R;
    if (b != cur_b) {
      if (cur_b >= 0 && grad && p.dcomp) flush_dcomp(cur_b);
      cur_b = b;
    }
    if (tid == 0 && tile + 1 < t_end) issue(tile + 1, stage ^ 1);   // that stage was released by the barrier ending tile-1
    float kw[NK];
    float wsum = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      kw[k] = (k < p.n_k) ? __ldg(p.w + b * p.n_k + k) : 0.f;
      wsum += kw[k];
    }
    // a tile lies inside one chunk (R | chunk_pad): one division per tile, additions per row
    const int t_ch = r0 / p.chunk_pad, t_i0 = r0 - t_ch * p.chunk_pad, t_n0 = t_ch * p.chunk + t_i0;
    const float* zs = hsm + stage * SM::STAGE;
    const float* ms = zs + NK * R * HD_DM;
    const float* rss = ms + NK * R;
    mbar_wait(bar0 + 8u * stage, (uint32_t)(it >> 1) & 1u);

    // ------------------------------------------------------------------ P1: y rows
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int rr = warp + 8 * i, r = r0 + rr;
      const bool rvalid = t_i0 + rr < p.chunk && t_n0 + rr < p.n_points;
      float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tc = ta;
      if (rvalid) {
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (k < p.n_k) {
            const float4* z4 = reinterpret_cast<const float4*>(zs + (k * R + rr) * HD_DM);
            const float4 za = z4[lane], zc = z4[32 + lane];
            const float m = ms[k * R + rr], a = kw[k] * rss[k * R + rr];
            ta.x += a * (za.x - m); ta.y += a * (za.y - m); ta.z += a * (za.z - m); ta.w += a * (za.w - m);
            tc.x += a * (zc.x - m); tc.y += a * (zc.y - m); tc.z += a * (zc.z - m); tc.w += a * (zc.w - m);
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

    // ------------------------------------------------------------------ P2a: partial logits over this warp's 32 channels
    {
      const int rg = lane & 7, cq = lane >> 3;
      float acc[RPT][CQ];
#pragma unroll
      for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < CQ; ++j) acc[i][j] = 0.f;
#pragma unroll 2
      for (int k4 = warp * 8; k4 < warp * 8 + 8; ++k4) {
        float4 yv[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) yv[i] = *reinterpret_cast<const float4*>(ys + (rg + 8 * i) * HD_LD + k4 * 4);
#pragma unroll
        for (int j = 0; j < CQ; ++j) {
          const float4 wv = *reinterpret_cast<const float4*>(Wsm + (cq + 4 * j) * HD_LD + k4 * 4);
#pragma unroll
          for (int i = 0; i < RPT; ++i) acc[i][j] += (yv[i].x * wv.x + yv[i].y * wv.y) + (yv[i].z * wv.z + yv[i].w * wv.w);
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < CQ; ++j) part[warp * (R * CMAX) + (rg + 8 * i) * CMAX + cq + 4 * j] = acc[i][j];
    }
    __syncthreads();

    // ------------------------------------------------------------------ P2b: softmax, loss, argmax, dlogits, counters
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int rr = warp + 8 * i;
      const int n = t_n0 + rr;
      const bool rvalid = t_i0 + rr < p.chunk && n < p.n_points;
      float lg[CL];
      float mx = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < CL; ++j) {
        const int c = lane + 32 * j;
        float s = 0.f;
        if (c < CMAX) {
#pragma unroll
          for (int w2 = 0; w2 < 8; ++w2) s += part[w2 * (R * CMAX) + rr * CMAX + c];
        }
        lg[j] = s;
        if (c < p.C && s > mx) { mx = s; bi = c; }
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (om > mx || (om == mx && oi < bi)) { mx = om; bi = oi; }
      }
      float e[CL], se = 0.f;
#pragma unroll
      for (int j = 0; j < CL; ++j) {
        e[j] = (lane + 32 * j < p.C) ? __expf(lg[j] - mx) : 0.f;
        se += e[j];
      }
      se = warp_sum(se);
      const long long lab = rvalid ? __ldg(p.labels + (long long)b * p.lab_stride + n) : (long long)p.ignore_index;
      const bool in_range = lab >= 0 && lab < p.C;
      const bool pvalid = rvalid && lab != p.ignore_index && in_range;
      const float inv = 1.f / se;
#pragma unroll
      for (int j = 0; j < CL; ++j) {
        const int c = lane + 32 * j;
        const bool hit = pvalid && c == (int)lab;
        if (hit) lacc += -(lg[j] - mx - __logf(se));
        if (c < CMAX) dls[rr * SM::DLS + c] = pvalid ? (e[j] * inv - (hit ? 1.f : 0.f)) * inv_nv : 0.f;
      }
      if (lane == 0) {
        if (pvalid) {
          atomicAdd(hist + bi, 1);
          atomicAdd(hist + CMAX + (int)lab, 1);
          if (bi == (int)lab) { atomicAdd(hist + 2 * CMAX + bi, 1); atomicAdd(hist + 3 * CMAX, 1); }
        } else if (rvalid && !in_range) {
          atomicAdd(hist + 3 * CMAX + 1, 1);
        }
      }
    }
    __syncthreads();
    if (!grad) continue;   // (the next tile's first barrier orders its ys writes after this tile's reads)

    // ------------------------------------------------------------------ P3a: dW[c][tid] += sum_r dl[r][c] * y[r][tid]
#pragma unroll 4
    for (int rr = 0; rr < R; ++rr) {
      const float yv = ys[rr * HD_LD + tid];
#pragma unroll
      for (int q = 0; q < CMAX / 4; ++q) {
        const float4 d4 = *reinterpret_cast<const float4*>(dls + rr * SM::DLS + q * 4);
        dw[4 * q] += d4.x * yv; dw[4 * q + 1] += d4.y * yv; dw[4 * q + 2] += d4.z * yv; dw[4 * q + 3] += d4.w * yv;
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ P3b: dy[r][ch] = sum_c dl[r][c] W[c][ch]  (into ys)
    {
      const int rg = lane >> 3, ch = warp * 32 + (lane & 7) * 4;
      float4 a4[RQ];
#pragma unroll
      for (int i = 0; i < RQ; ++i) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 3
      for (int c = 0; c < p.C; ++c) {
        const float4 wv = *reinterpret_cast<const float4*>(Wsm + c * HD_LD + ch);
#pragma unroll
        for (int i = 0; i < RQ; ++i) {
          const float d = dls[(rg + 4 * i) * SM::DLS + c];
          a4[i].x += d * wv.x; a4[i].y += d * wv.y; a4[i].z += d * wv.z; a4[i].w += d * wv.w;
        }
      }
#pragma unroll
      for (int i = 0; i < RQ; ++i) *reinterpret_cast<float4*>(ys + (rg + 4 * i) * HD_LD + ch) = a4[i];
    }
    __syncthreads();

    // ------------------------------------------------------------------ P4: dOutT rows, amax, dcomp
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int rr = warp + 8 * i, r = r0 + rr;
      const bool rvalid = t_i0 + rr < p.chunk && t_n0 + rr < p.n_points;
      const float4 da = *reinterpret_cast<const float4*>(ys + rr * HD_LD + lane * 4);
      const float4 dd = *reinterpret_cast<const float4*>(ys + rr * HD_LD + 128 + lane * 4);
      float4* o4 = reinterpret_cast<float4*>(p.dOutT + ((long long)b * p.rows_pad + r) * HD_DM);
      o4[lane] = da; o4[32 + lane] = dd;   // zero rows for pads and masked points
      amx = fmaxf(amx, fmaxf(fmaxf(fabsf(da.x), fabsf(da.y)), fmaxf(fabsf(da.z), fabsf(da.w))));
      amx = fmaxf(amx, fmaxf(fmaxf(fabsf(dd.x), fabsf(dd.y)), fmaxf(fabsf(dd.z), fabsf(dd.w))));
      if (!rvalid || p.dcomp == nullptr) continue;
      // dg = dy o gamma, s_b = <dy, beta>: <dy, LN(z_k)> = rs_k * <dg, z_k - mu_k> + s_b
      const float4 g0 = gs4[lane], g1 = gs4[32 + lane], b0 = gs4[64 + lane], b1 = gs4[96 + lane];
      const float4 ga = make_float4(da.x * g0.x, da.y * g0.y, da.z * g0.z, da.w * g0.w);
      const float4 gc = make_float4(dd.x * g1.x, dd.y * g1.y, dd.z * g1.z, dd.w * g1.w);
      const float sb = (da.x * b0.x + da.y * b0.y) + (da.z * b0.z + da.w * b0.w) + (dd.x * b1.x + dd.y * b1.y) + (dd.z * b1.z + dd.w * b1.w);
#pragma unroll
      for (int k = 0; k < NK; ++k)
        if (k < p.n_k) {
          const float4* z4 = reinterpret_cast<const float4*>(zs + (k * R + rr) * HD_DM);
          const float4 za = z4[lane], zc = z4[32 + lane];
          const float m = ms[k * R + rr];
          const float s = (ga.x * (za.x - m) + ga.y * (za.y - m)) + (ga.z * (za.z - m) + ga.w * (za.w - m)) +
                          (gc.x * (zc.x - m) + gc.y * (zc.y - m)) + (gc.z * (zc.z - m) + gc.w * (zc.w - m));
          dc[k] += (double)(rss[k * R + rr] * s) + (double)sb;
        }
    }
    __syncthreads();   // stage and ys are free again
  }

  // ---------------------------------------------------------------------- flush
  __syncthreads();
  if (grad && p.dcomp && cur_b >= 0) flush_dcomp(cur_b);
  lacc = warp_sum(lacc);
  if (lane == 0) red[warp] = lacc;
  __syncthreads();
  if (tid == 0) p.loss_part[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
  if (grad) {
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < p.C) p.dW_part[((long long)blockIdx.x * p.C + c) * HD_DM + tid] = dw[c];
    if (p.amax) {
      amx = warp_max(amx);
      if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(p.amax), __float_as_int(amx));
    }
  }
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
__global__ void __launch_bounds__(1024) head_dw_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int n_part, int C) {
  // block = 256 channels x 4 groups of partials (fixed assignment and order: deterministic), SMEM reduction of the groups
  __shared__ float red[4][HD_DM];
  const int c = blockIdx.x, ch = threadIdx.x & 255, grp = threadIdx.x >> 8;
  const int per = (n_part + 3) / 4, i0 = grp * per, i1 = min(n_part, i0 + per);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int i = i0;
  for (; i + 4 <= i1; i += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += __ldg(part + ((long long)(i + j) * C + c) * HD_DM + ch);
  }
  for (; i < i1; ++i) acc[0] += __ldg(part + ((long long)i * C + c) * HD_DM + ch);
  red[grp][ch] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (grp == 0) out[c * HD_DM + ch] = (red[0][ch] + red[1][ch]) + (red[2][ch] + red[3][ch]);
}

// loss = (sum of the per-CTA partial sums, fixed order) / max(n_valid, 1)
__global__ void head_loss_kernel(const float* __restrict__ loss_part, int n_part, const int* __restrict__ n_valid,
                                 float* __restrict__ loss) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n_part; i += 32) s += loss_part[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) loss[0] = s / (float)max(__ldg(n_valid), 1);
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

template <int CMAX, int NK, int R>
static int launch_head_r(const HeadArgs& a0, int n_b, int grid_cap, cudaStream_t s) {
  HeadArgs a = a0;
  a.tiles_per_b = a.rows_pad / R;
  a.n_tiles = n_b * a.tiles_per_b;
  auto kern = csa_head_kernel<CMAX, NK, R>;
  constexpr int smem = HeadSmem<CMAX, NK, R>::BYTES;
  static_assert(smem <= 227 * 1024, "csa_head_kernel: shared memory budget");
  int dev = 0;
  CSN_CUDA_OK(cudaGetDevice(&dev));
  static bool configured[64] = {false};
  if (dev < 64 && !configured[dev]) {
    CSN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = true;
  }
  const int grid = a.n_tiles < grid_cap ? a.n_tiles : grid_cap;
  kern<<<grid, 256, smem, s>>>(a);
  CSN_LAUNCH_OK("csa_head_kernel");
  return 0;
}

// 16-row tiles when the ring (2 stages x n_k x 16 KB) fits next to the class-dependent buffers, else 8-row tiles
template <int CMAX>
static int launch_head(const HeadArgs& a, int n_b, int grid_cap, cudaStream_t s) {
  if constexpr (CMAX == 16) {
    // 8-row tiles, ~98 KB of SMEM: two CTAs per SM (16 warps) hide each other's barriers and LDS latencies
    if (a.n_k <= 1) return launch_head_r<CMAX, 1, 8>(a, n_b, 2 * grid_cap, s);
    if (a.n_k <= 2) return launch_head_r<CMAX, 2, 8>(a, n_b, 2 * grid_cap, s);
    if (a.n_k <= 4) return launch_head_r<CMAX, 4, 8>(a, n_b, 2 * grid_cap, s);
    return launch_head_r<CMAX, 6, 8>(a, n_b, grid_cap, s);
  } else {
  if (a.n_k <= 2) return launch_head_r<CMAX, 2, 8>(a, n_b, grid_cap, s);
  if (a.n_k <= 4) return launch_head_r<CMAX, 4, 8>(a, n_b, grid_cap, s);
  return launch_head_r<CMAX, 6, 8>(a, n_b, grid_cap, s);
  }
}

}  // namespace csn

extern "C" int csn_csa_head_grid(int32_t n_b, int32_t rows_pad) {
  (void)n_b; (void)rows_pad;
  return 2 * csn::num_sms();   // upper bound of the grid (persistent CTAs, at most two per SM): size of loss_part / dW_part
}

extern "C" int csn_csa_head(const float* Z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            const int32_t* blk, const float* w, int32_t n_b, int32_t n_k, const float* W, int32_t n_classes,
                            const int64_t* labels, int64_t lab_stride, int32_t ignore_index, int32_t* n_valid,
                            int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, float* loss_part,
                            float* dOutT, float* amax, double* dcomp, float* dW_part, float* dW, int32_t* stats, float* y_out,
                            float* loss, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Z && mean && rstd && gamma && beta && blk && w && W && labels && n_valid && loss_part && stats,
                "csn_csa_head: null pointer");
  CSN_CHECK_ARG(n_k >= 1 && n_k <= HD_MAXK, "csn_csa_head: 1..%d attention blocks per query (got %d)", HD_MAXK, n_k);
  CSN_CHECK_ARG(n_classes >= 1 && n_classes <= 64, "csn_csa_head: 1..64 classes supported (got %d)", n_classes);
  CSN_CHECK_ARG(chunk_pad % 16 == 0 && rows_pad % chunk_pad == 0 && chunk <= chunk_pad, "csn_csa_head: bad padding");
  CSN_CHECK_ARG(!dOutT || (dW_part && dW), "csn_csa_head: the backward outputs need dW_part and dW");
  CSN_CHECK_ARG((reinterpret_cast<uintptr_t>(Z) & 15) == 0 && (reinterpret_cast<uintptr_t>(mean) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(rstd) & 15) == 0, "csn_csa_head: Z / mean / rstd must be 16-byte aligned");
  if (n_b == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  head_count_valid_kernel<<<148, 256, 0, s>>>(reinterpret_cast<const long long*>(labels), lab_stride, n_b, n_points, ignore_index, n_valid);
  CSN_LAUNCH_OK("head_count_valid_kernel");
  const int grid_cap = num_sms();
  HeadArgs a{Z, mean, rstd, gamma, beta, blk, w, n_k, W, n_classes, reinterpret_cast<const long long*>(labels), lab_stride,
             ignore_index, n_valid, n_points, chunk, chunk_pad, rows_pad, 0, 0,
             loss_part, dOutT, amax, dcomp, dW_part, stats, y_out};
  int rc;
  if (n_classes <= 16) rc = launch_head<16>(a, n_b, grid_cap, s);
  else if (n_classes <= 32) rc = launch_head<32>(a, n_b, grid_cap, s);
  else rc = launch_head<64>(a, n_b, grid_cap, s);
  if (rc) return rc;
  if (loss) {   // (loss_part is zero-initialised by the caller: entries beyond the launched grid do not contribute)
    head_loss_kernel<<<1, 32, 0, s>>>(loss_part, csn_csa_head_grid(n_b, rows_pad), n_valid, loss);
    CSN_LAUNCH_OK("head_loss_kernel");
  }
  if (dOutT) {
    // the grid the launch used (mirrors launch_head): two CTAs per SM for <= 16 classes and n_k <= 4
    const int tiles = n_b * (rows_pad / 8);
    const int cap = (n_classes <= 16 && n_k <= 4) ? 2 * grid_cap : grid_cap;
    const int grid = tiles < cap ? tiles : cap;
    head_dw_reduce_kernel<<<n_classes, 1024, 0, s>>>(dW_part, dW, grid, n_classes);
    CSN_LAUNCH_OK("head_dw_reduce_kernel");
  }
  return 0;
}
