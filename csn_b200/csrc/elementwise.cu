// HBM-bound pieces of the CSA/SSA layer (everything that is not a contraction):
//   pack        : channel-major fp32 (B,256,N,1) -> chunk-padded row-major 16-bit (+fp32) rows
//   softmax     : fwd (fp32 scores -> 16-bit probabilities) and bwd (dS = P o (dP - rowsum(P o dP)) * scale)
//   add + LN    : z = fc_out + residual ; y = LayerNorm(z) (eps 1e-6, biased variance) ; pooled column sums
//   LN bwd      : dz, dgamma, dbeta
//   combine     : out = sum_k comp[b,k] * Y[pair(b,k)]  written channel-major (csa_models.py:232-240), and its bwd
// Reference lines: MID-FC/csa_models.py:92-94 (layout), :141 (softmax), :116-118 (residual + LayerNorm),
// :211-219 (mean over points), :232-240 (weighted sum + transpose back).
// Rows of every intermediate live in "padded" coordinates: chunk c of `chunk` (=500) points occupies
// rows [c*chunk_pad, c*chunk_pad + chunk) of a shape's block (chunk_pad = 512); pad rows are zero.
#include <cstdlib>
#include "host_util.h"
#include "ptx.cuh"

namespace csn {

constexpr int DM = 256;  // d_model on this path

__device__ __forceinline__ uint32_t pack2(float a, float b, int dtype) {
  if (dtype == CSN_F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t w, int dtype) {
  if (dtype == CSN_F16) return __half22float2(*reinterpret_cast<__half2*>(&w));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}

// ------------------------------------------------------------------------------------------ pack
template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

struct PackArgs {
  const void* src;      // channel-major (fp32, or 16-bit for csn_pack_rows_src16): element (slot, c, n) at src + i0*src_s0 + i1*src_s1 + c*ch_stride + n
  void* dst16;          // [slot][rows_pad][256] 16-bit
  float* dst32;         // same layout fp32, or null
  long long ch_stride, src_s0, src_s1;
  long long dst_slot0, dst_s0, dst_s1;  // destination slot = dst_slot0 + i0*dst_s0 + i1*dst_s1
  int n0, n1;
  int n_points;         // points used from the source (n_chunks * chunk)
  int chunk, chunk_pad, rows_pad;
  int dtype;
  float* amax;          // optional: max |src| over everything read (atomic max on the bit pattern)
  float* chunk_sum;     // optional [slot*n_chunks + chunk][256]: per-chunk channel sums of the valid points (atomically accumulated)
};

template <typename TS>
__global__ void __launch_bounds__(256) pack_kernel(const PackArgs p) {
  __shared__ float tile[32][DM + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.y / p.n1, i1 = blockIdx.y % p.n1;
  const TS* src = reinterpret_cast<const TS*>(p.src) + i0 * p.src_s0 + i1 * p.src_s1;
  const long long slot = p.dst_slot0 + i0 * p.dst_s0 + i1 * p.dst_s1;
  const int r0 = blockIdx.x * 32;  // first padded row of this tile (32 | chunk_pad)
  const int ch = r0 / p.chunk_pad, i = r0 % p.chunk_pad + lane;
  const int n = ch * p.chunk + i;
  const bool valid = i < p.chunk && n < p.n_points;
  float amx = 0.f;
  float* csum = p.chunk_sum ? p.chunk_sum + (slot * (p.rows_pad / p.chunk_pad) + ch) * DM : nullptr;
  for (int c = warp; c < DM; c += 8) {
    const float v = valid ? ld_as_float<TS>(src + c * p.ch_stride + n) : 0.f;
    tile[lane][c] = v;
    amx = fmaxf(amx, fabsf(v));
    if (csum) {
      const float sv = warp_sum(v);
      if (lane == 0) atomicAdd(csum + c, sv);
    }
  }
  if (p.amax) {
    amx = warp_max(amx);
    if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(p.amax), __float_as_int(amx));
  }
  __syncthreads();
  for (int rr = warp; rr < 32; rr += 8) {
    const long long row = slot * p.rows_pad + r0 + rr;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = g * 64 + 2 * lane;
      const float a = tile[rr][c], b = tile[rr][c + 1];
      if (p.dst16) reinterpret_cast<uint32_t*>(p.dst16)[(row * DM + c) >> 1] = pack2(a, b, p.dtype);
      if (p.dst32) *reinterpret_cast<float2*>(p.dst32 + row * DM + c) = make_float2(a, b);
    }
  }
}

// 128 points x 64 channels per CTA (grid.z = the four channel groups): every channel row is read as 512 contiguous
// bytes (four 128-byte requests back to back) instead of 128, which is what the DRAM pages want; the other three
// quarters of each output row come from the sibling CTAs and merge in L2.
template <typename TS>
__global__ void __launch_bounds__(256) pack128_kernel(const PackArgs p) {
  __shared__ float tile[128][65];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.y / p.n1, i1 = blockIdx.y % p.n1;
  const TS* src = reinterpret_cast<const TS*>(p.src) + i0 * p.src_s0 + i1 * p.src_s1;
  const long long slot = p.dst_slot0 + i0 * p.dst_s0 + i1 * p.dst_s1;
  const int r0 = blockIdx.x * 128;            // first padded row of this tile (128 | chunk_pad)
  const int c0 = blockIdx.z * 64;             // first channel of this CTA
  const int ch = r0 / p.chunk_pad, ib = r0 % p.chunk_pad;
  float amx = 0.f;
  float* csum = p.chunk_sum ? p.chunk_sum + (slot * (p.rows_pad / p.chunk_pad) + ch) * DM + c0 : nullptr;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = warp + 8 * k;
    const TS* row = src + (long long)(c0 + c) * p.ch_stride + (long long)ch * p.chunk;
    float sv = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = ib + lane + 32 * j;
      const int n = ch * p.chunk + i;
      const float v = (i < p.chunk && n < p.n_points) ? ld_as_float<TS>(row + i) : 0.f;
      tile[lane + 32 * j][c] = v;
      amx = fmaxf(amx, fabsf(v));
      sv += v;
    }
    if (csum) {
      sv = warp_sum(sv);
      if (lane == 0) atomicAdd(csum + c, sv);
    }
  }
  if (p.amax) {
    amx = warp_max(amx);
    if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(p.amax), __float_as_int(amx));
  }
  __syncthreads();
#pragma unroll 4
  for (int rr = warp; rr < 128; rr += 8) {
    const long long row = slot * p.rows_pad + r0 + rr;
    const int c = c0 + 2 * lane;
    const float a = tile[rr][2 * lane], b = tile[rr][2 * lane + 1];
    if (p.dst16) reinterpret_cast<uint32_t*>(p.dst16)[(row * DM + c) >> 1] = pack2(a, b, p.dtype);
    if (p.dst32) *reinterpret_cast<float2*>(p.dst32 + row * DM + c) = make_float2(a, b);
  }
}

// ------------------------------------------------------------------------------------------ softmax fwd
// One warp per row. cols_pad = 128*NV (NV <= 4 keeps the row in registers). Columns >= cols_valid are
// masked; rows whose index inside their group (row % group_rows) is >= rows_valid are written as zeros.
template <int NV>
__global__ void softmax_fwd_kernel(const float* __restrict__ S, void* __restrict__ P, long long rows,
                                   int cols_valid, int group_rows, int rows_valid, int dtype) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int CP = 128 * NV;
  uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(P) + row * CP);
  if ((int)(row % group_rows) >= rows_valid) {
#pragma unroll
    for (int j = 0; j < NV; ++j) dst[j * 32 + lane] = make_uint2(0u, 0u);
    return;
  }
  const float4* src = reinterpret_cast<const float4*>(S + row * CP);
  float4 v[NV];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = __ldg(src + j * 32 + lane);
    const int c = j * 128 + lane * 4;
    if (c + 0 >= cols_valid) v[j].x = -INFINITY;
    if (c + 1 >= cols_valid) v[j].y = -INFINITY;
    if (c + 2 >= cols_valid) v[j].z = -INFINITY;
    if (c + 3 >= cols_valid) v[j].w = -INFINITY;
    m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x = __expf(v[j].x - m); v[j].y = __expf(v[j].y - m);
    v[j].z = __expf(v[j].z - m); v[j].w = __expf(v[j].w - m);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    dst[j * 32 + lane] = make_uint2(pack2(v[j].x * inv, v[j].y * inv, dtype), pack2(v[j].z * inv, v[j].w * inv, dtype));
}

// Generic width (three passes over the row, served by L1/L2).
__global__ void softmax_fwd_wide_kernel(const float* __restrict__ S, void* __restrict__ P, long long rows,
                                        int cols_pad, int cols_valid, int group_rows, int rows_valid, int dtype) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  uint16_t* dst = reinterpret_cast<uint16_t*>(P) + row * cols_pad;
  const bool dead = (int)(row % group_rows) >= rows_valid;
  const float* src = S + row * cols_pad;
  float m = -INFINITY;
  if (!dead) for (int c = lane; c < cols_valid; c += 32) m = fmaxf(m, src[c]);
  m = warp_max(m);
  float s = 0.f;
  if (!dead) for (int c = lane; c < cols_valid; c += 32) s += __expf(src[c] - m);
  s = warp_sum(s);
  const float inv = dead ? 0.f : 1.f / s;
  for (int c = lane; c < cols_pad; c += 32) {
    const float pv = (c < cols_valid && !dead) ? __expf(src[c] - m) * inv : 0.f;
    dst[c] = (uint16_t)(pack2(pv, 0.f, dtype) & 0xFFFFu);
  }
}

// ------------------------------------------------------------------------------------------ softmax bwd
// dS = P o (dP - sum_j P_j dP_j) * scale, written as 16-bit with zeros in pad rows / columns.
__global__ void softmax_bwd_kernel(const void* __restrict__ P, const float* __restrict__ dP, void* __restrict__ dS,
                                   long long rows, int cols_pad, int cols_valid, int group_rows, int rows_valid,
                                   float scale, int dtype) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint32_t* p2 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(P) + row * cols_pad);
  const float2* g2 = reinterpret_cast<const float2*>(dP + row * cols_pad);
  uint32_t* d2 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(dS) + row * cols_pad);
  const bool dead = (int)(row % group_rows) >= rows_valid;
  const int npair = cols_pad >> 1;
  float dot = 0.f;
  if (!dead)
    for (int c = lane; c < npair; c += 32) {
      const float2 pv = unpack2(__ldg(p2 + c), dtype);
      const float2 gv = __ldg(g2 + c);
      if (2 * c < cols_valid) dot += pv.x * gv.x;
      if (2 * c + 1 < cols_valid) dot += pv.y * gv.y;
    }
  dot = warp_sum(dot);
  for (int c = lane; c < npair; c += 32) {
    float a = 0.f, b = 0.f;
    if (!dead) {
      const float2 pv = unpack2(__ldg(p2 + c), dtype);
      const float2 gv = __ldg(g2 + c);
      if (2 * c < cols_valid) a = pv.x * (gv.x - dot) * scale;
      if (2 * c + 1 < cols_valid) b = pv.y * (gv.y - dot) * scale;
    }
    d2[c] = pack2(a, b, dtype);
  }
}

// ------------------------------------------------------------------------------------------ add + LayerNorm fwd
// Block = 8 warps, 64 consecutive rows (8 per warp), all inside one `block_rows`-row block (64 | block_rows).
struct AddLnArgs {
  float* Z;              // in: fc output [rows][256]; out: z = fc + residual (kept for backward)
  const float* R;        // residual source rows [*][256] fp32
  const int* res_block;  // per block (= row / block_rows): index of the residual block
  float* Y;              // LayerNorm output
  void* Y16;             // optional 16-bit copy of Y (null to skip)
  float* mean; float* rstd;
  const float* gamma; const float* beta;
  float* colsum;         // optional [rows/64][256]: per-CTA partial sums of Y over its valid rows (csn_colsum_reduce adds them in a fixed order)
  long long rows;
  int block_rows;        // rows per (pair) block, e.g. 10240
  int group_rows, rows_valid;  // pad structure inside a block: row % group_rows < rows_valid is valid
  float eps;
  int dtype;
  uint32_t drop_seed, drop_thresh; float drop_scale;   // dropout on the fc output (csa_models.py:115), off when thresh == 0
  const uint32_t* drop_epoch;
};

// dropout mask of the 8 fc-output values a lane holds for one row (columns lane*4.. and 128 + lane*4..): four hashes
__device__ __forceinline__ void drop_apply8(uint32_t seed, uint32_t thresh, float scale, uint32_t row, int lane, float4& a, float4& c) {
  const uint32_t rk = drop_row_key(seed, row);
  const uint32_t h0 = drop_pair(rk, 2u * lane), h1 = drop_pair(rk, 2u * lane + 1u);
  const uint32_t h2 = drop_pair(rk, 64u + 2u * lane), h3 = drop_pair(rk, 65u + 2u * lane);
  a.x = drop_keep_lo(h0, thresh) ? a.x * scale : 0.f; a.y = drop_keep_hi(h0, thresh) ? a.y * scale : 0.f;
  a.z = drop_keep_lo(h1, thresh) ? a.z * scale : 0.f; a.w = drop_keep_hi(h1, thresh) ? a.w * scale : 0.f;
  c.x = drop_keep_lo(h2, thresh) ? c.x * scale : 0.f; c.y = drop_keep_hi(h2, thresh) ? c.y * scale : 0.f;
  c.z = drop_keep_lo(h3, thresh) ? c.z * scale : 0.f; c.w = drop_keep_hi(h3, thresh) ? c.w * scale : 0.f;
}

__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const AddLnArgs p) {
  __shared__ float red[8][DM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * 64;
  const long long blk = row0 / p.block_rows;
  const long long rblk = p.res_block ? p.res_block[blk] : blk;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + lane), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta) + lane), b1 = __ldg(reinterpret_cast<const float4*>(p.beta) + 32 + lane);
  float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) {
    const long long row = row0 + warp * 8 + i;
    if (row >= p.rows) break;
    const int rin = (int)(row - blk * p.block_rows);
    float4* z4 = reinterpret_cast<float4*>(p.Z + row * DM);
    float4* y4 = p.Y ? reinterpret_cast<float4*>(p.Y + row * DM) : nullptr;
    const bool valid = (rin % p.group_rows) < p.rows_valid;
    if (!valid) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      z4[lane] = zero; z4[32 + lane] = zero;
      if (y4) { y4[lane] = zero; y4[32 + lane] = zero; }
      if (p.Y16) {
        uint2* y16 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.Y16) + row * DM);
        y16[lane] = make_uint2(0, 0); y16[32 + lane] = make_uint2(0, 0);
      }
      if (lane == 0) { p.mean[row] = 0.f; p.rstd[row] = 0.f; }
      continue;
    }
    const float4* r4 = reinterpret_cast<const float4*>(p.R + (rblk * p.block_rows + rin) * DM);
    float4 a = z4[lane], c = z4[32 + lane];
    if (p.drop_thresh) drop_apply8(drop_seed_eff(p.drop_seed, p.drop_epoch), p.drop_thresh, p.drop_scale, (uint32_t)row, lane, a, c);
    const float4 ra = __ldg(r4 + lane), rc = __ldg(r4 + 32 + lane);
    a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
    c.x += rc.x; c.y += rc.y; c.z += rc.z; c.w += rc.w;
    z4[lane] = a; z4[32 + lane] = c;
    const float mu = warp_sum((a.x + a.y) + (a.z + a.w) + (c.x + c.y) + (c.z + c.w)) * (1.f / DM);
    const float dx0 = a.x - mu, dx1 = a.y - mu, dx2 = a.z - mu, dx3 = a.w - mu;
    const float dx4 = c.x - mu, dx5 = c.y - mu, dx6 = c.z - mu, dx7 = c.w - mu;
    const float var = warp_sum((dx0 * dx0 + dx1 * dx1) + (dx2 * dx2 + dx3 * dx3) + (dx4 * dx4 + dx5 * dx5) + (dx6 * dx6 + dx7 * dx7)) * (1.f / DM);
    const float rs = rsqrtf(var + p.eps);
    float4 ya, yc;
    ya.x = dx0 * rs * g0.x + b0.x; ya.y = dx1 * rs * g0.y + b0.y; ya.z = dx2 * rs * g0.z + b0.z; ya.w = dx3 * rs * g0.w + b0.w;
    yc.x = dx4 * rs * g1.x + b1.x; yc.y = dx5 * rs * g1.y + b1.y; yc.z = dx6 * rs * g1.z + b1.z; yc.w = dx7 * rs * g1.w + b1.w;
    if (y4) { y4[lane] = ya; y4[32 + lane] = yc; }
    if (p.Y16) {
      uint2* y16 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.Y16) + row * DM);
      y16[lane] = make_uint2(pack2(ya.x, ya.y, p.dtype), pack2(ya.z, ya.w, p.dtype));
      y16[32 + lane] = make_uint2(pack2(yc.x, yc.y, p.dtype), pack2(yc.z, yc.w, p.dtype));
    }
    if (lane == 0) { p.mean[row] = mu; p.rstd[row] = rs; }
    cs[0] += ya.x; cs[1] += ya.y; cs[2] += ya.z; cs[3] += ya.w;
    cs[4] += yc.x; cs[5] += yc.y; cs[6] += yc.z; cs[7] += yc.w;
  }
  if (p.colsum) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[warp][lane * 4 + j] = cs[j]; red[warp][128 + lane * 4 + j] = cs[4 + j]; }
    __syncthreads();
    const int c = threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    p.colsum[(long long)blockIdx.x * DM + c] = s;
  }
}

// Per-64-row partial column sums of y = (z - mean)*rstd*gamma + beta over the valid rows, for rows whose statistics
// were already produced (csn_gemm_res_ln); the same partial layout as add_ln_fwd_kernel's colsum output.
__global__ void __launch_bounds__(256) ln_colsum_kernel(const float* __restrict__ Z, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ colsum,
                                                        int block_rows, int group_rows, int rows_valid) {
  __shared__ float red[8][DM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * 64;
  const int rin0 = (int)(row0 % block_rows);
  float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sc = sa;
  float cnt = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = row0 + warp * 8 + i;
    const int rin = rin0 + warp * 8 + i;
    if ((rin % group_rows) >= rows_valid) continue;
    const float4* z4 = reinterpret_cast<const float4*>(Z + row * DM);
    const float4 a = __ldg(z4 + lane), c = __ldg(z4 + 32 + lane);
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    sa.x += (a.x - mu) * rs; sa.y += (a.y - mu) * rs; sa.z += (a.z - mu) * rs; sa.w += (a.w - mu) * rs;
    sc.x += (c.x - mu) * rs; sc.y += (c.y - mu) * rs; sc.z += (c.z - mu) * rs; sc.w += (c.w - mu) * rs;
    cnt += 1.f;
  }
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
  red[warp][lane * 4 + 0] = sa.x * g0.x + cnt * b0.x; red[warp][lane * 4 + 1] = sa.y * g0.y + cnt * b0.y;
  red[warp][lane * 4 + 2] = sa.z * g0.z + cnt * b0.z; red[warp][lane * 4 + 3] = sa.w * g0.w + cnt * b0.w;
  red[warp][128 + lane * 4 + 0] = sc.x * g1.x + cnt * b1.x; red[warp][128 + lane * 4 + 1] = sc.y * g1.y + cnt * b1.y;
  red[warp][128 + lane * 4 + 2] = sc.z * g1.z + cnt * b1.z; red[warp][128 + lane * 4 + 3] = sc.w * g1.w + cnt * b1.w;
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  colsum[(long long)blockIdx.x * DM + c] = s;
}

// out[b][c] = scale * sum_{i < parts} part[(b*parts + i)][c], fixed order (deterministic pooled means)
__global__ void colsum_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int parts, float scale) {
  const int b = blockIdx.x, c = threadIdx.x;
  const float* src = part + (long long)b * parts * DM + c;
  float s = 0.f;
  for (int i = 0; i < parts; ++i) s += src[(long long)i * DM];
  out[(long long)b * DM + c] = s * scale;
}

// ------------------------------------------------------------------------------------------ LayerNorm bwd
struct LnBwdArgs {
  const float* dY; const float* Z; const float* mean; const float* rstd; const float* gamma;
  float* dZ;        // fp32 [rows][256] (also the gradient of the residual input)
  void* dZ16;       // 16-bit copy for the tensor-core contractions
  float* dgamma; float* dbeta;  // [256], atomically accumulated
  long long rows;
  int group_rows, rows_valid, block_rows;
  int dtype;
  const float* amax;   // optional: dY is multiplied by 2^floor(log2(128/amax)) on load (power-of-two loss scaling)
  const float* bcast;  // optional [n][256]: row vector added to every valid row of a block before anything else
  const int* bcast_idx;  // per block: row of `bcast` (or -1)
  float bcast_scale;
  // implicit upstream gradient: dY[block] = src_w[block] * dYsrc[src_idx[block]] (rows of one shape), or 0 if src_idx < 0
  const int* src_idx; const float* src_w;
  int debug;
  float* chunk_gsum;   // optional [rows/group_rows][256]: per-chunk column sums of dZ (atomically accumulated)
  // dropout on the fc output: the 16-bit copy (what flows on into fc / the attention output) is masked and scaled,
  // the fp32 dZ (gradient of the residual input) is not
  uint32_t drop_seed, drop_thresh; float drop_scale;
  const uint32_t* drop_epoch;
};

// Block = 8 warps x 8 rows = 64 consecutive rows: they lie inside one attention block AND one chunk (64 | group_rows),
// so the valid / pad split, the upstream-gradient source, the broadcast row and every base pointer are CTA constants;
// the row loop is branch-free FFMA chains (~130 instructions per row), two rows in flight per warp.
template <int MINB, bool GS, bool DROP>
__global__ void __launch_bounds__(256, MINB) ln_bwd_kernel(const LnBwdArgs p) {
  __shared__ float red[2][8][DM];
  // per-chunk column sums of dZ: every warp accumulates its rows in a private SMEM row (each lane owns its 8 channels:
  // plain read-modify-write, no atomics, no registers)
  __shared__ __align__(16) float gsm[GS ? 8 : 1][GS ? DM : 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (GS) {
#pragma unroll
    for (int w = 0; w < 8; ++w) gsm[w][threadIdx.x] = 0.f;
    __syncthreads();
  }
  const long long row0 = (long long)blockIdx.x * 64;
  const int blk = (int)(row0 / p.block_rows);
  const int rin0 = (int)(row0 - (long long)blk * p.block_rows);
  const int i0 = rin0 % p.group_rows;                         // first row's index inside its chunk
  const int nv = max(0, min(64, p.rows_valid - i0));          // rows [0, nv) of the CTA are real points, the rest padding
  const int w0 = warp * 8;
  const int nvw = max(0, min(8, nv - w0));                    // valid rows of this warp (warp-uniform)
  const float gscale = p.amax ? exp2f(floorf(log2f(128.f / fmaxf(__ldg(p.amax), 1e-30f)))) : 1.f;
  const int src_i = p.src_idx ? __ldg(p.src_idx + blk) : 0;
  const bool has_dy = !p.src_idx || src_i >= 0;               // CTA-uniform
  const float wg = (p.src_idx ? (has_dy ? __ldg(p.src_w + blk) : 0.f) : 1.f) * gscale;
  const int bc_i = p.bcast ? __ldg(p.bcast_idx + blk) : -1;
  // per-channel constants live in SMEM (two LDS.128 pairs per row instead of 16 registers held across the kernel):
  // gamma, and the block's broadcast row (pooled-mean gradient) pre-scaled
  __shared__ __align__(16) float cst[2][DM];
  cst[0][threadIdx.x] = __ldg(p.gamma + threadIdx.x);
  cst[1][threadIdx.x] = bc_i >= 0 ? __ldg(p.bcast + (long long)bc_i * DM + threadIdx.x) * (p.bcast_scale * gscale) : 0.f;
  __syncthreads();
  float dg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, db[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // base pointers of this warp's 8 rows (one 64-bit computation each)
  const long long wrow = row0 + w0;
  const float4* z4 = reinterpret_cast<const float4*>(p.Z + wrow * DM) + lane;
  const long long drow = p.src_idx ? ((long long)(has_dy ? src_i : 0) * p.block_rows + rin0 + w0) : wrow;
  const float4* d4 = reinterpret_cast<const float4*>(p.dY + drow * DM) + lane;
  const float* mup = p.mean + wrow;
  const float* rsp = p.rstd + wrow;
  uint2* dz16 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.dZ16) + wrow * DM) + lane;
  float4* dz4 = p.dZ ? reinterpret_cast<float4*>(p.dZ + wrow * DM) + lane : nullptr;

  struct Row { float4 za, zc, da, dc; float mu, rs; };
  auto load_row = [&](int i) -> Row {   // all global loads of a row, issued together
    Row r;
    r.za = __ldg(z4 + i * 64); r.zc = __ldg(z4 + i * 64 + 32);
    r.da = make_float4(0.f, 0.f, 0.f, 0.f); r.dc = r.da;
    if (has_dy) { r.da = __ldg(d4 + i * 64); r.dc = __ldg(d4 + i * 64 + 32); }
    r.mu = __ldg(mup + i); r.rs = __ldg(rsp + i);
    return r;
  };
  auto finish_row = [&](const Row& r, int i) {
    const float4 g0 = *reinterpret_cast<const float4*>(&cst[0][lane * 4]), g1 = *reinterpret_cast<const float4*>(&cst[0][128 + lane * 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&cst[1][lane * 4]), b1 = *reinterpret_cast<const float4*>(&cst[1][128 + lane * 4]);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float zz[8] = {r.za.x, r.za.y, r.za.z, r.za.w, r.zc.x, r.zc.y, r.zc.z, r.zc.w};
    const float dd[8] = {r.da.x, r.da.y, r.da.z, r.da.w, r.dc.x, r.dc.y, r.dc.z, r.dc.w};
    const float rs = r.rs, nb = -r.mu * r.rs;
    float xh[8], g[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = fmaf(zz[j], rs, nb);
      const float dy = fmaf(dd[j], wg, bc[j]);
      g[j] = dy * gm[j];
      s1 += g[j];
      s2 = fmaf(g[j], xh[j], s2);
      dg[j] = fmaf(dy, xh[j], dg[j]);
      db[j] += dy;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float c1 = -rs * s1 * (1.f / DM), c2 = -rs * s2 * (1.f / DM);
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = fmaf(xh[j], c2, fmaf(g[j], rs, c1));
    if (dz4) {
      dz4[i * 64] = make_float4(o8[0], o8[1], o8[2], o8[3]);
      dz4[i * 64 + 32] = make_float4(o8[4], o8[5], o8[6], o8[7]);
    }
    if (GS) {
      float4* ga = reinterpret_cast<float4*>(&gsm[warp][lane * 4]);
      float4* gc = reinterpret_cast<float4*>(&gsm[warp][128 + lane * 4]);
      float4 ta = *ga, tc = *gc;
      ta.x += o8[0]; ta.y += o8[1]; ta.z += o8[2]; ta.w += o8[3];
      tc.x += o8[4]; tc.y += o8[5]; tc.z += o8[6]; tc.w += o8[7];
      *ga = ta; *gc = tc;
    }
    if (DROP) {   // the 16-bit copy flows on into fc / the attention output: masked and scaled; the fp32 dZ is not
      float4 ma = make_float4(o8[0], o8[1], o8[2], o8[3]), mc = make_float4(o8[4], o8[5], o8[6], o8[7]);
      drop_apply8(drop_seed_eff(p.drop_seed, p.drop_epoch), p.drop_thresh, p.drop_scale, (uint32_t)(wrow + i), lane, ma, mc);
      o8[0] = ma.x; o8[1] = ma.y; o8[2] = ma.z; o8[3] = ma.w; o8[4] = mc.x; o8[5] = mc.y; o8[6] = mc.z; o8[7] = mc.w;
    }
    dz16[i * 64] = make_uint2(pack2(o8[0], o8[1], p.dtype), pack2(o8[2], o8[3], p.dtype));
    dz16[i * 64 + 32] = make_uint2(pack2(o8[4], o8[5], p.dtype), pack2(o8[6], o8[7], p.dtype));
  };
  auto zero_row = [&](int i) {
    if (dz4) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      dz4[i * 64] = zero; dz4[i * 64 + 32] = zero;
    }
    dz16[i * 64] = make_uint2(0u, 0u); dz16[i * 64 + 32] = make_uint2(0u, 0u);
  };
  if (nv == 64) {
    // 7 of the 8 CTAs of a chunk: every row real.  The condition depends on blockIdx and kernel parameters only, so
    // the branch is provably uniform (no warp-sync wrappers around the shuffles); fully unrolled, two rows in flight.
    Row cur = load_row(0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      Row nxt;
      if (i + 1 < 8) nxt = load_row(i + 1);
      finish_row(cur, i);
      if (i + 1 < 8) cur = nxt;
    }
  } else {
    // the chunk's tail (pad rows): one row at a time
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      if (i < nvw) finish_row(load_row(i), i);
      else zero_row(i);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][warp][lane * 4 + j] = dg[j]; red[0][warp][128 + lane * 4 + j] = dg[4 + j];
    red[1][warp][lane * 4 + j] = db[j]; red[1][warp][128 + lane * 4 + j] = db[4 + j];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f, t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { s += red[0][w][c]; t += red[1][w][c]; }
  if (p.debug & 1) return;
  atomicAdd(p.dgamma + c, s);
  atomicAdd(p.dbeta + c, t);
  if (GS) {   // gsm is complete after the barrier above
    float u = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) u += gsm[w][c];
    atomicAdd(p.chunk_gsum + (row0 / p.group_rows) * DM + c, u);
  }
}

// ------------------------------------------------------------------------------------------ combine fwd
// out[b][c][n] = sum_k w[b*n_k + k] * Y[blk[b*n_k + k]][padrow(n)][c]      (channel-major output)
struct LnParams {   // when mean != nullptr the row source holds PRE-LayerNorm rows z and y = (z-mean)*rstd*gamma+beta is formed on the fly
  const float* mean; const float* rstd; const float* gamma; const float* beta;
};

__device__ __forceinline__ void ln_apply(const LnParams& ln, long long row, int lane, float4& a, float4& c) {
  if (ln.mean == nullptr) return;
  const float mu = __ldg(ln.mean + row), rs = __ldg(ln.rstd + row);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(ln.gamma) + lane), g1 = __ldg(reinterpret_cast<const float4*>(ln.gamma) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(ln.beta) + lane), b1 = __ldg(reinterpret_cast<const float4*>(ln.beta) + 32 + lane);
  a.x = (a.x - mu) * rs * g0.x + b0.x; a.y = (a.y - mu) * rs * g0.y + b0.y; a.z = (a.z - mu) * rs * g0.z + b0.z; a.w = (a.w - mu) * rs * g0.w + b0.w;
  c.x = (c.x - mu) * rs * g1.x + b1.x; c.y = (c.y - mu) * rs * g1.y + b1.y; c.z = (c.z - mu) * rs * g1.z + b1.z; c.w = (c.w - mu) * rs * g1.w + b1.w;
}

struct CombineArgs {
  const float* Y; const int* blk; const float* w; float* out; void* rows16;  // rows16: optional 16-bit row-major copy [b][rows_pad][256]
  long long out_b_stride, out_ch_stride;
  int n_k, n_points, chunk, chunk_pad, rows_pad;
  int dtype;
  LnParams ln;
};

__global__ void __launch_bounds__(256) combine_fwd_kernel(const CombineArgs p) {
  __shared__ float tile[DM][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * 32;
  for (int rr = warp; rr < 32; rr += 8) {
    const int r = r0 + rr;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    for (int k = 0; k < p.n_k; ++k) {
      const float wk = __ldg(p.w + b * p.n_k + k);
      const long long yrow = (long long)__ldg(p.blk + b * p.n_k + k) * p.rows_pad + r;
      const float4* y4 = reinterpret_cast<const float4*>(p.Y + yrow * DM);
      float4 ya = __ldg(y4 + lane), yc = __ldg(y4 + 32 + lane);
      ln_apply(p.ln, yrow, lane, ya, yc);
      a.x += wk * ya.x; a.y += wk * ya.y; a.z += wk * ya.z; a.w += wk * ya.w;
      c.x += wk * yc.x; c.y += wk * yc.y; c.z += wk * yc.z; c.w += wk * yc.w;
    }
    tile[lane * 4 + 0][rr] = a.x; tile[lane * 4 + 1][rr] = a.y; tile[lane * 4 + 2][rr] = a.z; tile[lane * 4 + 3][rr] = a.w;
    tile[128 + lane * 4 + 0][rr] = c.x; tile[128 + lane * 4 + 1][rr] = c.y; tile[128 + lane * 4 + 2][rr] = c.z; tile[128 + lane * 4 + 3][rr] = c.w;
    if (p.rows16) {
      uint2* o16 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.rows16) + ((long long)b * p.rows_pad + r) * DM);
      o16[lane] = make_uint2(pack2(a.x, a.y, p.dtype), pack2(a.z, a.w, p.dtype));
      o16[32 + lane] = make_uint2(pack2(c.x, c.y, p.dtype), pack2(c.z, c.w, p.dtype));
    }
  }
  __syncthreads();
  const int ch = r0 / p.chunk_pad, i = r0 % p.chunk_pad + lane;
  const int n = ch * p.chunk + i;
  if (i < p.chunk && n < p.n_points) {
    float* o = p.out + b * p.out_b_stride + n;
    for (int c = warp; c < DM; c += 8) o[c * p.out_ch_stride] = tile[c][lane];
  }
}

// ------------------------------------------------------------------------------------------ combine bwd
// For every destination block j (one MHA output):
//   dY[j][row][c] = cw[j] * dOut[cb[j]][c][n(row)]  +  pw[j] * dpool[pb[j]][c]      (valid rows; 0 in pads)
// and, when cw_index[j] >= 0,  dcomp[cw_index[j]] += sum_{row,c} dOut[cb[j]][c][n] * Y[j][row][c].
struct CombineBwdArgs {
  const float* dOut; const float* Y; const float* dpool;
  const int* cb;        // batch item whose output gradient feeds block j (or -1)
  const float* cw;      // its compatibility weight
  const int* cw_index;  // flat index into dcomp (or -1)
  const int* pb;        // pooled-slot index feeding block j (or -1)
  float pool_scale;     // 1 / n_points
  float* dY; float* dcomp;
  long long out_b_stride, out_ch_stride;
  int n_points, chunk, chunk_pad, rows_pad;
  float* amax;          // optional: max |dY| over everything written (atomic max on the bit pattern)
  LnParams ln;
};

__global__ void __launch_bounds__(256) combine_bwd_kernel(const CombineBwdArgs p) {
  __shared__ float tile[DM][33];
  __shared__ float wred[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.y;
  const int r0 = blockIdx.x * 32;
  const int b = p.cb[j];
  const int pbj = p.pb[j];
  const int ch = r0 / p.chunk_pad, i = r0 % p.chunk_pad + lane;
  const int n = ch * p.chunk + i;
  const bool lane_valid = i < p.chunk && n < p.n_points;
  if (b >= 0) {
    const float* g = p.dOut + b * p.out_b_stride + n;
    for (int c = warp; c < DM; c += 8) tile[c][lane] = lane_valid ? __ldg(g + c * p.out_ch_stride) : 0.f;
  }
  __syncthreads();
  const float cwj = b >= 0 ? p.cw[j] : 0.f;
  float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pc = pa;
  if (pbj >= 0 && p.dpool != nullptr) {
    const float4* d4 = reinterpret_cast<const float4*>(p.dpool + (long long)pbj * DM);
    pa = __ldg(d4 + lane); pc = __ldg(d4 + 32 + lane);
    pa.x *= p.pool_scale; pa.y *= p.pool_scale; pa.z *= p.pool_scale; pa.w *= p.pool_scale;
    pc.x *= p.pool_scale; pc.y *= p.pool_scale; pc.z *= p.pool_scale; pc.w *= p.pool_scale;
  }
  float dot = 0.f;
  float amx = 0.f;
  for (int rr = warp; rr < 32; rr += 8) {
    const int r = r0 + rr;
    const int ii = r % p.chunk_pad;
    const bool valid = ii < p.chunk && (r / p.chunk_pad) * p.chunk + ii < p.n_points;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (valid) {
      a = pa; c = pc;
      if (b >= 0) {
        const float4 ga = make_float4(tile[lane * 4 + 0][rr], tile[lane * 4 + 1][rr], tile[lane * 4 + 2][rr], tile[lane * 4 + 3][rr]);
        const float4 gc = make_float4(tile[128 + lane * 4 + 0][rr], tile[128 + lane * 4 + 1][rr], tile[128 + lane * 4 + 2][rr], tile[128 + lane * 4 + 3][rr]);
        a.x += cwj * ga.x; a.y += cwj * ga.y; a.z += cwj * ga.z; a.w += cwj * ga.w;
        c.x += cwj * gc.x; c.y += cwj * gc.y; c.z += cwj * gc.z; c.w += cwj * gc.w;
        if (p.dcomp) {
        const long long yrow = (long long)j * p.rows_pad + r;
        const float4* y4 = reinterpret_cast<const float4*>(p.Y + yrow * DM);
        float4 ya = __ldg(y4 + lane), yc = __ldg(y4 + 32 + lane);
        ln_apply(p.ln, yrow, lane, ya, yc);
        dot += ga.x * ya.x + ga.y * ya.y + ga.z * ya.z + ga.w * ya.w + gc.x * yc.x + gc.y * yc.y + gc.z * yc.z + gc.w * yc.w;
        }
      }
    }
    if (p.dY) {
      float4* d4 = reinterpret_cast<float4*>(p.dY + ((long long)j * p.rows_pad + r) * DM);
      d4[lane] = a; d4[32 + lane] = c;
      amx = fmaxf(amx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      amx = fmaxf(amx, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
    }
  }
  if (p.amax && p.dY) {
    amx = warp_max(amx);
    if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(p.amax), __float_as_int(amx));
  }
  if (p.dcomp && b >= 0 && p.cw_index[j] >= 0) {
    dot = warp_sum(dot);
    if (lane == 0) wred[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += wred[w];
      atomicAdd(p.dcomp + p.cw_index[j], s);
    }
  }
}

// ------------------------------------------------------------------------------------------ block dot
// dcomp[out_idx[j]] += sum over the valid rows of block j of < G[src_idx[j]][row], y_j[row] >, where y_j is the
// (re-normalised) LayerNorm output of block j and G holds one block of rows per source (the transposed
// output gradient).  64 rows per CTA, one warp per row, one atomic per CTA.
struct BlockDotArgs {
  const float* G; const float* Z; const int* src_idx; const int* out_idx; float* out;
  long long rows; int block_rows, group_rows, rows_valid;
  LnParams ln;
};

__global__ void __launch_bounds__(256) block_dot_kernel(const BlockDotArgs p) {
  __shared__ float wred[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * 64;
  const int blk = (int)(row0 / p.block_rows);
  const int rin0 = (int)(row0 - (long long)blk * p.block_rows);
  const int si = __ldg(p.src_idx + blk), oi = __ldg(p.out_idx + blk);
  if (si < 0 || oi < 0) return;   // uniform per CTA
  float dot = 0.f;
  for (int i = 0; i < 8; ++i) {
    const int rin = rin0 + warp * 8 + i;
    const long long row = row0 + warp * 8 + i;
    if (row >= p.rows || (rin % p.group_rows) >= p.rows_valid) continue;
    const float4* g4 = reinterpret_cast<const float4*>(p.G + ((long long)si * p.block_rows + rin) * DM);
    const float4* z4 = reinterpret_cast<const float4*>(p.Z + row * DM);
    const float4 ga = __ldg(g4 + lane), gc = __ldg(g4 + 32 + lane);
    float4 ya = __ldg(z4 + lane), yc = __ldg(z4 + 32 + lane);
    ln_apply(p.ln, row, lane, ya, yc);
    dot += ga.x * ya.x + ga.y * ya.y + ga.z * ya.z + ga.w * ya.w + gc.x * yc.x + gc.y * yc.y + gc.z * yc.z + gc.w * yc.w;
  }
  dot = warp_sum(dot);
  if (lane == 0) wred[warp] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += wred[w];
    atomicAdd(p.out + oi, t);
  }
}

// x[i] *= 1 / 2^floor(log2(128 / amax)): undoes the power-of-two loss scaling csn_ln_bwd applied (same formula), for
// the flat buffer that holds every parameter gradient of a step (one launch instead of one per tensor)
__global__ void grad_unscale_kernel(float* __restrict__ x, long long n, const float* __restrict__ amax) {
  const float inv = 1.f / exp2f(floorf(log2f(128.f / fmaxf(__ldg(amax), 1e-30f))));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= inv;
}

template <typename K, typename A>
static int launch_simple(K kern, dim3 grid, dim3 block, const A& args, void* stream, const char* name) {
  kern<<<grid, block, 0, (cudaStream_t)stream>>>(args);
  CSN_LAUNCH_OK(name);
  return 0;
}

}  // namespace csn

namespace csn {
// Concatenated rows -> padded slots: out[s][r] = x[offsets[s] + r] for r < len(s), 0 for len(s) <= r < n_pad
// (fp32 copy and / or 16-bit copy).  One warp per row, 256 columns.
__global__ void __launch_bounds__(256) ragged_pad_kernel(const float* __restrict__ x, const long long* __restrict__ offsets,
                                                         int n_pad, float* __restrict__ out32, void* __restrict__ out16, int dtype,
                                                         float* __restrict__ amax) {
  const int s = blockIdx.y;
  const long long o0 = offsets[s];
  const int len = (int)(offsets[s + 1] - o0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float amx = 0.f;
  for (int r = blockIdx.x * 8 + warp; r < n_pad; r += gridDim.x * 8) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < len) {
      const float4* src = reinterpret_cast<const float4*>(x + (o0 + r) * 256) + lane * 2;
      a = __ldg(src);
      b = __ldg(src + 1);
      amx = fmaxf(amx, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                             fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
    }
    const long long row = (long long)s * n_pad + r;
    if (out32) {
      float4* d = reinterpret_cast<float4*>(out32 + row * 256) + lane * 2;
      d[0] = a;
      d[1] = b;
    }
    if (out16) {
      uint4 w;
      if (dtype == CSN_F16) {
        __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
        w = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      } else {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w), h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
        w = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      }
      reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out16) + row * 256)[lane] = w;
    }
  }
  if (amax) {   // max |x| over everything copied (non-negative floats order like their bit patterns)
    amx = warp_max(amx);
    if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(amx));
  }
}
}  // namespace csn

namespace csn {
// dst[s][e] = (dst[s][e] + sum over source blocks j with dst_block[j] == s of src[j][e]) * unscale:
// the residual path of the attention backward (dX[query slot] += dZ[block]) as a deterministic gather.
__global__ void __launch_bounds__(256) block_add_kernel(const float* __restrict__ src, const int* __restrict__ dst_block, int n_src,
                                                        float* __restrict__ dst, long long block_elems, const float* __restrict__ amax) {
  __shared__ int mine[256];
  __shared__ int n_mine;
  const int s = blockIdx.y;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int j = 0; j < n_src; ++j)
      if (dst_block[j] == s) {
        if (n == 256) asm volatile("trap;");   // more sources per destination than the table holds: fail loudly
        mine[n++] = j;
      }
    n_mine = n;
  }
  __syncthreads();
  const float inv = amax ? 1.f / exp2f(floorf(log2f(128.f / fmaxf(__ldg(amax), 1e-30f)))) : 1.f;
  const int n = n_mine;
  float4* d4 = reinterpret_cast<float4*>(dst + (long long)s * block_elems);
  const long long n4 = block_elems >> 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    float4 a = d4[e];
    for (int k = 0; k < n; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)mine[k] * block_elems) + e);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    d4[e] = a;
  }
}
}  // namespace csn

namespace csn {
// out[s][c] = mean of x[offsets[s] .. offsets[s+1])[c]: CTA = (segment, 32-column group), 32 row lanes per column,
// four loads in flight per thread (a handful of long segments must keep the memory system busy on their own).
__global__ void __launch_bounds__(1024) segment_mean_kernel(const float* __restrict__ x, const long long* __restrict__ offsets,
                                                            int n_cols, float* __restrict__ out) {
  __shared__ float red[32][33];
  const int seg = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const long long r0 = offsets[seg], r1 = offsets[seg + 1];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < n_cols) {
    const float* p = x + c;
    long long r = r0 + rl;
    for (; r + 96 < r1; r += 128) {
      a0 += __ldg(p + r * n_cols);
      a1 += __ldg(p + (r + 32) * n_cols);
      a2 += __ldg(p + (r + 64) * n_cols);
      a3 += __ldg(p + (r + 96) * n_cols);
    }
    for (; r < r1; r += 32) a0 += __ldg(p + r * n_cols);
  }
  red[rl][threadIdx.x & 31] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (rl == 0 && c < n_cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    out[(long long)seg * n_cols + c] = r1 > r0 ? t / (float)(r1 - r0) : 0.f;
  }
}
}  // namespace csn

template <typename TS>
static int launch_pack(const csn::PackArgs& a, void* stream) {
  using namespace csn;
  static const bool wide = getenv("CSN_PACK128") == nullptr || atoi(getenv("CSN_PACK128")) != 0;
  if (wide && a.chunk_pad % 128 == 0)
    return launch_simple(pack128_kernel<TS>, dim3(a.rows_pad / 128, a.n0 * a.n1, 4), dim3(256), a, stream, "pack128_kernel");
  return launch_simple(pack_kernel<TS>, dim3(a.rows_pad / 32, a.n0 * a.n1), dim3(256), a, stream, "pack_kernel");
}

static int pack_rows_common(const void* src, int32_t src_dtype, void* dst16, float* dst32, int64_t ch_stride, int32_t n0,
                            int64_t src_s0, int32_t n1, int64_t src_s1, int64_t dst_slot0, int64_t dst_s0, int64_t dst_s1,
                            int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, int32_t dtype,
                            float* amax, float* chunk_sum, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(src && (dst16 || dst32), "csn_pack_rows: null pointer");
  CSN_CHECK_ARG(chunk_pad % 32 == 0 && rows_pad % chunk_pad == 0 && chunk <= chunk_pad, "csn_pack_rows: bad padding (chunk=%d chunk_pad=%d rows_pad=%d)", chunk, chunk_pad, rows_pad);
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "csn_pack_rows: 16-bit destination only");
  if (n0 * n1 == 0) return 0;
  PackArgs a{src, dst16, dst32, ch_stride, src_s0, src_s1, dst_slot0, dst_s0, dst_s1, n0, n1, n_points, chunk, chunk_pad, rows_pad, dtype, amax, chunk_sum};
  if (src_dtype == CSN_F16) return launch_pack<__half>(a, stream);
  if (src_dtype == CSN_BF16) return launch_pack<__nv_bfloat16>(a, stream);
  return launch_pack<float>(a, stream);
}

extern "C" {

int csn_pack_rows(const float* src, void* dst16, float* dst32, int64_t ch_stride, int32_t n0, int64_t src_s0,
                  int32_t n1, int64_t src_s1, int64_t dst_slot0, int64_t dst_s0, int64_t dst_s1,
                  int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, int32_t dtype,
                  float* amax, float* chunk_sum, void* stream) {
  return pack_rows_common(src, -1, dst16, dst32, ch_stride, n0, src_s0, n1, src_s1, dst_slot0, dst_s0, dst_s1, n_points, chunk,
                          chunk_pad, rows_pad, dtype, amax, chunk_sum, stream);
}

int csn_pack_rows_src16(const void* src, int32_t src_dtype, void* dst16, float* dst32, int64_t ch_stride, int32_t n0,
                        int64_t src_s0, int32_t n1, int64_t src_s1, int64_t dst_slot0, int64_t dst_s0, int64_t dst_s1,
                        int32_t n_points, int32_t chunk, int32_t chunk_pad, int32_t rows_pad, int32_t dtype,
                        float* amax, float* chunk_sum, void* stream) {
  if (src_dtype != CSN_F16 && src_dtype != CSN_BF16) {
    csn::clear_error();
    CSN_CHECK_ARG(false, "csn_pack_rows_src16: source dtype must be f16 or bf16");
  }
  return pack_rows_common(src, src_dtype, dst16, dst32, ch_stride, n0, src_s0, n1, src_s1, dst_slot0, dst_s0, dst_s1, n_points,
                          chunk, chunk_pad, rows_pad, dtype, amax, chunk_sum, stream);
}

int csn_softmax_fwd(const float* S, void* P, int64_t rows, int32_t cols_pad, int32_t cols_valid,
                    int32_t group_rows, int32_t rows_valid, int32_t dtype, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(S && P, "csn_softmax_fwd: null pointer");
  CSN_CHECK_ARG(cols_pad % 8 == 0 && cols_valid <= cols_pad && cols_valid > 0, "csn_softmax_fwd: bad widths %d/%d", cols_valid, cols_pad);
  if (rows == 0) return 0;
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  cudaStream_t s = (cudaStream_t)stream;
  if (cols_pad == 512) softmax_fwd_kernel<4><<<grid, wpb * 32, 0, s>>>(S, P, rows, cols_valid, group_rows, rows_valid, dtype);
  else if (cols_pad == 256) softmax_fwd_kernel<2><<<grid, wpb * 32, 0, s>>>(S, P, rows, cols_valid, group_rows, rows_valid, dtype);
  else if (cols_pad == 128) softmax_fwd_kernel<1><<<grid, wpb * 32, 0, s>>>(S, P, rows, cols_valid, group_rows, rows_valid, dtype);
  else softmax_fwd_wide_kernel<<<grid, wpb * 32, 0, s>>>(S, P, rows, cols_pad, cols_valid, group_rows, rows_valid, dtype);
  CSN_LAUNCH_OK("softmax_fwd_kernel");
  return 0;
}

int csn_softmax_bwd(const void* P, const float* dP, void* dS, int64_t rows, int32_t cols_pad, int32_t cols_valid,
                    int32_t group_rows, int32_t rows_valid, float scale, int32_t dtype, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(P && dP && dS, "csn_softmax_bwd: null pointer");
  CSN_CHECK_ARG(cols_pad % 8 == 0 && cols_valid <= cols_pad, "csn_softmax_bwd: bad widths");
  if (rows == 0) return 0;
  const int wpb = 8;
  softmax_bwd_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(P, dP, dS, rows, cols_pad, cols_valid, group_rows, rows_valid, scale, dtype);
  CSN_LAUNCH_OK("softmax_bwd_kernel");
  return 0;
}

int csn_add_ln_fwd(float* Z, const float* R, const int32_t* res_block, float* Y, void* Y16, float* mean, float* rstd,
                   const float* gamma, const float* beta, float* colsum, int64_t rows, int32_t block_rows,
                   int32_t group_rows, int32_t rows_valid, float eps, int32_t dtype, uint32_t drop_seed, float drop_p,
                   void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Z && R && mean && rstd && gamma && beta, "csn_add_ln_fwd: null pointer");
  CSN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "csn_add_ln_fwd: dropout probability outside [0, 1)");
  CSN_CHECK_ARG(block_rows % 64 == 0 && rows % 64 == 0, "csn_add_ln_fwd: rows (%lld) and block_rows (%d) must be multiples of 64", (long long)rows, block_rows);
  if (rows == 0) return 0;
  const uint32_t dth = drop_thresh16(drop_p);
  AddLnArgs a{Z, R, res_block, Y, Y16, mean, rstd, gamma, beta, colsum, rows, block_rows, group_rows, rows_valid, eps, dtype,
              drop_seed, dth, drop_scale_of(dth), dth ? drop_epoch_ptr() : nullptr};
  return launch_simple(add_ln_fwd_kernel, dim3((unsigned)(rows / 64)), dim3(256), a, stream, "add_ln_fwd_kernel");
}

int csn_ln_colsum(const float* Z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  float* colsum, int64_t rows, int32_t block_rows, int32_t group_rows, int32_t rows_valid, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Z && mean && rstd && gamma && beta && colsum, "csn_ln_colsum: null pointer");
  CSN_CHECK_ARG(block_rows % 64 == 0 && rows % 64 == 0, "csn_ln_colsum: rows (%lld) and block_rows (%d) must be multiples of 64", (long long)rows, block_rows);
  if (rows == 0) return 0;
  ln_colsum_kernel<<<(unsigned)(rows / 64), 256, 0, (cudaStream_t)stream>>>(Z, mean, rstd, gamma, beta, colsum, block_rows, group_rows, rows_valid);
  CSN_LAUNCH_OK("ln_colsum_kernel");
  return 0;
}

int csn_colsum_reduce(const float* part, float* out, int32_t n_blocks, int32_t parts, float scale, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(part && out, "csn_colsum_reduce: null pointer");
  if (n_blocks == 0) return 0;
  colsum_reduce_kernel<<<n_blocks, DM, 0, (cudaStream_t)stream>>>(part, out, parts, scale);
  CSN_LAUNCH_OK("colsum_reduce_kernel");
  return 0;
}

int csn_ln_bwd(const float* dY, const float* Z, const float* mean, const float* rstd, const float* gamma, float* dZ,
               void* dZ16, float* dgamma, float* dbeta, int64_t rows, int32_t block_rows, int32_t group_rows,
               int32_t rows_valid, int32_t dtype, const float* amax, const float* bcast, const int32_t* bcast_idx,
               float bcast_scale, const int32_t* src_idx, const float* src_w, float* chunk_gsum, uint32_t drop_seed,
               float drop_p, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(dY && Z && mean && rstd && gamma && dZ16 && dgamma && dbeta, "csn_ln_bwd: null pointer");
  CSN_CHECK_ARG(!chunk_gsum || group_rows % 64 == 0, "csn_ln_bwd: chunk_gsum needs group_rows to be a multiple of 64");
  CSN_CHECK_ARG(!bcast || bcast_idx, "csn_ln_bwd: bcast needs bcast_idx");
  CSN_CHECK_ARG(!src_idx || src_w, "csn_ln_bwd: src_idx needs src_w");
  CSN_CHECK_ARG(rows % 64 == 0 && block_rows % 64 == 0 && group_rows % 64 == 0 && block_rows % group_rows == 0,
                "csn_ln_bwd: rows, block_rows and group_rows must be multiples of 64, block_rows a multiple of group_rows");
  if (rows == 0) return 0;
  LnBwdArgs a{dY, Z, mean, rstd, gamma, dZ, dZ16, dgamma, dbeta, rows, group_rows, rows_valid, block_rows, dtype, amax, bcast, bcast_idx, bcast_scale, src_idx, src_w, getenv("CSN_LN_BWD_DEBUG") ? atoi(getenv("CSN_LN_BWD_DEBUG")) : 0, chunk_gsum,
              drop_seed, drop_thresh16(drop_p), drop_scale_of(drop_thresh16(drop_p)), drop_thresh16(drop_p) ? drop_epoch_ptr() : nullptr};
  CSN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "csn_ln_bwd: dropout probability outside [0, 1)");
  CSN_CHECK_ARG(!(chunk_gsum && drop_p > 0.f), "csn_ln_bwd: chunk_gsum (V centring) is not combined with dropout");
  static const int occ = getenv("CSN_LN_BWD_OCC") ? atoi(getenv("CSN_LN_BWD_OCC")) : 3;   // resident CTAs/SM (tuning knob)
  const dim3 grid((unsigned)(rows / 64)), block(256);
  const bool drop = a.drop_thresh != 0;
  if (chunk_gsum) return launch_simple(ln_bwd_kernel<3, true, false>, grid, block, a, stream, "ln_bwd_kernel");
  if (drop) return launch_simple(ln_bwd_kernel<3, false, true>, grid, block, a, stream, "ln_bwd_kernel");
  if (occ >= 4) return launch_simple(ln_bwd_kernel<4, false, false>, grid, block, a, stream, "ln_bwd_kernel");
  if (occ >= 3) return launch_simple(ln_bwd_kernel<3, false, false>, grid, block, a, stream, "ln_bwd_kernel");
  return launch_simple(ln_bwd_kernel<2, false, false>, grid, block, a, stream, "ln_bwd_kernel");
}

int csn_segment_mean(const float* x, const int64_t* offsets, int32_t n_seg, int32_t n_cols, float* out, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(x && offsets && out, "csn_segment_mean: null pointer");
  CSN_CHECK_ARG(n_cols > 0 && n_seg >= 0, "csn_segment_mean: bad shape");
  if (n_seg == 0) return 0;
  segment_mean_kernel<<<dim3(n_seg, (n_cols + 31) / 32), 1024, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<const long long*>(offsets), n_cols, out);
  CSN_LAUNCH_OK("segment_mean_kernel");
  return 0;
}

int csn_ragged_pad(const float* x, const int64_t* offsets, int32_t n_slots, int32_t n_pad, float* out32, void* out16,
                   int32_t dtype, float* amax, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(x && offsets && (out32 || out16), "csn_ragged_pad: null pointer");
  CSN_CHECK_ARG(!out16 || dtype == CSN_F16 || dtype == CSN_BF16, "csn_ragged_pad: 16-bit output needs dtype f16 / bf16");
  if (n_slots == 0 || n_pad == 0) return 0;
  int gx = (n_pad + 7) / 8;
  const int cap = (8 * num_sms() + n_slots - 1) / n_slots;
  if (gx > cap) gx = cap;
  ragged_pad_kernel<<<dim3((unsigned)gx, (unsigned)n_slots), 256, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<const long long*>(offsets), n_pad, out32, out16, dtype, amax);
  CSN_LAUNCH_OK("ragged_pad_kernel");
  return 0;
}

int csn_block_add(const float* src, const int32_t* dst_block, int32_t n_src, float* dst, int32_t n_dst, int64_t block_elems,
                  const float* amax, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(src && dst_block && dst, "csn_block_add: null pointer");
  CSN_CHECK_ARG(block_elems % 4 == 0, "csn_block_add: block size must be a multiple of 4 elements");
  CSN_CHECK_ARG(n_src >= 0 && n_src <= 4096 && n_dst >= 0, "csn_block_add: bad block counts");
  if (n_dst == 0 || block_elems == 0) return 0;
  const long long n4 = block_elems / 4;
  long long gx = (n4 + 1023) / 1024;   // 4 float4 per thread
  const long long cap = (4LL * num_sms() + n_dst - 1) / n_dst;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  block_add_kernel<<<dim3((unsigned)gx, (unsigned)n_dst), 256, 0, (cudaStream_t)stream>>>(src, dst_block, n_src, dst, block_elems, amax);
  CSN_LAUNCH_OK("block_add_kernel");
  return 0;
}

int csn_grad_unscale(float* x, int64_t n, const float* amax, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(x && amax, "csn_grad_unscale: null pointer");
  if (n == 0) return 0;
  const unsigned grid = (unsigned)((n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592);
  grad_unscale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, amax);
  CSN_LAUNCH_OK("grad_unscale_kernel");
  return 0;
}

int csn_combine_fwd(const float* Y, const int32_t* blk, const float* w, float* out, void* rows16, int32_t n_b,
                    int32_t n_k, int64_t out_b_stride, int64_t out_ch_stride, int32_t n_points, int32_t chunk,
                    int32_t chunk_pad, int32_t rows_pad, int32_t dtype, const float* ln_mean, const float* ln_rstd,
                    const float* ln_gamma, const float* ln_beta, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Y && blk && w && out, "csn_combine_fwd: null pointer");
  CSN_CHECK_ARG(!ln_mean || (ln_rstd && ln_gamma && ln_beta), "csn_combine_fwd: incomplete LayerNorm parameters");
  CSN_CHECK_ARG(chunk_pad % 32 == 0 && rows_pad % chunk_pad == 0, "csn_combine_fwd: bad padding");
  if (n_b == 0) return 0;
  CombineArgs a{Y, blk, w, out, rows16, out_b_stride, out_ch_stride, n_k, n_points, chunk, chunk_pad, rows_pad, dtype, {ln_mean, ln_rstd, ln_gamma, ln_beta}};
  return launch_simple(combine_fwd_kernel, dim3(rows_pad / 32, n_b), dim3(256), a, stream, "combine_fwd_kernel");
}

int csn_block_dot(const float* G, const float* Z, const int32_t* src_idx, const int32_t* out_idx, float* out,
                  int64_t rows, int32_t block_rows, int32_t group_rows, int32_t rows_valid, const float* ln_mean,
                  const float* ln_rstd, const float* ln_gamma, const float* ln_beta, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(G && Z && src_idx && out_idx && out, "csn_block_dot: null pointer");
  CSN_CHECK_ARG(rows % 64 == 0 && block_rows % 64 == 0, "csn_block_dot: rows and block_rows must be multiples of 64");
  if (rows == 0) return 0;
  BlockDotArgs a{G, Z, src_idx, out_idx, out, rows, block_rows, group_rows, rows_valid, {ln_mean, ln_rstd, ln_gamma, ln_beta}};
  return launch_simple(block_dot_kernel, dim3((unsigned)(rows / 64)), dim3(256), a, stream, "block_dot_kernel");
}

int csn_combine_bwd(const float* dOut, const float* Y, const float* dpool, const int32_t* cb, const float* cw,
                    const int32_t* cw_index, const int32_t* pb, float pool_scale, float* dY, float* dcomp,
                    int32_t n_blocks, int64_t out_b_stride, int64_t out_ch_stride, int32_t n_points, int32_t chunk,
                    int32_t chunk_pad, int32_t rows_pad, float* amax, const float* ln_mean, const float* ln_rstd,
                    const float* ln_gamma, const float* ln_beta, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(dOut && Y && cb && cw && cw_index && pb && (dY || dcomp), "csn_combine_bwd: null pointer");
  if (n_blocks == 0) return 0;
  CombineBwdArgs a{dOut, Y, dpool, cb, cw, cw_index, pb, pool_scale, dY, dcomp, out_b_stride, out_ch_stride, n_points, chunk, chunk_pad, rows_pad, amax, {ln_mean, ln_rstd, ln_gamma, ln_beta}};
  return launch_simple(combine_bwd_kernel, dim3(rows_pad / 32, n_blocks), dim3(256), a, stream, "combine_bwd_kernel");
}

}  // extern "C"
