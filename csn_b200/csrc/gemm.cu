// csn_gemm: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[b][m][n] = alpha * sum_k A[b][m][k] * B[b][n][k]
//
// Roles inside one 256-thread CTA (one CTA per SM, grid = min(#tiles, #SMs)):
//   warp 0   : TMA producer  — fills a ring of {A 128x64, B BNx64} 16-bit stages (SWIZZLE_128B)
//   warp 1   : MMA issuer    — one lane issues tcgen05.mma kind::f16 (M=128, N=BN, K=16) x4 per
//                              stage, accumulating fp32 in TMEM; tcgen05.commit releases the stage
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      — tcgen05.ld (lane = output row), scale, convert, store / atomic add
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.
//
// Operands may be K-major (rows of k) or MN-major (rows of mn, i.e. a transposed operand consumed
// in place): the difference is confined to the TMA box shape and the UMMA descriptor strides.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace csn {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 x 16-bit = 128 B = one swizzle atom row
constexpr int GEMM_THREADS = 384;   // warps 0-3: producer / MMA / TMEM alloc / idle, warps 4-11: epilogue

struct GemmArgs {
  int M, N, K;
  int nb0, nb1, nb2, nb3;
  int tiles_m, tiles_n, split_k, kb_per_split, kb_total;
  long long total_tiles;
  long long a_mn_off[4], a_k_off[4], b_mn_off[4], b_k_off[4];
  void* D;
  long long ldd;
  long long d_off[4];
  int out_dtype, transposed, accumulate, vec_ok;
  float alpha;
  uint32_t idesc;
  // TMA-store epilogue (row-major, non-accumulating outputs): coordinates of a batch's origin
  int tma_store;
  int d_row_off[4], d_col_off[4];
  int stages;     // smem ring depth actually used (<= Cfg::STAGES)
  int stg_bufs;   // epilogue staging slabs per warp
  int tempty_count;  // arrivals that free an accumulator buffer
  int alt_tiles;     // LN epilogue with 8 warps: warps 4-7 drain accumulator 0 (even tiles), warps 8-11 accumulator 1 (odd tiles)
  int epi_warps;  // 4, or 8 for short-K problems where the epilogue is the critical path (two warps per TMEM lane quadrant, half the columns each)
  // fused residual + LayerNorm-statistics epilogue (csn_gemm_res_ln): z = alpha*acc + residual, Z = z, mean/rstd per row
  int ln;
  const int* res_sel;         // per block: which of the two channel-major residual tensors (tensor maps tmR0 / tmR1)
  const int* res_row;         // per block: first channel row of the block's [256][n_points] matrix in that tensor's 2-D view
  int block_rows, group_rows, rows_valid, n_points;
  float eps;
  float* mean; float* rstd;
  // fused delta epilogue (csn_gemm_delta): D = dO (16-bit), delta[(blk*n_head + head)*rows_pad + r] = rowsum_head(dO o (O + O_lo/scale))
  int dl;
  float* delta;
  int dl_rows_pad, dl_n_head, dl_d_head;
  float dl_lo_inv;   // 1/2048 (fp16) or 1/256 (bf16); 0 when there is no O_lo
  int debug;   // CSN_GEMM_DEBUG (diagnostics only): 1 = epilogue drains TMEM but stores nothing
  // column bias per row group (csn_gemm_colbias): D[row][col] -= cbias[(row / cb_group)*cb_ld + col - cb_col0] for
  // col >= cb_col0 and row % cb_group < cb_valid (TMA-store epilogue only).  V is centred on its per-chunk key mean.
  const float* cbias; long long cb_ld; int cb_col0, cb_group, cb_valid;
  // csn_gemm_res_ln: optional per-chunk row vector added to z, zbias[(row / group_rows)*256 + col]
  const float* zbias;
  // csn_gemm_res_ln: dropout on the projection output before the residual add (csa_models.py:115), off when thresh == 0
  uint32_t drop_seed, drop_thresh; float drop_scale;
  const uint32_t* drop_epoch;
  // csn_gemm_dual: a second problem of the same shape whose tiles are interleaved with the first one's, batch by batch
  // (problem 1: A consumed MN-major through tensor map slot tmR0, B through tmR1): the two problems of the attention
  // backward that read the same dS tile — dQ = dS K and dK = dS^T Q — run back to back, so the second read hits L2
  int dual;
  long long a2_mn_off[4], a2_k_off[4], b2_mn_off[4], b2_k_off[4];
  int d2_row_off[4], d2_col_off[4], d2_col_base;
  uint32_t idesc2;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // 128 / 256 / 512: powers of two
  static constexpr int BAR_BYTES = 512;
  static constexpr int STG_BYTES = 4 * 2 * 4096;  // epilogue staging: 4 warps x 2 buffers x (32 rows x 128 B)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;  // +1024: alignment slack
};

struct TileCoord {
  int b0, b1, b2, b3, mt, nt, ks, prob;
};

// 32-bit arithmetic (the host rejects problems with >= 2^31 tiles); divisions by 1 are skipped, which is the
// common case for the batch dims and split_k.
__device__ __forceinline__ TileCoord decode_tile(long long t64, const GemmArgs& p) {
  TileCoord c;
  uint32_t t = (uint32_t)t64;
  auto step = [&](int n) -> int {
    if (n == 1) return 0;
    const uint32_t q = t / (uint32_t)n;
    const int r = (int)(t - q * (uint32_t)n);
    t = q;
    return r;
  };
  if (p.split_k > 1) {
    // split-K (weight gradients: M x N small, K = all points): the CTAs that run concurrently take the m- / n-tiles
    // of the SAME k range, so every operand k-slice is fetched from HBM once and re-read from L2 by the other tiles
    c.mt = step(p.tiles_m);
    c.nt = step(p.tiles_n);
    c.ks = step(p.split_k);
  } else {
    c.ks = 0;
    c.nt = step(p.tiles_n);
    c.mt = step(p.tiles_m);
  }
  c.prob = p.dual ? step(2) : 0;   // dual launch: [all tiles of problem 0][all tiles of problem 1] of one batch, then the next batch
  c.b0 = step(p.nb0);
  c.b1 = step(p.nb1);
  c.b2 = step(p.nb2);
  c.b3 = (int)t;
  return c;
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int dtype) {
  if (dtype == CSN_F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// Store 32 consecutive columns [n, n+32) of output row `row` held in v[] (already scaled).
__device__ __forceinline__ void store_chunk(const GemmArgs& p, long long base, int row, int n,
                                            const float (&v)[32]) {
  const int ncols = min(32, p.N - n);
  if (!p.transposed) {
    const long long o = base + (long long)row * p.ldd + n;
    if (p.accumulate) {
      float* d = reinterpret_cast<float*>(p.D) + o;
      for (int j = 0; j < ncols; ++j) atomicAdd(d + j, v[j]);
    } else if (p.out_dtype == CSN_F32) {
      float* d = reinterpret_cast<float*>(p.D) + o;
      if (ncols == 32 && p.vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < ncols; ++j) d[j] = v[j];
      }
    } else {
      uint16_t* d = reinterpret_cast<uint16_t*>(p.D) + o;
      if (ncols == 32 && p.vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 q;
          q.x = pack16(v[j], v[j + 1], p.out_dtype);
          q.y = pack16(v[j + 2], v[j + 3], p.out_dtype);
          q.z = pack16(v[j + 4], v[j + 5], p.out_dtype);
          q.w = pack16(v[j + 6], v[j + 7], p.out_dtype);
          *reinterpret_cast<uint4*>(d + j) = q;
        }
      } else {
        for (int j = 0; j < ncols; ++j) {
          uint32_t w = pack16(v[j], 0.f, p.out_dtype);
          d[j] = (uint16_t)(w & 0xFFFFu);
        }
      }
    }
  } else {
    // D[n*ld + m]: for a fixed column the 32 lanes of the warp hold 32 consecutive m -> coalesced.
    const long long o = base + (long long)n * p.ldd + row;
    if (p.accumulate) {
      float* d = reinterpret_cast<float*>(p.D) + o;
      for (int j = 0; j < ncols; ++j) atomicAdd(d + (long long)j * p.ldd, v[j]);
    } else if (p.out_dtype == CSN_F32) {
      float* d = reinterpret_cast<float*>(p.D) + o;
      for (int j = 0; j < ncols; ++j) d[(long long)j * p.ldd] = v[j];
    } else {
      uint16_t* d = reinterpret_cast<uint16_t*>(p.D) + o;
      for (int j = 0; j < ncols; ++j) {
        uint32_t w = pack16(v[j], 0.f, p.out_dtype);
        d[(long long)j * p.ldd] = (uint16_t)(w & 0xFFFFu);
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int CL, bool LN = false, bool DL = false, bool CB = false, bool DROP = false, bool DUAL = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR0,
            const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  // [ring: p.stages x STAGE_BYTES][staging: 4 warps x p.stg_bufs x 4 KB][barriers]; the host trades ring depth for
  // staging depth (same total) when K is short and the TMA-store epilogue is the critical path
  const int NST = p.stages;
  const uint32_t stg_base = smem_base + NST * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + p.epi_warps * p.stg_bufs * 4096;
  // barrier layout: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * NST + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * NST + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bar_base - smem_base) + 8 * (2 * NST + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);   // CL == 2: the B half-tiles are multicast, both CTAs must have drained the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), p.tempty_count);
    }
    if (LN || DL) for (int i = 0; i < 16; ++i) mbar_init(bar_base + 256 + 8u * i, 1);   // epilogue operand slabs: (up to) 8 warps x 2
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // CL == 2: the two CTAs of a cluster own m-tiles 2i and 2i+1 of the same (batch, n-tile, k-split); each loads
  // half of every B stage and multicasts it to both, so B crosses L2 -> SM once per pair instead of once per tile.
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const long long t_first = (long long)blockIdx.x / CL, t_stride = (long long)gridDim.x / CL;
  constexpr uint16_t MC_MASK = (1u << CL) - 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (long long t = t_first; t < p.total_tiles; t += t_stride) {
        TileCoord c = decode_tile(t, p);
        c.mt = c.mt * CL + rank;
        const bool pr = DUAL && c.prob != 0;
        const bool a_mnm = DUAL ? pr : A_MN;      // dual: problem 1 consumes A MN-major (dS^T)
        const long long* amo = pr ? p.a2_mn_off : p.a_mn_off;
        const long long* ako = pr ? p.a2_k_off : p.a_k_off;
        const long long* bmo = pr ? p.b2_mn_off : p.b_mn_off;
        const long long* bko = pr ? p.b2_k_off : p.b_k_off;
        const CUtensorMap* tA = pr ? &tmR0 : &tmA;
        const CUtensorMap* tB = pr ? &tmR1 : &tmB;
        const long long a_mn = c.b0 * amo[0] + c.b1 * amo[1] + c.b2 * amo[2] + c.b3 * amo[3] + (long long)c.mt * GEMM_BM;
        const long long b_mn = c.b0 * bmo[0] + c.b1 * bmo[1] + c.b2 * bmo[2] + c.b3 * bmo[3] + (long long)c.nt * BN;
        const long long a_k0 = c.b0 * ako[0] + c.b1 * ako[1] + c.b2 * ako[2] + c.b3 * ako[3];
        const long long b_k0 = c.b0 * bko[0] + c.b1 * bko[1] + c.b2 * bko[2] + c.b3 * bko[3];
        const int kb0 = c.ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(st), ph ^ 1);
          const uint32_t sA = smem_base + st * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          mbar_arrive_expect_tx(full_bar(st), Cfg::STAGE_BYTES);
          const int ka = (int)(a_k0 + (long long)kb * GEMM_BK);
          const int kbb = (int)(b_k0 + (long long)kb * GEMM_BK);
          if (!a_mnm) {
            tma_load_2d(sA, tA, full_bar(st), ka, (int)a_mn);
          } else {
#pragma unroll
            for (int at = 0; at < GEMM_BM / 64; ++at)
              tma_load_2d(sA + at * 8192, tA, full_bar(st), (int)a_mn + at * 64, ka);
          }
          if (CL == 1) {
            if (!B_MN) {
              tma_load_2d(sB, tB, full_bar(st), kbb, (int)b_mn);
            } else {
#pragma unroll
              for (int at = 0; at < BN / 64; ++at)
                tma_load_2d(sB + at * 8192, tB, full_bar(st), (int)b_mn + at * 64, kbb);
            }
          } else {
            if (!B_MN) {   // this CTA's half of the BN rows (box = BN/2 rows), delivered to both CTAs
              tma_load_2d_mc(sB + rank * (BN / 2) * 128, tB, full_bar(st), kbb, (int)b_mn + rank * (BN / 2), MC_MASK);
            } else {
#pragma unroll
              for (int at = 0; at < BN / 64; ++at)
                if ((at % CL) == rank) tma_load_2d_mc(sB + at * 8192, tB, full_bar(st), (int)b_mn + at * 64, kbb, MC_MASK);
            }
          }
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (long long t = t_first; t < p.total_tiles; t += t_stride) {
        const TileCoord c = decode_tile(t, p);
        const int kb0 = c.ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const bool a_mnm = DUAL ? (c.prob != 0) : A_MN;
        const uint32_t idesc_t = (DUAL && c.prob != 0) ? p.idesc2 : p.idesc;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(st), ph);
          tc_fence_after();
          const uint32_t sA = smem_base + st * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // K-major: 16 elements = 32 B further along the swizzled row.
            // MN-major: 16 K-rows of 128 B further down.
            const uint64_t ad = a_mnm ? umma_desc_sw128(sA + k * 2048, 8192, 1024)
                                      : umma_desc_sw128(sA + k * 32, 0, 1024);
            const uint64_t bd = B_MN ? umma_desc_sw128(sB + k * 2048, 8192, 1024)
                                     : umma_desc_sw128(sB + k * 32, 0, 1024);
            umma_f16_ss(d_tmem, ad, bd, idesc_t, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CL == 1) umma_commit(empty_bar(st)); else umma_commit_mc(empty_bar(st), MC_MASK);  // stage reusable once these MMAs have read it
          if (++st == NST) { st = 0; ph ^= 1; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + p.epi_warps) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;  // with 8 epilogue warps: which half of the tile's columns
    const int ew = warp - 4;
    int stg_flip = 0;
    uint32_t res_ph = 0;   // LN epilogue: phase bits of the two residual-slab barriers
    int it = 0;            // tiles seen so far: tile `it` lives in accumulator it & 1, use (it >> 1) of that buffer
    for (long long t = t_first; t < p.total_tiles; t += t_stride, ++it) {
      if (LN && p.alt_tiles && (it & 1) != half) continue;   // the other warp set drains this accumulator
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      TileCoord c = decode_tile(t, p);
      c.mt = c.mt * CL + rank;
      const int kb0 = c.ks * p.kb_per_split;
      const bool has_k = kb0 < p.kb_total;
      if (DL && lane == 0) {
        // delta epilogue: the first two O / O_lo slabs of this tile are requested before waiting for the accumulator
        const int prow0 = c.mt * GEMM_BM + q * 32, pn0 = c.nt * BN;
        const int pn_units = min(BN / 64, (p.N - pn0 + 63) / 64);
        const uint32_t pbuf = stg_base + ew * 6 * 4096;
        for (int u = 0; u < 2 && u < pn_units; ++u) {
          const uint32_t bar = bar_base + 256 + 8u * (ew * 2 + u);
          mbar_arrive_expect_tx(bar, p.dl_lo_inv != 0.f ? 8192 : 4096);
          tma_load_2d(pbuf + (2 + u) * 4096, &tmR0, bar, pn0 + u * 64, prow0);
          if (p.dl_lo_inv != 0.f) tma_load_2d(pbuf + (4 + u) * 4096, &tmR1, bar, pn0 + u * 64, prow0);
        }
      }
      if ((CB || LN) && lane < 8) {
        // per-row-group bias rows (column bias / zbias): pull the tile's 1 KB row into L1 while the MMAs run
        const int prow = c.mt * GEMM_BM + q * 32;
        const float* src = nullptr;
        if (CB && c.nt * BN >= p.cb_col0) src = p.cbias + (long long)(prow / p.cb_group) * p.cb_ld - p.cb_col0 + c.nt * BN;
        if (LN && p.zbias) src = p.zbias + (long long)(prow / p.group_rows) * BN;
        if (src) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + lane * 32));
      }
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const int row = c.mt * GEMM_BM + q * 32 + lane;
      const int n0 = c.nt * BN;
      const long long base = c.b0 * p.d_off[0] + c.b1 * p.d_off[1] + c.b2 * p.d_off[2] + c.b3 * p.d_off[3];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      if (p.debug & 1) {
      } else if (DL) {
        // dO tile (16-bit out) fused with delta = rowsum_head(dO o O): the thread owns its row; O (and its rounding
        // residual O_lo) arrive as warp-private [32 rows x 64 cols] TMA boxes two 64-column units ahead (issued before
        // the wait for the accumulator, so they load under the MMAs).  Per warp: slabs 0,1 output staging, 2,3 O,
        // 4,5 O_lo.  delta is formed from the ROUNDED dO, the value every later kernel sees.
        const int row0 = c.mt * GEMM_BM + q * 32;
        const uint32_t wbuf = stg_base + ew * 6 * 4096;
        const uint32_t rbar0 = bar_base + 256 + 8u * (ew * 2);
        const bool has_lo = p.dl_lo_inv != 0.f;
        const int n_units = min(BN / 64, (p.N - n0 + 63) / 64);
        auto fetch_o = [&](int u) {   // lane 0 only
          const uint32_t bar = rbar0 + 8u * (u & 1);
          mbar_arrive_expect_tx(bar, has_lo ? 8192 : 4096);
          tma_load_2d(wbuf + (2 + (u & 1)) * 4096, &tmR0, bar, n0 + u * 64, row0);
          if (has_lo) tma_load_2d(wbuf + (4 + (u & 1)) * 4096, &tmR1, bar, n0 + u * 64, row0);
        };
        const int blk = row / p.dl_rows_pad, rin = row - blk * p.dl_rows_pad;
        const bool f16 = p.out_dtype == CSN_F16;
        const float al = p.alpha;
        float dsum = 0.f;
        const int units_per_head = p.dl_d_head / 64;
#pragma unroll 1
        for (int u = 0; u < n_units; ++u) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + u * 64, r0);
          tmem_ld_32x32(taddr + u * 64 + 32, r1);
          tmem_ld_wait();
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            w[j] = pack16(__uint_as_float(r0[2 * j]) * al, __uint_as_float(r0[2 * j + 1]) * al, p.out_dtype);
            w[16 + j] = pack16(__uint_as_float(r1[2 * j]) * al, __uint_as_float(r1[2 * j + 1]) * al, p.out_dtype);
          }
          mbar_wait(rbar0 + 8u * (u & 1), (res_ph >> (u & 1)) & 1u);
          res_ph ^= 1u << (u & 1);
          const uint32_t so = wbuf + (2 + (u & 1)) * 4096 + lane * 128, sl = wbuf + (4 + (u & 1)) * 4096 + lane * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint32_t sw = (((uint32_t)ch ^ ((uint32_t)lane & 7u)) << 4);
            uint32_t o4[4], l4[4] = {0u, 0u, 0u, 0u};
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o4[0]), "=r"(o4[1]), "=r"(o4[2]), "=r"(o4[3]) : "r"(so + sw));
            if (has_lo) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(l4[0]), "=r"(l4[1]), "=r"(l4[2]), "=r"(l4[3]) : "r"(sl + sw));
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              float2 dv, ov, lv;
              const uint32_t dw = w[ch * 4 + t2];
              if (f16) {
                dv = __half22float2(*reinterpret_cast<const __half2*>(&dw));
                ov = __half22float2(*reinterpret_cast<const __half2*>(&o4[t2]));
                lv = __half22float2(*reinterpret_cast<const __half2*>(&l4[t2]));
              } else {
                dv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dw));
                ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o4[t2]));
                lv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l4[t2]));
              }
              dsum += dv.x * (ov.x + lv.x * p.dl_lo_inv) + dv.y * (ov.y + lv.y * p.dl_lo_inv);
            }
          }
          __syncwarp();
          if (lane == 0 && u + 2 < n_units) fetch_o(u + 2);
          if ((u + 1) % units_per_head == 0) {
            const int head = (n0 + u * 64) / p.dl_d_head;
            if (row < p.M) p.delta[((long long)blk * p.dl_n_head + head) * p.dl_rows_pad + rin] = dsum;
            dsum = 0.f;
          }
          const uint32_t buf = wbuf + stg_flip * 4096;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint32_t a = rowaddr + (((uint32_t)ch ^ ((uint32_t)lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * ch]), "r"(w[4 * ch + 1]), "r"(w[4 * ch + 2]), "r"(w[4 * ch + 3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            tma_store_2d(&tmD, buf, n0 + u * 64, row0);
            tma_store_commit();
          }
          stg_flip ^= 1;
        }
      } else if (LN) {
        // Fused residual + LayerNorm statistics (BN == N == 256: the thread owns its whole output row).
        // z = alpha*acc + residual.  The residual comes from the reference's channel-major fp32 tensors: a
        // [32 channels x 32 points] TMA box per 32-column unit lands in a warp-private slab two units ahead of
        // its use (for a fixed channel the 32 lanes read 32 consecutive floats: conflict-free).  Z leaves through
        // the slab/TMA-store path; the row's mean and 1/std come from per-unit (mean, M2) pairs merged with
        // Chan's formula.  Per warp: slabs 0,1 = output staging, slabs 2,3 = residual.
        const int row0 = c.mt * GEMM_BM + q * 32;
        const int last_blk = (p.M - 1) / p.block_rows;
        const int blk0 = min(row0 / p.block_rows, last_blk), rin0 = row0 - blk0 * p.block_rows;   // 32 | block_rows
        const int grp = rin0 / p.group_rows, j0 = rin0 - grp * p.group_rows;
        const int pt0 = grp * p.rows_valid + j0;
        const bool valid = row < p.M && (j0 + lane) < p.rows_valid && (pt0 + lane) < p.n_points;
        const int sel = __ldg(p.res_sel + blk0), rrow = __ldg(p.res_row + blk0);
        const CUtensorMap* tmR = sel ? &tmR1 : &tmR0;
        const uint32_t wbuf = stg_base + ew * 4 * 4096;
        const uint32_t rbar0 = bar_base + 256 + 8u * (ew * 2);
        const float al = p.alpha;
        auto fetch_res = [&](int u) {   // lane 0 only
          const uint32_t bar = rbar0 + 8u * (u & 1);
          mbar_arrive_expect_tx(bar, 4096);
          tma_load_2d(wbuf + (2 + (u & 1)) * 4096, tmR, bar, pt0, rrow + u * 32);
        };
        if (lane == 0) { fetch_res(0); fetch_res(1); }
        float mu = 0.f, m2 = 0.f;
        uint32_t ra[32], rb[32];
        const uint32_t drop_rk = drop_row_key(drop_seed_eff(p.drop_seed, p.drop_epoch), (uint32_t)row);
        auto emit = [&](uint32_t (&r)[32], int u) {
          // residual slab of this unit
          mbar_wait(rbar0 + 8u * (u & 1), (res_ph >> (u & 1)) & 1u);
          res_ph ^= 1u << (u & 1);
          const uint32_t rs = wbuf + (2 + (u & 1)) * 4096 + (lane & 3) * 4;
          float z[32];
          float su = 0.f;
          if (p.zbias) {   // warp-uniform row (the warp's 32 rows lie in one chunk): broadcast loads
            const float4* zb4 = reinterpret_cast<const float4*>(p.zbias + (long long)(row0 / p.group_rows) * BN + u * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bv = __ldg(zb4 + j4);
              z[4 * j4] = bv.x; z[4 * j4 + 1] = bv.y; z[4 * j4 + 2] = bv.z; z[4 * j4 + 3] = bv.w;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) z[jj] = 0.f;
          }
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            float rv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rv) : "r"(rs + jj * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)jj & 7u)) << 4)));
            float fc = __uint_as_float(r[jj]) * al;
            if (DROP) {
              const uint32_t hh = drop_pair(drop_rk, (uint32_t)(u * 32 + jj) >> 1);   // (compiler shares the hash of a pair)
              fc = ((jj & 1) ? drop_keep_hi(hh, p.drop_thresh) : drop_keep_lo(hh, p.drop_thresh)) ? fc * p.drop_scale : 0.f;
            }
            z[jj] = valid ? fc + rv + z[jj] : 0.f;
            su += z[jj];
          }
          __syncwarp();
          if (lane == 0 && u + 2 < BN / 32) fetch_res(u + 2);   // the slab is free again: prefetch two units ahead
          const float mu_u = su * (1.f / 32.f);
          float q2 = 0.f;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) { const float dlt = z[jj] - mu_u; q2 += dlt * dlt; }
          {   // Chan merge of (32u, mu, m2) with (32, mu_u, q2)
            const float w = 1.f / (float)(u + 1), dlt = mu_u - mu;
            mu += dlt * w;
            m2 += q2 + dlt * dlt * (32.f * (float)u * w);
          }
          const uint32_t buf = wbuf + stg_flip * 4096;
          const uint32_t rowaddr = buf + lane * 128;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint32_t a = rowaddr + (((uint32_t)ch ^ ((uint32_t)lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__float_as_uint(z[4 * ch])), "r"(__float_as_uint(z[4 * ch + 1])),
                         "r"(__float_as_uint(z[4 * ch + 2])), "r"(__float_as_uint(z[4 * ch + 3])) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            tma_store_2d(&tmD, buf, u * 32, row0);
            tma_store_commit();
          }
          stg_flip ^= 1;
        };
        constexpr int NU = BN / 32;
        tmem_ld_32x32(taddr, ra);
#pragma unroll 1
        for (int u = 0; u < NU; u += 2) {
          tmem_ld_wait();
          tmem_ld_32x32(taddr + (u + 1) * 32, rb);
          emit(ra, u);
          tmem_ld_wait();
          if (u + 2 < NU) tmem_ld_32x32(taddr + (u + 2) * 32, ra);
          emit(rb, u + 1);
        }
        if (row < p.M) {
          p.mean[row] = valid ? mu : 0.f;
          p.rstd[row] = valid ? rsqrtf(m2 * (1.f / (float)BN) + p.eps) : 0.f;
        }
      } else if (p.tma_store) {
        // Coalesced path: each warp stages [32 rows x 128 B] slabs (128B-swizzled, conflict-free 16-byte
        // writes) and hands them to the TMA store engine. Work unit = one x32 TMEM load (32 fp32 columns):
        // a whole slab for fp32 outputs, half a slab for 16-bit outputs. The load of unit u+1 is in flight
        // while unit u is converted and staged (tcgen05.ld latency is several hundred cycles while the MMA
        // pipe is writing the other accumulator).
        const bool o32 = p.out_dtype == CSN_F32;
        const int W = o32 ? 32 : 64;  // columns per 128-byte slab
        const int* dro = (DUAL && c.prob != 0) ? p.d2_row_off : p.d_row_off;
        const int* dco = (DUAL && c.prob != 0) ? p.d2_col_off : p.d_col_off;
        const int row0 = c.b0 * dro[0] + c.b1 * dro[1] + c.b2 * dro[2] + c.b3 * dro[3] + c.mt * GEMM_BM + q * 32;
        const int col0 = c.b0 * dco[0] + c.b1 * dco[1] + c.b2 * dco[2] + c.b3 * dco[3] + n0 + ((DUAL && c.prob != 0) ? p.d2_col_base : 0);
        const bool rows_live = c.mt * GEMM_BM + q * 32 < p.M;  // warp-uniform
        // this warp's unit range [u0, u1): all live units, or one half of them (split on a slab boundary)
        const int g = o32 ? 1 : 2;                                          // units per slab
        const int n_slabs = min(BN / W, (p.N - n0 + W - 1) / W);
        const int s_half = p.epi_warps == 8 ? (n_slabs + 1) / 2 : n_slabs;
        const int u0 = half * s_half * g, u1 = min(n_slabs, (half + 1) * s_half) * g;
        const float al = p.alpha;
        uint32_t ra[32], rb[32];
        // column bias (CB): the 32 rows of this warp lie in one row group, so the bias row is warp-uniform (broadcast
        // loads that hit L1: the row's lines are prefetched before the wait for the accumulator)
        const bool cb_rows = CB && ((row0 + lane) % p.cb_group) < p.cb_valid;
        const float* cb_row = CB ? p.cbias + (long long)(row0 / p.cb_group) * p.cb_ld - p.cb_col0 : nullptr;
        const bool cb_tile = CB && has_k && n0 >= p.cb_col0;
        auto emit = [&](const uint32_t (&r)[32], int u) {
          float bv[32];
          if (CB) {
            if (cb_tile) {
              const float4* b4 = reinterpret_cast<const float4*>(cb_row + n0 + u * 32);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 t4 = __ldg(b4 + j4);
                bv[4 * j4] = cb_rows ? t4.x : 0.f; bv[4 * j4 + 1] = cb_rows ? t4.y : 0.f;
                bv[4 * j4 + 2] = cb_rows ? t4.z : 0.f; bv[4 * j4 + 3] = cb_rows ? t4.w : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) bv[j] = 0.f;
            }
          }
          if (p.debug & 8) { if (r[0] == 0x7fc12345u && r[31] == 0x7fc12345u) atomicAdd(reinterpret_cast<int*>(p.D), 1); return; }   // diagnostics: TMEM loads only
          const uint32_t buf = stg_base + (ew * p.stg_bufs + stg_flip) * 4096;
          const uint32_t rowaddr = buf + lane * 128;
          const bool first = o32 || !(u & 1), last = o32 || (u & 1);
          if (first) {   // the slab buffer must have been drained by the store issued stg_bufs slabs ago
            if (lane == 0) { if (p.stg_bufs == 4) tma_store_wait_read<3>(); else tma_store_wait_read<1>(); }
            __syncwarp();
          }
          if (o32) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint32_t a = rowaddr + (((uint32_t)ch ^ ((uint32_t)lane & 7u)) << 4);
              uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
              if (has_k) {
                w0 = __float_as_uint(__uint_as_float(r[4 * ch]) * al - (CB ? bv[4 * ch] : 0.f));
                w1 = __float_as_uint(__uint_as_float(r[4 * ch + 1]) * al - (CB ? bv[4 * ch + 1] : 0.f));
                w2 = __float_as_uint(__uint_as_float(r[4 * ch + 2]) * al - (CB ? bv[4 * ch + 2] : 0.f));
                w3 = __float_as_uint(__uint_as_float(r[4 * ch + 3]) * al - (CB ? bv[4 * ch + 3] : 0.f));
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            }
          } else {
            const bool f16 = p.out_dtype == CSN_F16;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t w[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = has_k ? __uint_as_float(r[8 * ch + 2 * j]) * al - (CB ? bv[8 * ch + 2 * j] : 0.f) : 0.f;
                const float y = has_k ? __uint_as_float(r[8 * ch + 2 * j + 1]) * al - (CB ? bv[8 * ch + 2 * j + 1] : 0.f) : 0.f;
                if (f16) { __half2 h = __floats2half2_rn(x, y); w[j] = *reinterpret_cast<uint32_t*>(&h); }
                else { __nv_bfloat162 h = __floats2bfloat162_rn(x, y); w[j] = *reinterpret_cast<uint32_t*>(&h); }
              }
              const uint32_t a = rowaddr + (((uint32_t)(ch + 4 * (u & 1)) ^ ((uint32_t)lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
          }
          if (last) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && rows_live && !(p.debug & 4)) {
              const int sl = u / g;
              if (p.accumulate) { if (has_k) tma_reduce_add_2d(&tmD, buf, col0 + sl * W, row0); }   // split-K partial: += in L2
              else tma_store_2d(&tmD, buf, col0 + sl * W, row0);
              tma_store_commit();
            }
            if (++stg_flip == p.stg_bufs) stg_flip = 0;
          }
        };
        if (p.debug & 16) {   // diagnostics: staging + stores without touching TMEM
#pragma unroll
          for (int j = 0; j < 32; ++j) ra[j] = 0x3c003c00u + j;
#pragma unroll 1
          for (int u = u0; u < u1; ++u) emit(ra, u);
        } else {
        if (u0 < u1) tmem_ld_32x32(taddr + u0 * 32, ra);
#pragma unroll 1
        for (int u = u0; u < u1; u += 2) {
          tmem_ld_wait();
          if (u + 1 < u1) tmem_ld_32x32(taddr + (u + 1) * 32, rb);
          emit(ra, u);
          if (u + 1 < u1) {
            tmem_ld_wait();
            if (u + 2 < u1) tmem_ld_32x32(taddr + (u + 2) * 32, ra);
            emit(rb, u + 1);
          }
        }
        }
      } else
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        if (n0 + cc >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(taddr + cc, r);
        tmem_ld_wait();
        if (row < p.M) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = has_k ? __uint_as_float(r[j]) * p.alpha : 0.f;
          store_chunk(p, base, row, n0 + cc, v);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();  // staged slabs must be read before the CTA exits
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN, int CL, bool LN = false, bool DL = false, bool CB = false, bool DROP = false, bool DUAL = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                       const GemmArgs& args, cudaStream_t stream, const CUtensorMap* tmR0 = nullptr,
                       const CUtensorMap* tmR1 = nullptr) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, CL, LN, DL, CB, DROP, DUAL>;
  CSN_SET_MAX_SMEM(kern, Cfg::SMEM_BYTES);
  const long long workers = num_sms() / CL;
  const long long grid = (args.total_tiles < workers ? args.total_tiles : workers) * CL;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, tmR0 ? *tmR0 : tmA, tmR1 ? *tmR1 : tmA, args));
  CSN_LAUNCH_OK("gemm_kernel");
  return 0;
}

template <int BN, int CL>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const CUtensorMap& tmD, const GemmArgs& args, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, CL>(tmA, tmB, tmD, args, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, CL>(tmA, tmB, tmD, args, stream);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false, CL>(tmA, tmB, tmD, args, stream);
  return launch_gemm<BN, true, true, CL>(tmA, tmB, tmD, args, stream);
}

struct LnEpilogue {
  const float* res0; long long res0_rows; const float* res1; long long res1_rows; const int* res_sel; const int* res_row;
  long long res_ld; int block_rows, group_rows, rows_valid, n_points; float eps; float* mean; float* rstd;
  const float* zbias; uint32_t drop_seed; float drop_p;
};

struct ColBias {
  const float* bias; long long ld; int col0, group_rows, rows_valid;
};

struct DualProblem {   // the second problem of csn_gemm_dual (same M, N, K and batch extents as the first)
  const csn_mat* A; const csn_mat* B; const csn_out* D;
};

struct DeltaEpilogue {
  const void* O; const void* O_lo; long long ldo; long long o_rows; float* delta; int rows_pad, n_head, d_head;
};

}  // namespace csn

static int gemm_impl(const csn_mat* A, const csn_mat* B, const csn_out* D, int32_t M, int32_t N,
                     int32_t K, const int32_t nb[4], float alpha, int32_t split_k, void* stream,
                     const csn::LnEpilogue* ln, const csn::DeltaEpilogue* dl = nullptr, const csn::ColBias* cb = nullptr,
                     const csn::DualProblem* du = nullptr) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(A && B && D && nb, "csn_gemm: null argument");
  CSN_CHECK_ARG(M > 0 && N > 0 && K > 0, "csn_gemm: empty problem M=%d N=%d K=%d", M, N, K);
  CSN_CHECK_ARG(A->dtype == B->dtype, "csn_gemm: operand dtypes differ");
  CSN_CHECK_ARG(A->dtype == CSN_F16 || A->dtype == CSN_BF16, "csn_gemm: operands must be f16/bf16");
  CSN_CHECK_ARG(nb[0] > 0 && nb[1] > 0 && nb[2] > 0 && nb[3] > 0, "csn_gemm: batch extents must be >= 1");
  CSN_CHECK_ARG(split_k >= 1, "csn_gemm: split_k must be >= 1");
  CSN_CHECK_ARG(split_k == 1 || D->accumulate, "csn_gemm: split_k > 1 requires accumulate");
  CSN_CHECK_ARG(!D->accumulate || D->dtype == CSN_F32, "csn_gemm: accumulate requires fp32 output");
  CSN_CHECK_ARG(D->ptr != nullptr && A->ptr != nullptr && B->ptr != nullptr, "csn_gemm: null data pointer");

  const int BN = (N > 128) ? 256 : (N > 64 ? 128 : 64);
  const bool a_mn = A->major == CSN_MAJOR_MN, b_mn = B->major == CSN_MAJOR_MN;
  const int tiles_m_all = (M + GEMM_BM - 1) / GEMM_BM;
  static const int debug_flags = getenv("CSN_GEMM_DEBUG") ? atoi(getenv("CSN_GEMM_DEBUG")) : 0;
  static const bool cluster_ok = getenv("CSN_GEMM_CLUSTER") == nullptr || atoi(getenv("CSN_GEMM_CLUSTER")) != 0;
  // pairs of m-tiles share their B tiles through a 2-CTA cluster with multicast loads
  const int CL = (cluster_ok && BN >= 128 && tiles_m_all % 2 == 0) ? 2 : 1;

  CUtensorMap tmA, tmB;
  int rc;
  rc = a_mn ? make_tmap_2d(&tmA, A->ptr, A->dtype, A->inner, A->outer, A->ld, 64, 64)
            : make_tmap_2d(&tmA, A->ptr, A->dtype, A->inner, A->outer, A->ld, 64, GEMM_BM);
  if (rc) return rc;
  rc = b_mn ? make_tmap_2d(&tmB, B->ptr, B->dtype, B->inner, B->outer, B->ld, 64, 64)
            : make_tmap_2d(&tmB, B->ptr, B->dtype, B->inner, B->outer, B->ld, 64, (uint32_t)(BN / CL));
  if (rc) return rc;

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K;
  g.debug = debug_flags;
  g.nb0 = nb[0]; g.nb1 = nb[1]; g.nb2 = nb[2]; g.nb3 = nb[3];
  g.tiles_m = tiles_m_all / CL;   // decode space: m-tile PAIRS when CL == 2
  g.tiles_n = (N + BN - 1) / BN;
  g.kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  g.split_k = split_k > g.kb_total ? g.kb_total : split_k;
  g.kb_per_split = (g.kb_total + g.split_k - 1) / g.split_k;
  g.split_k = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;  // no empty splits
  g.total_tiles = (long long)g.nb0 * g.nb1 * g.nb2 * g.nb3 * g.tiles_m * g.tiles_n * g.split_k;
  CSN_CHECK_ARG(g.total_tiles < (1ll << 31), "csn_gemm: too many tiles (%lld)", g.total_tiles);
  for (int i = 0; i < 4; ++i) {
    g.a_mn_off[i] = A->mn_off[i]; g.a_k_off[i] = A->k_off[i];
    g.b_mn_off[i] = B->mn_off[i]; g.b_k_off[i] = B->k_off[i];
    g.d_off[i] = D->off[i];
  }
  g.D = D->ptr; g.ldd = D->ld; g.out_dtype = D->dtype; g.transposed = D->transposed;
  g.accumulate = D->accumulate; g.alpha = alpha;
  const int esz = D->dtype == CSN_F32 ? 4 : 2;
  bool vec = ((reinterpret_cast<uintptr_t>(D->ptr) & 15) == 0) && ((D->ld * esz) % 16 == 0);
  for (int i = 0; i < 4; ++i) vec = vec && ((D->off[i] * esz) % 16 == 0);
  g.vec_ok = vec ? 1 : 0;
  g.idesc = umma_idesc_f16(A->dtype == CSN_F16 ? 0u : 1u, a_mn ? 1u : 0u, b_mn ? 1u : 0u, (uint32_t)BN);

  // TMA-store epilogue when the output is a plain row-major tile grid: every batch origin must be a
  // (row, col) coordinate of one 2-D view with leading dimension ld, and tiles must not spill into a
  // neighbouring batch (M multiple of 128 / N multiple of the slab width, or a single batch).
  CUtensorMap tmD = tmA;
  g.tma_store = 0;
  {
    const int W = D->dtype == CSN_F32 ? 32 : 64;
    const long long nbt = (long long)nb[0] * nb[1] * nb[2] * nb[3];
    bool ok = (!D->accumulate || D->dtype == CSN_F32) && !D->transposed && vec && D->ld >= N;
    ok = ok && (nbt == 1 || (M % GEMM_BM == 0 && N % W == 0));
    long long max_row = M, max_col = N;
    for (int i = 0; i < 4 && ok; ++i) {
      const long long ro = D->off[i] / D->ld, co = D->off[i] % D->ld;
      if (D->off[i] < 0 || ro > 0x3fffffff) ok = false;
      g.d_row_off[i] = (int)ro;
      g.d_col_off[i] = (int)co;
      max_row += ro * (nb[i] - 1);
      max_col += co * (nb[i] - 1);
    }
    ok = ok && max_col <= D->ld;
    if (ok) {
      rc = make_tmap_2d_any(&tmD, D->ptr, D->dtype, nbt == 1 ? N : D->ld, max_row, D->ld, (uint32_t)W, 32);
      if (rc) return rc;
      g.tma_store = 1;
    }
  }
  // smem split: deep ring for long K, deep epilogue staging (4 slabs in flight per warp) for short K
  {
    const int full = BN == 256 ? 4 : (BN == 128 ? 6 : 8), shallow = BN == 256 ? 3 : (BN == 128 ? 5 : 6);
    const bool deep_staging = g.tma_store && g.kb_per_split <= 16;
    g.stages = deep_staging ? shallow : full;
    g.stg_bufs = 2;
    g.epi_warps = deep_staging ? 8 : 4;
    g.tempty_count = 32 * g.epi_warps;
  }
  if (cb) {
    CSN_CHECK_ARG(g.tma_store && !D->accumulate, "csn_gemm_colbias: needs a row-major, 16-byte aligned, non-accumulating output");
    CSN_CHECK_ARG(cb->group_rows % 32 == 0 && cb->col0 % 32 == 0 && cb->ld % 4 == 0 && (reinterpret_cast<uintptr_t>(cb->bias) & 15) == 0,
                  "csn_gemm_colbias: group_rows and col0 must be multiples of 32, the bias rows 16-byte aligned");
    CSN_CHECK_ARG((long long)nb[0] * nb[1] * nb[2] * nb[3] == 1, "csn_gemm_colbias: one batch");
    g.cbias = cb->bias; g.cb_ld = cb->ld; g.cb_col0 = cb->col0; g.cb_group = cb->group_rows; g.cb_valid = cb->rows_valid;
  }
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  if (du) {
    const csn_mat* A1 = du->A; const csn_mat* B1 = du->B; const csn_out* D1 = du->D;
    CSN_CHECK_ARG(BN == 256 && !a_mn && b_mn && A1->major == CSN_MAJOR_MN && B1->major == CSN_MAJOR_MN,
                  "csn_gemm_dual: N > 128; problem 0 = (K-major A, MN-major B), problem 1 = (MN-major A, MN-major B)");
    CSN_CHECK_ARG(A1->dtype == A->dtype && B1->dtype == A->dtype && D1->dtype == D->dtype && D1->ld == D->ld && !D1->accumulate &&
                  !D1->transposed && split_k == 1 && g.tma_store, "csn_gemm_dual: the two problems must share dtypes, the output "
                  "leading dimension and a row-major, 16-byte aligned, non-accumulating output");
    const long long delta = (reinterpret_cast<const char*>(D1->ptr) - reinterpret_cast<const char*>(D->ptr)) / esz;
    CSN_CHECK_ARG(delta >= 0 && delta + N <= D->ld, "csn_gemm_dual: D1 must start inside the rows of D0 (same buffer, column offset)");
    rc = make_tmap_2d(&tmA2, A1->ptr, A1->dtype, A1->inner, A1->outer, A1->ld, 64, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB2, B1->ptr, B1->dtype, B1->inner, B1->outer, B1->ld, 64, 64);
    if (rc) return rc;
    for (int i = 0; i < 4; ++i) {
      g.a2_mn_off[i] = A1->mn_off[i]; g.a2_k_off[i] = A1->k_off[i];
      g.b2_mn_off[i] = B1->mn_off[i]; g.b2_k_off[i] = B1->k_off[i];
      CSN_CHECK_ARG(D1->off[i] >= 0, "csn_gemm_dual: negative output offset");
      g.d2_row_off[i] = (int)(D1->off[i] / D1->ld);
      g.d2_col_off[i] = (int)(D1->off[i] % D1->ld);
    }
    g.d2_col_base = (int)delta;
    g.idesc2 = umma_idesc_f16(A->dtype == CSN_F16 ? 0u : 1u, 1u, 1u, (uint32_t)BN);
    g.dual = 1;
    g.total_tiles *= 2;
    CSN_CHECK_ARG(g.total_tiles < (1ll << 31), "csn_gemm_dual: too many tiles");
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (du) {
    if (CL == 2) return launch_gemm<256, false, true, 2, false, false, false, false, true>(tmA, tmB, tmD, g, s, &tmA2, &tmB2);
    return launch_gemm<256, false, true, 1, false, false, false, false, true>(tmA, tmB, tmD, g, s, &tmA2, &tmB2);
  }
  if (dl) {
    const long long nbt = (long long)nb[0] * nb[1] * nb[2] * nb[3];
    CSN_CHECK_ARG(BN == 256 && !a_mn && nbt == 1 && split_k == 1 && D->dtype != CSN_F32 && g.tma_store,
                  "csn_gemm_delta: needs N > 128, a K-major A, one batch, a 16-byte aligned row-major 16-bit dO");
    CSN_CHECK_ARG((dl->d_head == 64 || dl->d_head == 256) && N == dl->n_head * dl->d_head && dl->rows_pad % 32 == 0,
                  "csn_gemm_delta: d_head must be 64 or 256, N = n_head*d_head, rows_pad a multiple of 32");
    CUtensorMap tmO, tmOlo;
    rc = make_tmap_2d(&tmO, dl->O, D->dtype, N, dl->o_rows, dl->ldo, 64, 32);
    if (rc) return rc;
    tmOlo = tmO;
    if (dl->O_lo) {
      rc = make_tmap_2d(&tmOlo, dl->O_lo, D->dtype, N, dl->o_rows, dl->ldo, 64, 32);
      if (rc) return rc;
    }
    g.dl = 1;
    g.delta = dl->delta; g.dl_rows_pad = dl->rows_pad; g.dl_n_head = dl->n_head; g.dl_d_head = dl->d_head;
    g.dl_lo_inv = dl->O_lo ? (D->dtype == CSN_F16 ? 1.f / 2048.f : 1.f / 256.f) : 0.f;
    // the thread owns its whole row: 4 epilogue warps, per warp 2 output + 2 O + 2 O_lo slabs (96 KB) next to a 2-deep ring
    g.stages = 2; g.stg_bufs = 6; g.epi_warps = 4; g.alt_tiles = 0; g.tempty_count = 128;
    if (b_mn) {
      if (CL == 2) return launch_gemm<256, false, true, 2, false, true>(tmA, tmB, tmD, g, s, &tmO, &tmOlo);
      return launch_gemm<256, false, true, 1, false, true>(tmA, tmB, tmD, g, s, &tmO, &tmOlo);
    }
    if (CL == 2) return launch_gemm<256, false, false, 2, false, true>(tmA, tmB, tmD, g, s, &tmO, &tmOlo);
    return launch_gemm<256, false, false, 1, false, true>(tmA, tmB, tmD, g, s, &tmO, &tmOlo);
  }
  if (ln) {
    const long long nbt = (long long)nb[0] * nb[1] * nb[2] * nb[3];
    CSN_CHECK_ARG(N == 256 && !a_mn && !b_mn && nbt == 1 && split_k == 1 && D->dtype == CSN_F32 && g.tma_store,
                  "csn_gemm_res_ln: needs N == 256, K-major operands, one batch, a 16-byte aligned row-major fp32 Z");
    CSN_CHECK_ARG(ln->block_rows > 0 && ln->group_rows > 0 && ln->rows_valid > 0 && ln->rows_valid <= ln->group_rows &&
                  ln->block_rows % 32 == 0 && ln->group_rows % 32 == 0,
                  "csn_gemm_res_ln: bad row structure (block_rows and group_rows must be multiples of 32)");
    CUtensorMap tmR0, tmR1;
    rc = make_tmap_2d_any(&tmR0, ln->res0, CSN_F32, ln->res_ld, ln->res0_rows, ln->res_ld, 32, 32);
    if (rc) return rc;
    rc = make_tmap_2d_any(&tmR1, ln->res1, CSN_F32, ln->res_ld, ln->res1_rows, ln->res_ld, 32, 32);
    if (rc) return rc;
    g.ln = 1;
    g.res_sel = ln->res_sel; g.res_row = ln->res_row;
    g.block_rows = ln->block_rows; g.group_rows = ln->group_rows; g.rows_valid = ln->rows_valid; g.n_points = ln->n_points;
    g.eps = ln->eps; g.mean = ln->mean; g.rstd = ln->rstd; g.zbias = ln->zbias;
    g.drop_seed = ln->drop_seed; g.drop_thresh = drop_thresh16(ln->drop_p); g.drop_scale = drop_scale_of(g.drop_thresh);
    g.drop_epoch = g.drop_thresh ? drop_epoch_ptr() : nullptr;
    // the thread owns its whole row; two warp sets alternate tiles (one per accumulator buffer); per warp 2 output
    // + 2 residual slabs -> 128 KB of staging next to a 2-deep operand ring (K is short: the epilogue is the critical path)
    g.stages = 2; g.stg_bufs = 4; g.epi_warps = 8; g.alt_tiles = 1; g.tempty_count = 128;
    if (g.drop_thresh) {   // training mode: dropout on the projection output (compile-time variant)
      if (CL == 2) return launch_gemm<256, false, false, 2, true, false, false, true>(tmA, tmB, tmD, g, s, &tmR0, &tmR1);
      return launch_gemm<256, false, false, 1, true, false, false, true>(tmA, tmB, tmD, g, s, &tmR0, &tmR1);
    }
    if (CL == 2) return launch_gemm<256, false, false, 2, true>(tmA, tmB, tmD, g, s, &tmR0, &tmR1);
    return launch_gemm<256, false, false, 1, true>(tmA, tmB, tmD, g, s, &tmR0, &tmR1);
  }
  if (cb) {
    CSN_CHECK_ARG(BN == 256 && !a_mn && !b_mn, "csn_gemm_colbias: needs N > 128 and K-major operands");
    if (CL == 2) return launch_gemm<256, false, false, 2, false, false, true>(tmA, tmB, tmD, g, s);
    return launch_gemm<256, false, false, 1, false, false, true>(tmA, tmB, tmD, g, s);
  }
  if (CL == 2) {
    if (BN == 256) return dispatch_major<256, 2>(a_mn, b_mn, tmA, tmB, tmD, g, s);
    return dispatch_major<128, 2>(a_mn, b_mn, tmA, tmB, tmD, g, s);
  }
  if (BN == 256) return dispatch_major<256, 1>(a_mn, b_mn, tmA, tmB, tmD, g, s);
  if (BN == 128) return dispatch_major<128, 1>(a_mn, b_mn, tmA, tmB, tmD, g, s);
  return dispatch_major<64, 1>(a_mn, b_mn, tmA, tmB, tmD, g, s);
}

extern "C" int csn_gemm(const csn_mat* A, const csn_mat* B, const csn_out* D, int32_t M, int32_t N,
                        int32_t K, const int32_t nb[4], float alpha, int32_t split_k, void* stream) {
  return gemm_impl(A, B, D, M, N, K, nb, alpha, split_k, stream, nullptr);
}

extern "C" int csn_gemm_dual(const csn_mat* A0, const csn_mat* B0, const csn_out* D0, const csn_mat* A1, const csn_mat* B1,
                             const csn_out* D1, int32_t M, int32_t N, int32_t K, const int32_t nb[4], float alpha,
                             void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(A1 && B1 && D1, "csn_gemm_dual: null argument");
  DualProblem du{A1, B1, D1};
  return gemm_impl(A0, B0, D0, M, N, K, nb, alpha, 1, stream, nullptr, nullptr, nullptr, &du);
}

extern "C" int csn_gemm_res_ln(const csn_mat* A, const csn_mat* B, float* Z, int64_t ldz, int32_t M, int32_t K,
                               float alpha, const float* res0, int64_t res0_rows, const float* res1,
                               int64_t res1_rows, const int32_t* res_sel, const int32_t* res_row, int64_t res_ld,
                               int32_t n_points, int32_t block_rows, int32_t group_rows, int32_t rows_valid,
                               float eps, float* mean, float* rstd, const float* zbias, uint32_t drop_seed, float drop_p,
                               void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Z && res0 && res_sel && res_row && mean && rstd, "csn_gemm_res_ln: null pointer");
  CSN_CHECK_ARG(n_points > 0 && n_points <= res_ld, "csn_gemm_res_ln: n_points (%d) must be in (0, res_ld]", n_points);
  csn_out D;
  memset(&D, 0, sizeof(D));
  D.ptr = Z; D.dtype = CSN_F32; D.ld = ldz;
  LnEpilogue ln{res0, res0_rows, res1 ? res1 : res0, res1 ? res1_rows : res0_rows, res_sel, res_row, res_ld,
                block_rows, group_rows, rows_valid, n_points, eps, mean, rstd, zbias, drop_seed, drop_p};
  CSN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "csn_gemm_res_ln: dropout probability outside [0, 1)");
  CSN_CHECK_ARG(!zbias || (reinterpret_cast<uintptr_t>(zbias) & 15) == 0, "csn_gemm_res_ln: zbias must be 16-byte aligned");
  const int32_t nb[4] = {1, 1, 1, 1};
  return gemm_impl(A, B, &D, M, 256, K, nb, alpha, 1, stream, &ln);
}

extern "C" int csn_gemm_colbias(const csn_mat* A, const csn_mat* B, const csn_out* D, int32_t M, int32_t N, int32_t K,
                                float alpha, const float* bias, int64_t bias_ld, int32_t col0, int32_t group_rows,
                                int32_t rows_valid, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(bias != nullptr, "csn_gemm_colbias: null bias");
  ColBias cb{bias, bias_ld, col0, group_rows, rows_valid};
  const int32_t nb[4] = {1, 1, 1, 1};
  return gemm_impl(A, B, D, M, N, K, nb, alpha, 1, stream, nullptr, nullptr, &cb);
}

extern "C" int csn_gemm_delta(const csn_mat* A, const csn_mat* B, void* dO, int64_t lddo, int32_t M, int32_t N, int32_t K,
                              float alpha, const void* O, const void* O_lo, int64_t ldo, float* delta,
                              int32_t rows_pad, int32_t n_head, int32_t d_head, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(A && B && dO && O && delta, "csn_gemm_delta: null pointer");
  csn_out D;
  memset(&D, 0, sizeof(D));
  D.ptr = dO; D.dtype = A->dtype; D.ld = lddo;
  DeltaEpilogue dl{O, O_lo, ldo, (long long)M, delta, rows_pad, n_head, d_head};
  const int32_t nb[4] = {1, 1, 1, 1};
  return gemm_impl(A, B, &D, M, N, K, nb, alpha, 1, stream, nullptr, &dl);
}
