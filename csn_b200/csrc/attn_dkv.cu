// Key-stationary attention backward for d_head = 64 (MinkowskiNet CSA head, h = 4):
//   dV = P^T dO   and   dK = dS^T Q,   dS = P o (dP - delta) / sqrt(d),   P = exp(S / sqrt(d) - lse)
// for one tile of 128 keys per work item, with P and dS recomputed tile by tile and never written to HBM
// (reference: autograd of ScaledDotProductAttention.forward, MinkowskiNet/models/attention.py:69-75, which
// keeps the full (B, h, Lq, Lk) attention matrix).  Together with csn_attn_bwd_dq (dQ, query-stationary) this
// is the whole attention backward in two passes and O(L) memory: at d_head 64 the two [128 x 128] score
// accumulators, BOTH [128 x 64] output accumulators and the 16-bit P^T / dS^T operands of the output MMAs fit in
// TMEM (128 + 128 + 64 + 64 + 64 + 64 = 512 columns): P^T and dS^T go from the element-wise warps' registers
// straight back into TMEM (tcgen05.st) and are consumed as the A operand from there (lane = key = M, 16-bit pairs
// along the query = K dimension), so they never touch shared memory.  (With 128B-swizzled SMEM staging tiles the
// st.shared traffic alone was 31 % of the kernel: profiles/r2_experiments.md.)
//
// CTA = 640 threads, one per SM, persistent over items:
//   warp 0     TMA producer : K_i, V_i resident per item; ring of 16 KB slots streaming Q_j, dO_j (for the score
//                             MMAs, K-major) and again dO_j, Q_j (for the output MMAs, MN-major: same bytes)
//   warp 1     MMA issuer   : S^T = K_i Q_j^T, dP^T = V_i dO_j^T (lane = key, column = query), then
//                             dV += P^T dO_j, dK += dS^T Q_j; the score MMAs of tile j+1 go ahead of the output
//                             MMAs of tile j
//   warp 2     TMEM allocator
//   warps 4-19 element-wise : four warps per TMEM lane quadrant (32 query columns each).  lse / delta are per
//                             COLUMN here: staged per tile in SMEM and read as broadcast float4.
//                             P^T and dS^T -> TMEM (16-bit pairs) = A operands of the output MMAs.
//                             Epilogue: warps 4-7 write dV, warps 8-11 dK (TMA stores).
#include <cstdlib>
#include <type_traits>

#include "attn_common.cuh"

namespace csn {

struct DkvItem {
  int k_row0;    // first resident key row (row of the K and V views)
  int k_valid;   // rows of the tile that exist (<= 128); the others are written as zeros
  int q_row0;    // first streamed query row (row of the Q view)
  int q_len;     // number of streamed query rows
  int o_row0;    // first output row (dK and dV)
  int col0;      // head * 64
  int stat_off;  // lse[stat_off + n], delta[stat_off + n] for streamed query n
  int do_row0;   // first streamed row of dO
  int key0;      // index of the first resident key inside its chunk (dropout mask column)
  int pad0, pad1, pad2;
};

struct DkvArgs {
  const DkvItem* items;
  int n_items;
  const float* lse;
  const float* delta;
  const float* lse2;    // dkv kernel: lse * log2(e) and delta / sqrt(d), pre-scaled (bulk-copied per tile by the producer)
  const float* dlt_s;
  float scale_log2, scale;
  int dtype;
  uint32_t idesc_s;   // M=128, N=128, K-major x K-major
  uint32_t idesc_o;   // M=128, N=64, A K-major, B MN-major
  uint32_t drop_seed, drop_thresh;
  float drop_scale;
  const uint32_t* drop_epoch;
};

struct DkvCfg {
  static constexpr int TILE_BYTES = 128 * 64 * 2;     // K_i, V_i
  static constexpr int OUT_BYTES = 8 * 4096;          // output slabs of the epilogue: [32 rows x 64 columns] per warp
  static constexpr int SLOT_BYTES = 128 * 64 * 2;
  static constexpr int NST = 8;
#ifndef CSN_DKV_EW_WARPS
#define CSN_DKV_EW_WARPS 16
#endif
  static constexpr int EW_WARPS = CSN_DKV_EW_WARPS;    // 8: 64 query columns per thread in steps of 32; 16: 32 in steps of 16
  static constexpr int COLS = 128 / (EW_WARPS / 4);
  static constexpr int CW = EW_WARPS == 16 ? 16 : 32;   // columns per tcgen05.ld
  static constexpr int EW_THREADS = 32 * EW_WARPS;
  static constexpr int THREADS = 128 + EW_THREADS;
  static constexpr int NSTAT = 4;                      // ring of per-tile statistics: [lse | delta][128 queries]
  static constexpr int STAT_BYTES = NSTAT * 2 * 128 * 4;
  static constexpr int BAR_BYTES = 384;
  static constexpr int SMEM_BYTES = 2 * TILE_BYTES + OUT_BYTES + NST * SLOT_BYTES + STAT_BYTES + BAR_BYTES + 1024;
  // TMEM: S^T @0, dP^T @128 (fp32), dV @256, dK @320 (fp32), P^T @384, dS^T @448 (16-bit pairs: 128 queries = 64 columns)
  static constexpr int DV_COL = 256, DK_COL = 320, PT_COL = 384, DST_COL = 448;
};

template <bool DROP>
__global__ void __launch_bounds__(DkvCfg::THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                    const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
                    const __grid_constant__ DkvArgs p) {
  using Cfg = DkvCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + Cfg::TILE_BYTES;
  const uint32_t sOut = sV + Cfg::TILE_BYTES;
  const uint32_t sRing = sOut + Cfg::OUT_BYTES;
  float* stat = reinterpret_cast<float*>(smem + 2 * Cfg::TILE_BYTES + Cfg::OUT_BYTES + Cfg::NST * Cfg::SLOT_BYTES);
  uint8_t* bar_ptr = reinterpret_cast<uint8_t*>(stat) + Cfg::STAT_BYTES;
  const uint32_t bar_base = smem_u32(bar_ptr);
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t bres_full = bar_base + 8u * (2 * Cfg::NST + 0);   // K_i, V_i landed
  const uint32_t bres_empty = bar_base + 8u * (2 * Cfg::NST + 1);  // every score MMA of the item has been issued
  const uint32_t sdp_full = bar_base + 8u * (2 * Cfg::NST + 2);    // S^T, dP^T of tile j are in TMEM
  const uint32_t sdp_empty = bar_base + 8u * (2 * Cfg::NST + 3);   // ... and have been read
  const uint32_t st_full = bar_base + 8u * (2 * Cfg::NST + 4);     // P^T, dS^T of tile j are in TMEM
  const uint32_t st_empty = bar_base + 8u * (2 * Cfg::NST + 5);    // output MMAs of tile j done reading them
  const uint32_t acc_full = bar_base + 8u * (2 * Cfg::NST + 6);
  const uint32_t acc_empty = bar_base + 8u * (2 * Cfg::NST + 7);
  auto stat_full = [&](int s) { return bar_base + 8u * (2 * Cfg::NST + 8 + s); };
  auto stat_empty = [&](int s) { return bar_base + 8u * (2 * Cfg::NST + 8 + Cfg::NSTAT + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 8 + 2 * Cfg::NSTAT));
  const uint32_t sStat = smem_u32(stat);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dseed = DROP ? drop_seed_eff(p.drop_seed, p.drop_epoch) : 0u;   // (one load per thread, train-mode variants only)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::NST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int s = 0; s < Cfg::NSTAT; ++s) {
      mbar_init(stat_full(s), 1);
      mbar_init(stat_empty(s), Cfg::EW_THREADS);
    }
    mbar_init(bres_full, 1);
    mbar_init(bres_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, Cfg::EW_THREADS);
    mbar_init(st_full, Cfg::EW_THREADS);
    mbar_init(st_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, Cfg::EW_THREADS);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int st = 0, ss = 0;
      uint32_t ph = 0, r_ph = 0, s_ph = 0;
      auto load_slot = [&](const CUtensorMap* tm, int col0, int row0) {
        mbar_wait(kv_empty(st), ph ^ 1);
        mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
        tma_load_2d(sRing + st * Cfg::SLOT_BYTES, tm, kv_full(st), col0, row0);
        if (++st == Cfg::NST) { st = 0; ph ^= 1; }
      };
      auto load_stats = [&](const DkvItem& it, int j) {   // lse2 | delta_s of the 128 queries of tile j: two 512 B bulk copies
        mbar_wait(stat_empty(ss), s_ph ^ 1);
        mbar_arrive_expect_tx(stat_full(ss), 1024);
        bulk_load_1d(sStat + ss * 1024, p.lse2 + it.stat_off + j * 128, 512, stat_full(ss));
        bulk_load_1d(sStat + ss * 1024 + 512, p.dlt_s + it.stat_off + j * 128, 512, stat_full(ss));
        if (++ss == Cfg::NSTAT) { ss = 0; s_ph ^= 1; }
      };
      for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
        const DkvItem it = p.items[wk];
        const int nq = (it.q_len + 127) >> 7;
        mbar_wait(bres_empty, r_ph ^ 1);
        mbar_arrive_expect_tx(bres_full, 2 * Cfg::TILE_BYTES);
        tma_load_2d(sK, &tmK, bres_full, it.col0, it.k_row0);
        tma_load_2d(sV, &tmV, bres_full, it.col0, it.k_row0);
        r_ph ^= 1;
        // consumption order of the MMA warp: [Q0, dO0], [Q1, dO1], dO0, Q0, [Q2, dO2], dO1, Q1, ...
        load_stats(it, 0);
        load_slot(&tmQ, it.col0, it.q_row0);
        load_slot(&tmDO, it.col0, it.do_row0);
        for (int j = 0; j < nq; ++j) {
          if (j + 1 < nq) {
            load_stats(it, j + 1);
            load_slot(&tmQ, it.col0, it.q_row0 + (j + 1) * 128);
            load_slot(&tmDO, it.col0, it.do_row0 + (j + 1) * 128);
          }
          load_slot(&tmDO, it.col0, it.do_row0 + j * 128);   // for dV (same bytes, consumed MN-major)
          load_slot(&tmQ, it.col0, it.q_row0 + j * 128);     // for dK
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, r_ph = 0, sdp_ph = 0, stf_ph = 0, acc_ph = 0;
      auto mma_scores = [&](uint32_t a_tile, uint32_t d_tmem) {   // D[key][query] = A_tile (resident) x slot^T, K = 64
        mbar_wait(kv_full(st), ph);
        tc_fence_after();
        const uint32_t b_tile = sRing + st * Cfg::SLOT_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(d_tmem, umma_desc_sw128(a_tile + k * 32, 0, 1024), umma_desc_sw128(b_tile + k * 32, 0, 1024),
                      p.idesc_s, k ? 1u : 0u);
        umma_commit(kv_empty(st));
        if (++st == Cfg::NST) { st = 0; ph ^= 1; }
      };
      auto mma_out = [&](uint32_t a_tmem, uint32_t d_tmem, bool accumulate) {   // D[key][d] += A[key][query] (TMEM) x slot
        mbar_wait(kv_full(st), ph);
        tc_fence_after();
        const uint32_t b_tile = sRing + st * Cfg::SLOT_BYTES;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 128 queries in steps of 16 = 8 TMEM columns of 16-bit pairs
          umma_f16_ts(d_tmem, a_tmem + k * 8, umma_desc_sw128(b_tile + k * 2048, 16384, 1024), p.idesc_o,
                      (accumulate || k) ? 1u : 0u);
        umma_commit(kv_empty(st));
        if (++st == Cfg::NST) { st = 0; ph ^= 1; }
      };
      for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
        const DkvItem it = p.items[wk];
        const int nq = (it.q_len + 127) >> 7;
        mbar_wait(bres_full, r_ph);
        r_ph ^= 1;
        tc_fence_after();
        auto issue_scores = [&](int j) {
          mbar_wait(sdp_empty, sdp_ph ^ 1);   // the scores of the previous tile have been read
          tc_fence_after();
          mma_scores(sK, tmem_base);          // S^T  = K_i Q_j^T
          mma_scores(sV, tmem_base + 128);    // dP^T = V_i dO_j^T
          umma_commit(sdp_full);
          if (j == nq - 1) umma_commit(bres_empty);
          sdp_ph ^= 1;
        };
        issue_scores(0);
        for (int j = 0; j < nq; ++j) {
          if (j + 1 < nq) issue_scores(j + 1);
          mbar_wait(st_full, stf_ph);
          stf_ph ^= 1;
          if (j == 0) mbar_wait(acc_empty, acc_ph ^ 1);   // previous item's accumulators have been read out
          tc_fence_after();
          mma_out(tmem_base + Cfg::PT_COL, tmem_base + Cfg::DV_COL, j != 0);    // dV += P^T  dO_j
          mma_out(tmem_base + Cfg::DST_COL, tmem_base + Cfg::DK_COL, j != 0);   // dK += dS^T Q_j
          umma_commit(st_empty);
        }
        umma_commit(acc_full);
        acc_ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================================================== element-wise stage + epilogue
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;            // query columns [COLS half, COLS half + COLS)
    constexpr int COLS = Cfg::COLS, CW = Cfg::CW;
    const int r = q * 32 + lane;                 // key row of the tile owned by this thread
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t sdp_ph = 0, ste_ph = 0, accf_ph = 0, stat_ph = 0;
    int ss = 0;
    for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
      const DkvItem it = p.items[wk];
      const int nq = (it.q_len + 127) >> 7;
      const bool kvalid = r < it.k_valid;
      for (int j = 0; j < nq; ++j) {
        const int nvalid = min(128, it.q_len - j * 128);
        // per-query statistics of this tile (lse in log2 units, delta pre-scaled): bulk-copied by the producer warp
        // into a ring; no block-wide barrier anywhere in this loop (it cost 23 % of the warps' time: ncu)
        mbar_wait(stat_full(ss), stat_ph);
        const uint32_t st_l = sStat + ss * 1024, st_d = st_l + 512;   // shared-window addresses (ld.shared, not generic loads)
        mbar_wait(sdp_full, sdp_ph);
        sdp_ph ^= 1;
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr;
        // all of this thread's scores into registers first: the MMA warp gets the accumulators back before the
        // arithmetic starts (the chain scores -> element-wise -> scores is what bounds the kernel)
        static_assert(COLS == 2 * CW, "two register chunks per thread");
        uint32_t svA[CW], dvA[CW], svB[CW], dvB[CW];
        tmem_ld_cols(s_addr + half * COLS, svA);
        tmem_ld_cols(s_addr + 128 + half * COLS, dvA);
        tmem_ld_cols(s_addr + half * COLS + CW, svB);
        tmem_ld_cols(s_addr + 128 + half * COLS + CW, dvB);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(sdp_empty);
        auto ew_tile = [&](auto F16C, auto FULLC) {
          constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = half * COLS + cc * CW;
            const uint32_t (&sv)[CW] = cc ? svB : svA;
            const uint32_t (&dv)[CW] = cc ? dvB : dvA;
            uint32_t pp[CW / 2], pd[CW / 2];
#pragma unroll
            for (int i = 0; i < CW; i += 4) {
              const float4 l4 = lds_f4(st_l + (c + i) * 4);
              const float4 d4 = lds_f4(st_d + (c + i) * 4);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, ds_[4] = {d4.x, d4.y, d4.z, d4.w};
              float pv[4], dsv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float m = 1.f;   // d P_dropped / d P = mask / (1 - p)
                if (DROP) {      // the forward mask: row id = query, pair = two adjacent keys
                  const uint32_t hh = drop_pair(drop_row_key(dseed, (uint32_t)(it.stat_off + j * 128 + c + i + u)),
                                                (uint32_t)(it.key0 + r) >> 1);
                  m = (((it.key0 + r) & 1) ? drop_keep_hi(hh, p.drop_thresh) : drop_keep_lo(hh, p.drop_thresh)) ? p.drop_scale : 0.f;
                }
                const float pr = fast_exp2(__uint_as_float(sv[i + u]) * p.scale_log2 - ls[u]);
                float pe = DROP ? pr * m : pr;
                float de = pr * fmaf(__uint_as_float(dv[i + u]), DROP ? m * p.scale : p.scale, -ds_[u]);
                if (!FULL) {
                  if (!(kvalid && c + i + u < nvalid)) { pe = 0.f; de = 0.f; }
                }
                pv[u] = pe;
                dsv[u] = de;
              }
              if (F16) {
                __half2 a = __floats2half2_rn(pv[0], pv[1]), b = __floats2half2_rn(pv[2], pv[3]);
                __half2 e = __floats2half2_rn(dsv[0], dsv[1]), f = __floats2half2_rn(dsv[2], dsv[3]);
                pp[i >> 1] = *reinterpret_cast<uint32_t*>(&a); pp[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&b);
                pd[i >> 1] = *reinterpret_cast<uint32_t*>(&e); pd[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&f);
              } else {
                __nv_bfloat162 a = __floats2bfloat162_rn(pv[0], pv[1]), b = __floats2bfloat162_rn(pv[2], pv[3]);
                __nv_bfloat162 e = __floats2bfloat162_rn(dsv[0], dsv[1]), f = __floats2bfloat162_rn(dsv[2], dsv[3]);
                pp[i >> 1] = *reinterpret_cast<uint32_t*>(&a); pp[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&b);
                pd[i >> 1] = *reinterpret_cast<uint32_t*>(&e); pd[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&f);
              }
            }
            if (c == half * COLS) {
              mbar_wait(st_empty, ste_ph ^ 1);   // output MMAs of the previous tile no longer read P^T / dS^T
              ste_ph ^= 1;
              tc_fence_after();
            }
            tmem_st_cols(s_addr + Cfg::PT_COL + (c >> 1), pp);    // queries c .. c+CW-1 = 16-bit pairs in CW/2 columns
            tmem_st_cols(s_addr + Cfg::DST_COL + (c >> 1), pd);
          }
        };
        const bool full = (nvalid == 128) && __all_sync(0xffffffffu, kvalid);
        if (p.dtype == CSN_F16) {
          if (full) ew_tile(std::true_type{}, std::true_type{}); else ew_tile(std::true_type{}, std::false_type{});
        } else {
          if (full) ew_tile(std::false_type{}, std::true_type{}); else ew_tile(std::false_type{}, std::false_type{});
        }
        mbar_arrive(stat_empty(ss));   // (every read of the statistics precedes the stores above in program order)
        if (++ss == Cfg::NSTAT) { ss = 0; stat_ph ^= 1; }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(st_full);
      }
      // ---- epilogue: warps 4-7 write dV, warps 8-11 dK; each warp stages its 32 rows in its own rows of one of the
      //      (now idle) staging tiles and hands the [32 x 64] slab to the TMA store engine
      mbar_wait(acc_full, accf_ph);
      accf_ph ^= 1;
      tc_fence_after();
      if (half >= 2) {
        mbar_arrive(acc_empty);
      } else {
        const uint32_t o_addr = tmem_base + lane_addr + (half == 0 ? Cfg::DV_COL : Cfg::DK_COL);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(o_addr, v0);
        tmem_ld_32x32(o_addr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(acc_empty);   // the accumulators are in registers
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = kvalid ? __uint_as_float(v0[i]) : 0.f, a1 = kvalid ? __uint_as_float(v0[i + 1]) : 0.f;
          const float b0 = kvalid ? __uint_as_float(v1[i]) : 0.f, b1 = kvalid ? __uint_as_float(v1[i + 1]) : 0.f;
          if (p.dtype == CSN_F16) {
            __half2 x = __floats2half2_rn(a0, a1), y = __floats2half2_rn(b0, b1);
            w[i >> 1] = *reinterpret_cast<uint32_t*>(&x); w[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&y);
          } else {
            __nv_bfloat162 x = __floats2bfloat162_rn(a0, a1), y = __floats2bfloat162_rn(b0, b1);
            w[i >> 1] = *reinterpret_cast<uint32_t*>(&x); w[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&y);
          }
        }
        const uint32_t buf = sOut + (half * 4 + q) * 4096;
        const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t a = rowaddr + (((uint32_t)t ^ ((uint32_t)lane & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * t]), "r"(w[4 * t + 1]), "r"(w[4 * t + 2]), "r"(w[4 * t + 3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(half == 0 ? &tmDV : &tmDK, buf, it.col0, it.o_row0 + q * 32);
          tma_store_commit();
          tma_store_wait_read<0>();   // the slab is rewritten by the next item's epilogue
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Query-stationary counterpart: dQ = dS K of one 128-query tile per item, nothing written but dQ (the variant
// of attn_bwd.cu's kernel used when no dS buffer is requested).  Same machinery: S = Q_i K_j^T and dP = dO_i V_j^T
// (lane = query, column = key), 16 element-wise warps, dS goes back into TMEM as 16-bit pairs (two buffers) and is
// the A operand of dQ += dS K_j.  TMEM: S @0, dP @128, dQ @256, dS16 @320 / @384.
struct Dq64Item {   // = AttnBwdItem of attn_bwd.cu
  int q_row0, q_valid, kv_row0, kv_len, o_row0, col0, stat_off, ds_row0, ds_col0, flags, pad0, pad1;
};

struct Dq64Cfg {
  static constexpr int TILE_BYTES = 128 * 64 * 2;
  static constexpr int OUT_BYTES = 4 * 4096;
  static constexpr int SLOT_BYTES = 128 * 64 * 2;
  static constexpr int NST = 9;
  static constexpr int EW_WARPS = 16;
  static constexpr int EW_THREADS = 32 * EW_WARPS;
  static constexpr int THREADS = 128 + EW_THREADS;
  static constexpr int COLS = 32, CW = 16;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 2 * TILE_BYTES + OUT_BYTES + NST * SLOT_BYTES + BAR_BYTES + 1024;
  static constexpr int DQ_COL = 256, DS_COL = 320;
};

template <bool DROP>
__global__ void __launch_bounds__(Dq64Cfg::THREADS, 1)
attn_bwd_dq64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ DkvArgs p) {
  using Cfg = Dq64Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sDO = sQ + Cfg::TILE_BYTES;
  const uint32_t sOut = sDO + Cfg::TILE_BYTES;
  const uint32_t sRing = sOut + Cfg::OUT_BYTES;
  uint8_t* bar_ptr = smem + 2 * Cfg::TILE_BYTES + Cfg::OUT_BYTES + Cfg::NST * Cfg::SLOT_BYTES;
  const uint32_t bar_base = smem_u32(bar_ptr);
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t bres_full = bar_base + 8u * (2 * Cfg::NST + 0);
  const uint32_t bres_empty = bar_base + 8u * (2 * Cfg::NST + 1);
  const uint32_t sdp_full = bar_base + 8u * (2 * Cfg::NST + 2);
  const uint32_t sdp_empty = bar_base + 8u * (2 * Cfg::NST + 3);
  auto ds_full = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 4 + b); };    // dS_j is in TMEM buffer b
  auto ds_empty = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 6 + b); };   // dQ MMAs of tile j done reading it
  const uint32_t acc_full = bar_base + 8u * (2 * Cfg::NST + 8);
  const uint32_t acc_empty = bar_base + 8u * (2 * Cfg::NST + 9);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 10));
  const Dq64Item* items = reinterpret_cast<const Dq64Item*>(p.items);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dseed = DROP ? drop_seed_eff(p.drop_seed, p.drop_epoch) : 0u;   // (one load per thread, train-mode variants only)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::NST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(bres_full, 1);
    mbar_init(bres_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(sdp_empty, Cfg::EW_THREADS);
    for (int b = 0; b < 2; ++b) {
      mbar_init(ds_full(b), Cfg::EW_THREADS);
      mbar_init(ds_empty(b), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, Cfg::EW_THREADS);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, r_ph = 0;
      auto load_slot = [&](const CUtensorMap* tm, int col0, int row0) {
        mbar_wait(kv_empty(st), ph ^ 1);
        mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
        tma_load_2d(sRing + st * Cfg::SLOT_BYTES, tm, kv_full(st), col0, row0);
        if (++st == Cfg::NST) { st = 0; ph ^= 1; }
      };
      for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
        const Dq64Item it = items[wk];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bres_empty, r_ph ^ 1);
        mbar_arrive_expect_tx(bres_full, 2 * Cfg::TILE_BYTES);
        tma_load_2d(sQ, &tmQ, bres_full, it.col0, it.q_row0);
        tma_load_2d(sDO, &tmDO, bres_full, it.col0, it.o_row0);
        r_ph ^= 1;
        // consumption order of the MMA warp: [K0, V0], [K1, V1], K0 (dQ), [K2, V2], K1 (dQ), ...
        load_slot(&tmK, it.col0, it.kv_row0);
        load_slot(&tmV, it.col0, it.kv_row0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) {
            load_slot(&tmK, it.col0, it.kv_row0 + (j + 1) * 128);
            load_slot(&tmV, it.col0, it.kv_row0 + (j + 1) * 128);
          }
          load_slot(&tmK, it.col0, it.kv_row0 + j * 128);   // for dQ (same bytes, consumed MN-major)
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, r_ph = 0, sdp_ph = 0, acc_ph = 0;
      uint32_t dsf_ph[2] = {0, 0};
      auto mma_scores = [&](uint32_t a_tile, uint32_t d_tmem) {
        mbar_wait(kv_full(st), ph);
        tc_fence_after();
        const uint32_t b_tile = sRing + st * Cfg::SLOT_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(d_tmem, umma_desc_sw128(a_tile + k * 32, 0, 1024), umma_desc_sw128(b_tile + k * 32, 0, 1024),
                      p.idesc_s, k ? 1u : 0u);
        umma_commit(kv_empty(st));
        if (++st == Cfg::NST) { st = 0; ph ^= 1; }
      };
      for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
        const Dq64Item it = items[wk];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bres_full, r_ph);
        r_ph ^= 1;
        tc_fence_after();
        auto issue_scores = [&](int j) {
          mbar_wait(sdp_empty, sdp_ph ^ 1);
          tc_fence_after();
          mma_scores(sQ, tmem_base);           // S  = Q_i  K_j^T
          mma_scores(sDO, tmem_base + 128);    // dP = dO_i V_j^T
          umma_commit(sdp_full);
          if (j == nkv - 1) umma_commit(bres_empty);
          sdp_ph ^= 1;
        };
        issue_scores(0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) issue_scores(j + 1);
          const int b = j & 1;
          mbar_wait(ds_full(b), dsf_ph[b]);
          dsf_ph[b] ^= 1;
          if (j == 0) mbar_wait(acc_empty, acc_ph ^ 1);
          tc_fence_after();
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          const uint32_t b_tile = sRing + st * Cfg::SLOT_BYTES;
#pragma unroll
          for (int k = 0; k < 8; ++k)   // 128 keys in steps of 16 = 8 TMEM columns of 16-bit pairs
            umma_f16_ts(tmem_base + Cfg::DQ_COL, tmem_base + Cfg::DS_COL + b * 64 + k * 8,
                        umma_desc_sw128(b_tile + k * 2048, 16384, 1024), p.idesc_o, (j | k) ? 1u : 0u);
          umma_commit(kv_empty(st));
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          umma_commit(ds_empty(b));
        }
        umma_commit(acc_full);
        acc_ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================================================== element-wise stage + epilogue
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;            // key columns [32 half, 32 half + 32)
    constexpr int COLS = Cfg::COLS, CW = Cfg::CW;
    const int r = q * 32 + lane;                 // query row of the tile owned by this thread
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t sdp_ph = 0, accf_ph = 0;
    uint32_t dse_ph[2] = {0, 0};
    constexpr float LOG2E = 1.4426950408889634f;
    for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
      const Dq64Item it = items[wk];
      const int nkv = (it.kv_len + 127) >> 7;
      const bool valid = r < it.q_valid;
      const float lse_l2 = valid ? p.lse[it.stat_off + r] * LOG2E : 0.f;
      const float dlt_s = valid ? p.delta[it.stat_off + r] * p.scale : 0.f;
      const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.stat_off + r));
      for (int j = 0; j < nkv; ++j) {
        const int b = j & 1;
        const int nvalid = min(128, it.kv_len - j * 128);
        mbar_wait(sdp_full, sdp_ph);
        sdp_ph ^= 1;
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr;
        // all of this thread's scores (2 x 2 x 16 columns) into registers first: the MMA warp gets the accumulators
        // back before the arithmetic starts
        uint32_t svA[CW], dvA[CW], svB[CW], dvB[CW];
        tmem_ld_cols(s_addr + half * COLS, svA);
        tmem_ld_cols(s_addr + 128 + half * COLS, dvA);
        tmem_ld_cols(s_addr + half * COLS + CW, svB);
        tmem_ld_cols(s_addr + 128 + half * COLS + CW, dvB);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(sdp_empty);
        auto ew_tile = [&](auto F16C, auto FULLC) {
          constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = half * COLS + cc * CW;
            const uint32_t (&sv)[CW] = cc ? svB : svA;
            const uint32_t (&dv)[CW] = cc ? dvB : dvA;
            uint32_t pd[CW / 2];
#pragma unroll
            for (int i = 0; i < CW; i += 2) {
              float m0 = p.scale, m1 = p.scale;   // (d P_dropped / d P = mask / (1 - p)) * scale
              if (DROP) {
                const uint32_t hh = drop_pair(rk, (uint32_t)(j * 128 + c + i) >> 1);
                m0 = drop_keep_lo(hh, p.drop_thresh) ? p.drop_scale * p.scale : 0.f;
                m1 = drop_keep_hi(hh, p.drop_thresh) ? p.drop_scale * p.scale : 0.f;
              }
              float d0 = fast_exp2(__uint_as_float(sv[i]) * p.scale_log2 - lse_l2) * fmaf(__uint_as_float(dv[i]), m0, -dlt_s);
              float d1 = fast_exp2(__uint_as_float(sv[i + 1]) * p.scale_log2 - lse_l2) * fmaf(__uint_as_float(dv[i + 1]), m1, -dlt_s);
              if (!FULL) {
                if (!(valid && c + i < nvalid)) d0 = 0.f;
                if (!(valid && c + i + 1 < nvalid)) d1 = 0.f;
              }
              if (F16) {
                __half2 h = __floats2half2_rn(d0, d1);
                pd[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
              } else {
                __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
                pd[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
              }
            }
            if (c == half * COLS) {
              mbar_wait(ds_empty(b), dse_ph[b] ^ 1);   // dQ MMAs of tile j-2 no longer read this buffer
              dse_ph[b] ^= 1;
              tc_fence_after();
            }
            tmem_st_cols(s_addr + Cfg::DS_COL + b * 64 + (c >> 1), pd);
          }
        };
        const bool full = (nvalid == 128) && __all_sync(0xffffffffu, valid);
        if (p.dtype == CSN_F16) {
          if (full) ew_tile(std::true_type{}, std::true_type{}); else ew_tile(std::true_type{}, std::false_type{});
        } else {
          if (full) ew_tile(std::false_type{}, std::true_type{}); else ew_tile(std::false_type{}, std::false_type{});
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(ds_full(b));
      }
      // ---- epilogue: warps 4-7 write the dQ tile (rows >= q_valid: zeros)
      mbar_wait(acc_full, accf_ph);
      accf_ph ^= 1;
      tc_fence_after();
      if (half != 0) {
        mbar_arrive(acc_empty);
      } else {
        const uint32_t o_addr = tmem_base + lane_addr + Cfg::DQ_COL;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(o_addr, v0);
        tmem_ld_32x32(o_addr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(acc_empty);
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = valid ? __uint_as_float(v0[i]) : 0.f, a1 = valid ? __uint_as_float(v0[i + 1]) : 0.f;
          const float b0 = valid ? __uint_as_float(v1[i]) : 0.f, b1 = valid ? __uint_as_float(v1[i + 1]) : 0.f;
          if (p.dtype == CSN_F16) {
            __half2 x = __floats2half2_rn(a0, a1), y = __floats2half2_rn(b0, b1);
            w[i >> 1] = *reinterpret_cast<uint32_t*>(&x); w[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&y);
          } else {
            __nv_bfloat162 x = __floats2bfloat162_rn(a0, a1), y = __floats2bfloat162_rn(b0, b1);
            w[i >> 1] = *reinterpret_cast<uint32_t*>(&x); w[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&y);
          }
        }
        const uint32_t buf = sOut + q * 4096;
        const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t a = rowaddr + (((uint32_t)t ^ ((uint32_t)lane & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * t]), "r"(w[4 * t + 1]), "r"(w[4 * t + 2]), "r"(w[4 * t + 3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&tmDQ, buf, it.col0, it.o_row0 + q * 32);
          tma_store_commit();
          tma_store_wait_read<0>();   // the slab is rewritten by the next item's epilogue
        }
        __syncwarp();
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool DROP>
static int launch_dq64_d(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                         const CUtensorMap& tmDQ, const DkvArgs& a, cudaStream_t stream) {
  auto kern = attn_bwd_dq64_kernel<DROP>;
  CSN_SET_MAX_SMEM(kern, Dq64Cfg::SMEM_BYTES);
  const int grid = a.n_items < num_sms() ? a.n_items : num_sms();
  kern<<<grid, Dq64Cfg::THREADS, Dq64Cfg::SMEM_BYTES, stream>>>(tmQ, tmDO, tmK, tmV, tmDQ, a);
  CSN_LAUNCH_OK("attn_bwd_dq64_kernel");
  return 0;
}

// called by csn_attn_bwd_dq (attn_bwd.cu) for d_head 64 when no dS buffer is requested
int launch_dq64(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDQ, const int32_t* items, int32_t n_items, const float* lse, const float* delta,
                int32_t dtype, uint32_t drop_seed, float drop_p, cudaStream_t stream) {
  DkvArgs a;
  a.items = reinterpret_cast<const DkvItem*>(items);
  a.n_items = n_items;
  a.lse = lse;
  a.delta = delta;
  a.lse2 = nullptr;
  a.dlt_s = nullptr;
  a.scale = 1.0f / sqrtf(64.f);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.dtype = dtype;
  const uint32_t fmt = dtype == CSN_F16 ? 0u : 1u;
  a.idesc_s = umma_idesc_f16(fmt, 0, 0, 128);
  a.idesc_o = umma_idesc_f16(fmt, 0, 1, 64);
  a.drop_seed = drop_seed;
  a.drop_thresh = drop_thresh16(drop_p);
  a.drop_scale = drop_scale_of(a.drop_thresh);
  a.drop_epoch = a.drop_thresh ? drop_epoch_ptr() : nullptr;
  return a.drop_thresh ? launch_dq64_d<true>(tmQ, tmDO, tmK, tmV, tmDQ, a, stream) : launch_dq64_d<false>(tmQ, tmDO, tmK, tmV, tmDQ, a, stream);
}

// lse -> lse * log2(e), delta -> delta / sqrt(d): what the dK/dV kernel's element-wise stage consumes per COLUMN
__global__ void dkv_stats_scale_kernel(const float* __restrict__ lse, const float* __restrict__ delta, long long n, float scale,
                                       float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[i] = lse[i] * 1.4426950408889634f;
    out[n + i] = delta[i] * scale;
  }
}

template <bool DROP>
static int launch_dkv(const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmQ, const CUtensorMap& tmDO,
                      const CUtensorMap& tmDK, const CUtensorMap& tmDV, const DkvArgs& a, cudaStream_t stream) {
  auto kern = attn_bwd_dkv_kernel<DROP>;
  CSN_SET_MAX_SMEM(kern, DkvCfg::SMEM_BYTES);
  const int grid = a.n_items < num_sms() ? a.n_items : num_sms();
  kern<<<grid, DkvCfg::THREADS, DkvCfg::SMEM_BYTES, stream>>>(tmK, tmV, tmQ, tmDO, tmDK, tmDV, a);
  CSN_LAUNCH_OK("attn_bwd_dkv_kernel");
  return 0;
}

}  // namespace csn

extern "C" int csn_attn_bwd_dkv(const void* K, const void* V, const void* Q, const void* dO, int64_t kv_rows, int64_t q_rows,
                                int64_t do_rows, int64_t width, int64_t ldk, int64_t ldv, int64_t ldq, int64_t lddo,
                                int32_t d_head, int32_t dtype, const int32_t* items, int32_t n_items, void* dK, void* dV,
                                int64_t out_rows, int64_t ldout, const float* lse, const float* delta, int64_t n_stats,
                                float* stat_scratch, uint32_t drop_seed, float drop_p, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(K && V && Q && dO && items && dK && dV && lse && delta && stat_scratch, "csn_attn_bwd_dkv: null pointer");
  CSN_CHECK_ARG(n_stats > 0 && n_stats % 4 == 0 && (reinterpret_cast<uintptr_t>(stat_scratch) & 15) == 0,
                "csn_attn_bwd_dkv: statistics scratch must be 16B aligned, n_stats a multiple of 4");
  CSN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "csn_attn_bwd_dkv: dropout probability %f outside [0, 1)", (double)drop_p);
  CSN_CHECK_ARG(d_head == 64, "csn_attn_bwd_dkv: d_head=%d not supported (64: both output accumulators must fit in TMEM)", d_head);
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "csn_attn_bwd_dkv: 16-bit operands only");
  CSN_CHECK_ARG((ldout * 2) % 16 == 0, "csn_attn_bwd_dkv: output stride must be a 16B multiple");
  if (n_items == 0) return 0;
  CUtensorMap tmK, tmV, tmQ, tmDO, tmDK, tmDV;
  int rc = make_tmap_2d(&tmK, K, dtype, width, kv_rows, ldk, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmV, V, dtype, width, kv_rows, ldv, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmQ, Q, dtype, width, q_rows, ldq, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmDO, dO, dtype, width, do_rows, lddo, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmDK, dK, dtype, width, out_rows, ldout, 64, 32);
  if (rc) return rc;
  rc = make_tmap_2d(&tmDV, dV, dtype, width, out_rows, ldout, 64, 32);
  if (rc) return rc;
  DkvArgs a;
  a.items = reinterpret_cast<const DkvItem*>(items);
  a.n_items = n_items;
  a.lse = lse;
  a.delta = delta;
  a.lse2 = stat_scratch;
  a.dlt_s = stat_scratch + n_stats;
  a.scale = 1.0f / sqrtf((float)d_head);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.dtype = dtype;
  const uint32_t fmt = dtype == CSN_F16 ? 0u : 1u;
  a.idesc_s = umma_idesc_f16(fmt, 0, 0, 128);
  a.idesc_o = umma_idesc_f16(fmt, 0, 1, 64);
  a.drop_seed = drop_seed;
  a.drop_thresh = drop_thresh16(drop_p);
  a.drop_scale = drop_scale_of(a.drop_thresh);
  a.drop_epoch = a.drop_thresh ? drop_epoch_ptr() : nullptr;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dkv_stats_scale_kernel<<<(unsigned)((n_stats + 1023) / 1024 < 1184 ? (n_stats + 1023) / 1024 : 1184), 256, 0, s>>>(
      lse, delta, n_stats, a.scale, stat_scratch);
  CSN_LAUNCH_OK("dkv_stats_scale_kernel");
  return a.drop_thresh ? launch_dkv<true>(tmK, tmV, tmQ, tmDO, tmDK, tmDV, a, s) : launch_dkv<false>(tmK, tmV, tmQ, tmDO, tmDK, tmDV, a, s);
}
