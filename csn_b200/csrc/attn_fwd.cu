// Fused attention forward (flash-style) for sm_100a:   O = softmax(Q K^T / sqrt(d)) V   per work item,
// with the score matrix living only in TMEM (reference: ScaledDotProductAttention.forward,
// MID-FC/csa_models.py:138-144 == MinkowskiNet/models/attention.py:69-75, which materialise it).
//
// Work item (table in device memory) = 128 query rows of one (block, head) attending to the key rows
// [kv_row0, kv_row0 + kv_len).  MID-FC: kv range = the 500-point chunk of the query tile
// (block-diagonal attention, csa_models.py:83-90); MinkowskiNet: the whole key shape.
//
// CTA = 256 threads, one CTA per SM, persistent over items:
//   warp 0    TMA producer : Q tile (resident per item) + ring of K / V slots (SWIZZLE_128B)
//   warp 1    MMA issuer   : S_j = Q K_j^T (kind::f16, fp32 in TMEM, 2 S buffers) ; O += P_j V_j
//                            (V consumed MN-major straight from the projection output)
//   warp 2    TMEM allocator (512 columns: S0 | S1 | O)
//   warps 4-7 softmax      : lane = query row. tcgen05.ld S_j, running max / sum (online softmax, lazy
//                            rescale of O only when the max grows by > 2^8), P_j -> SMEM (16-bit,
//                            128B-swizzled K-major tile = A operand of the P V MMA), final O / l and
//                            LSE written by the same warps.
#include <stdlib.h>

#include <type_traits>

#include "attn_common.cuh"

namespace csn {

template <int DH>
struct AttnCfg {
  static constexpr int KB = DH / 64;                       // k-blocks of the QK^T contraction
  static constexpr int Q_BYTES = 128 * DH * 2;             // resident query tile
  static constexpr int P_BYTES = 128 * 128 * 2;            // probabilities of one KV tile (2 k-blocks)
  static constexpr int SLOT_BYTES = (DH >= 128) ? 32768 : 128 * DH * 2;
  // K tile (128 keys x DH): slots of [128 keys x (SLOT_BYTES/256) columns]
  static constexpr int K_SLOTS = (128 * DH * 2) / SLOT_BYTES;
  static constexpr int KB_PER_KSLOT = KB / K_SLOTS;
  // V tile (128 keys x DH), MN-major: slots of [KEYS_PER_VSLOT keys x DH columns]
  static constexpr int V_SLOTS = K_SLOTS;
  static constexpr int KEYS_PER_VSLOT = 128 / V_SLOTS;
  // (tried for d_head 256: a 3-deep ring + 32 KB of extra output slabs, four per warp -> 426 vs 399 us: the ring
  //  depth matters more than the slab waits; XSTG_BYTES > 0 re-enables that layout)
  static constexpr int NST = (DH >= 128) ? 4 : 8;
  static constexpr int XSTG_BYTES = 0;
  static constexpr int BAR_BYTES = 256;
  // softmax warps: one per TMEM lane quadrant, or (d_head 64, where the exponentials and not the MMAs bound the
  // kernel: 1118 us with, 493 us without them on the MinkowskiNet batch) two, the second taking key columns 64-127
#ifndef CSN_FWD256_SM_WARPS
#define CSN_FWD256_SM_WARPS 4   // 8 (column halves, row statistics exchanged through SMEM) was measured: 402 vs 389-400 us, no gain
#endif
  static constexpr int SM_WARPS = (DH == 64) ? 8 : CSN_FWD256_SM_WARPS;
  static constexpr int SM_THREADS = 32 * SM_WARPS;
  static constexpr int THREADS = 128 + SM_THREADS;
  static constexpr int COLS = 128 / (SM_WARPS / 4);          // key columns of a tile per softmax thread
  static constexpr int XCH_BYTES = (SM_WARPS == 8) ? 1024 : 0;   // row max / row sum exchange between the two halves
  static constexpr int SMEM_BYTES = Q_BYTES + P_BYTES + NST * SLOT_BYTES + XSTG_BYTES + BAR_BYTES + XCH_BYTES + 1024;
  static constexpr int O_COL = 256;                        // TMEM: S0 @0, S1 @128, O @256 (DH <= 256)
  // d_head 64: the 16-bit probabilities go back into TMEM (two buffers of 64 columns = 128 keys as 16-bit pairs,
  // behind the 64 columns of O) and are the A operand of P V from there; no SMEM round trip, P double-buffered
  // and the two softmax warps of a lane quadrant (key columns 0-63 / 64-127 of every tile) each own a PRIVATE
  // running max, row sum and [128 x 64] accumulator (O_0 @256, O_1 @320): no cross-warp agreement inside the tile
  // loop; the two partial results are merged once per item in the epilogue.
  static constexpr int P_COL = 384;
};

template <int DH, int MODE, int CL, bool DROP>
__global__ void __launch_bounds__(AttnCfg<DH>::THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmOlo, const __grid_constant__ AttnFwdArgs p) {
  using Cfg = AttnCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sP = sQ + Cfg::Q_BYTES;
  const uint32_t sKV = sP + Cfg::P_BYTES;
  const uint32_t sX = sKV + Cfg::NST * Cfg::SLOT_BYTES;   // extra output staging (d_head 256)
  const uint32_t bar_base = sX + Cfg::XSTG_BYTES;
  uint8_t* bar_ptr = smem + Cfg::Q_BYTES + Cfg::P_BYTES + Cfg::NST * Cfg::SLOT_BYTES + Cfg::XSTG_BYTES;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t bq_full = bar_base + 8u * (2 * Cfg::NST + 0);
  const uint32_t bq_empty = bar_base + 8u * (2 * Cfg::NST + 1);
  auto s_full = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 2 + b); };
  auto s_empty = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 4 + b); };
  constexpr bool TSP = (DH == 64 && MODE == 0);   // P through TMEM (forward, d_head 64)
  constexpr bool PRIV = TSP;                      // ... with one accumulator per column half
  auto bp_full = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + (b ? 10 : 6)); };
  auto bp_empty = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + (b ? 11 : 7)); };
  const uint32_t bo_full = bar_base + 8u * (2 * Cfg::NST + 8);
  const uint32_t bo_empty = bar_base + 8u * (2 * Cfg::NST + 9);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 12));
  float* xch = reinterpret_cast<float*>(bar_ptr + Cfg::BAR_BYTES);   // [2 halves][128 rows] (SM_WARPS == 8 only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dseed = DROP ? drop_seed_eff(p.drop_seed, p.drop_epoch) : 0u;   // (one load per thread, train-mode variants only)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::NST; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), CL);   // the slot is refilled by multicast: every CTA of the cluster must have drained it
    }
    mbar_init(bq_full, 1);
    mbar_init(bq_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), Cfg::SM_THREADS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bp_full(b), Cfg::SM_THREADS);
      mbar_init(bp_empty(b), 1);
    }
    mbar_init(bo_full, 1);
    mbar_init(bo_empty, 128);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // barrier inits visible cluster-wide before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Work distribution. CL == 2: the two CTAs of a cluster take items 2m and 2m+1, which by construction
  // of the table share every streamed tile (same key range); each CTA loads half of every streamed slot
  // and multicasts it to both, halving the L2 -> SM traffic of the streamed operands.
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int worker = (int)blockIdx.x / CL, n_workers = (int)gridDim.x / CL;
  const int n_work = p.n_items / CL;
  constexpr uint16_t MC_MASK = (1u << CL) - 1;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, q_ph = 0;
      auto load_k = [&](const AttnItem& it, int j) {
#pragma unroll 1
        for (int s = 0; s < Cfg::K_SLOTS; ++s) {
          mbar_wait(kv_empty(st), ph ^ 1);
          mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
#pragma unroll
          for (int kb = 0; kb < Cfg::KB_PER_KSLOT; ++kb) {
            const uint32_t dst = sKV + st * Cfg::SLOT_BYTES + kb * 16384;
            const int c0 = it.col0 + (s * Cfg::KB_PER_KSLOT + kb) * 64, c1 = it.kv_row0 + j * 128;
            if (CL == 1) tma_load_2d(dst, &tmK, kv_full(st), c0, c1);
            else if ((kb % CL) == rank) tma_load_2d_mc(dst, &tmK, kv_full(st), c0, c1, MC_MASK);
          }
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      auto load_v = [&](const AttnItem& it, int j) {
#pragma unroll 1
        for (int s = 0; s < Cfg::V_SLOTS; ++s) {
          mbar_wait(kv_empty(st), ph ^ 1);
          mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
          // MN-major: one box of [KEYS_PER_VSLOT key rows x 64 columns] per 64-column atom
#pragma unroll
          for (int a = 0; a < Cfg::KB; ++a) {
            const uint32_t dst = sKV + st * Cfg::SLOT_BYTES + a * (Cfg::KEYS_PER_VSLOT * 128);
            const int c0 = it.col0 + a * 64, c1 = it.v_row0 + j * 128 + s * Cfg::KEYS_PER_VSLOT;
            if (CL == 1) tma_load_2d(dst, &tmV, kv_full(st), c0, c1);
            else if ((a % CL) == rank) tma_load_2d_mc(dst, &tmV, kv_full(st), c0, c1, MC_MASK);
          }
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnItem it = p.items[wk * CL + rank];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bq_empty, q_ph ^ 1);
        mbar_arrive_expect_tx(bq_full, Cfg::Q_BYTES);
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb)
          tma_load_2d(sQ + kb * 16384, &tmQ, bq_full, it.col0 + kb * 64, it.q_row0);
        q_ph ^= 1;
        // consumption order of the MMA warp: K0, [K1, V0], [K2, V1], ..., V_last
        load_k(it, 0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) load_k(it, j + 1);
          load_v(it, j);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, q_ph = 0, o_ph = 0;
      uint32_t p_ph[2] = {0, 0};
      uint32_t s_ph[2] = {0, 0};
      auto issue_qk = [&](int j, bool last) {
        const int b = j & 1;
        mbar_wait(s_empty(b), s_ph[b] ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * 128;
#pragma unroll 1
        for (int s = 0; s < Cfg::K_SLOTS; ++s) {
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < Cfg::KB_PER_KSLOT; ++kb) {
            const uint32_t a_tile = sQ + (s * Cfg::KB_PER_KSLOT + kb) * 16384;
            const uint32_t b_tile = sKV + st * Cfg::SLOT_BYTES + kb * 16384;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // forward: S[query][key] = Q(resident) K(streamed)^T ; dV mode: S[query][key] = Q(streamed) K(resident)^T
              const uint64_t da = umma_desc_sw128((MODE == 0 ? a_tile : b_tile) + k * 32, 0, 1024);
              const uint64_t db = umma_desc_sw128((MODE == 0 ? b_tile : a_tile) + k * 32, 0, 1024);
              umma_f16_ss(d_tmem, da, db, p.idesc_qk, (s | kb | k) ? 1u : 0u);
            }
          }
          if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
        umma_commit(s_full(b));
        if (last) umma_commit(bq_empty);  // every QK^T of this item has been issued: Q tile may be replaced
        s_ph[b] ^= 1;
      };
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnItem it = p.items[wk * CL + rank];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bq_full, q_ph);
        q_ph ^= 1;
        issue_qk(0, nkv == 1);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) issue_qk(j + 1, j + 2 == nkv);
          const int pb = TSP ? (j & 1) : 0;
          mbar_wait(bp_full(pb), p_ph[pb]);  // P_j is in SMEM / TMEM (and O was rescaled if needed)
          p_ph[pb] ^= 1;
          if (j == 0) mbar_wait(bo_empty, o_ph ^ 1);  // previous item's O has been read out
          tc_fence_after();
          const uint32_t o_tmem = tmem_base + Cfg::O_COL;
#pragma unroll 1
          for (int s = 0; s < Cfg::V_SLOTS; ++s) {
            mbar_wait(kv_full(st), ph);
            tc_fence_after();
            const uint32_t v_tile = sKV + st * Cfg::SLOT_BYTES;
#pragma unroll
            for (int k = 0; k < Cfg::KEYS_PER_VSLOT / 16; ++k) {
              const int key = s * Cfg::KEYS_PER_VSLOT + k * 16;  // offset along the contraction inside the 128-row tile
              // forward: A = P[query][key], K-major (contraction over keys: 32 B per step inside a swizzled row)
              // dV mode: A = P^T: the same tile [query rows][keys] read MN-major (M = keys contiguous, two
              //          64-key atoms 16 KB apart; contraction over queries: 16 rows = 2048 B per step)
              if constexpr (TSP) {   // A = P from TMEM: 16 keys = 8 columns of 16-bit pairs; keys 64-127 -> O_1
                umma_f16_ts(o_tmem + (key >> 6) * 64, tmem_base + Cfg::P_COL + pb * 64 + (key >> 1),
                            umma_desc_sw128(v_tile + k * 2048, Cfg::KEYS_PER_VSLOT * 128, 1024), p.idesc_pv,
                            (j | (key & 63)) ? 1u : 0u);
              } else {
                const uint64_t da = (MODE == 0) ? umma_desc_sw128(sP + (key >> 6) * 16384 + (key & 63) * 2, 0, 1024)
                                                : umma_desc_sw128(sP + (key >> 4) * 2048, 16384, 1024);
                umma_f16_ss(o_tmem, da, umma_desc_sw128(v_tile + k * 2048, Cfg::KEYS_PER_VSLOT * 128, 1024), p.idesc_pv,
                            (j | s | k) ? 1u : 0u);
              }
            }
            if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
            if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          }
          umma_commit(bp_empty(pb));  // P buffer free / O accumulation of tile j complete
        }
        umma_commit(bo_full);    // (same completion point as the last bp_empty)
        o_ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================================================== softmax / correction / epilogue
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;   // SM_WARPS == 8: warp q + 4 shares TMEM lane quadrant q, columns 64-127
    constexpr int COLS = Cfg::COLS;
    constexpr bool SPLIT = Cfg::SM_WARPS == 8;
    const int c_lo = half * COLS, c_hi = half * COLS + COLS;
    const int r = q * 32 + lane;  // row of the tile owned by this thread
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t s_ph[2] = {0, 0};
    uint32_t pe_ph[2] = {0, 0};
    uint32_t of_ph = 0;
    // SPLIT: combine a per-row value of the two halves (both threads of a row return the same result)
    auto row_exchange = [&](float v, bool is_max) -> float {
      xch[half * 128 + r] = v;
      asm volatile("bar.sync 2, %0;" ::"n"(Cfg::SM_THREADS) : "memory");
      const float o = xch[(half ^ 1) * 128 + r];
      asm volatile("bar.sync 2, %0;" ::"n"(Cfg::SM_THREADS) : "memory");   // the slots may be written again
      return is_max ? fmaxf(v, o) : v + o;
    };
    auto row_exchange_raw = [&](float v) -> float {   // the partner's value
      xch[half * 128 + r] = v;
      asm volatile("bar.sync 2, %0;" ::"n"(Cfg::SM_THREADS) : "memory");
      const float o = xch[(half ^ 1) * 128 + r];
      asm volatile("bar.sync 2, %0;" ::"n"(Cfg::SM_THREADS) : "memory");
      return o;
    };
    // SPLIT: OR of a predicate over every softmax thread (a retry decision must be taken by both halves of a row)
    auto any_thread = [&](bool pred) -> bool {
      if (!SPLIT) return __any_sync(0xffffffffu, pred);
      uint32_t out;
      asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, 2, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(out) : "r"((uint32_t)pred), "n"(Cfg::SM_THREADS) : "memory");
      return out != 0;
    };
    uint8_t* sP_ptr = smem + Cfg::Q_BYTES;
    constexpr float LOG2E = 1.4426950408889634f;
    // 64 columns (two tcgen05.ld in flight) -> 16-bit pairs -> the swizzled A-operand tile of the P V MMA
    auto store_p64 = [&](int c, const uint32_t (&pk)[32]) {
      const uint32_t rowp = sP + (c >> 6) * 16384 + r * 128;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int ch = t ^ (r & 7);
        sts_v4(rowp + ch * 16, pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
      }
    };
    auto pack_pair = [&](float a, float b) -> uint32_t {
      if (p.dtype == CSN_F16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    };
    auto pack_t = [](auto F16C, float a, float b) -> uint32_t {   // operand type known at compile time
      if (decltype(F16C)::value) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    };
    for (int wk = worker; wk < n_work; wk += n_workers) {
      const AttnItem it = p.items[wk * CL + rank];
      const int nkv = (it.kv_len + 127) >> 7;
      float m_used = -INFINITY;  // reference max (raw score units)
      float l = 0.f;
      for (int j = 0; j < nkv; ++j) {
        const int b = j & 1;
        mbar_wait(s_full(b), s_ph[b]);
        s_ph[b] ^= 1;
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr + b * 128;
        const int nvalid = min(128, it.kv_len - j * 128);
        const int pb = TSP ? (j & 1) : 0;
        bool waited_p = false;
        if (p.debug & 1) {
          mbar_wait(bp_empty(pb), pe_ph[pb] ^ 1);
          pe_ph[pb] ^= 1;
          l = 1.f;
        } else if (MODE == 1) {
          // ---- dV mode: lane = query row of the streamed tile, columns = the resident keys.
          //      P[query][key] = exp(s*scale - lse[query]); statistics come from the forward pass.
          const bool rvalid = r < nvalid;                       // query row exists
          const float lse_l2 = rvalid ? p.lse[it.lse_off + j * 128 + r] * LOG2E : 0.f;
          const int cvalid = rvalid ? it.q_valid : 0;           // resident key columns that exist
          const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.lse_off + j * 128 + r));
          auto dv_tile = [&](auto F16C, auto FULLC) {
            constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
#pragma unroll 1
            for (int c = c_lo; c < c_hi; c += 64) {
              uint32_t v0[32], v1[32];
              tmem_ld_32x32(s_addr + c, v0);
              tmem_ld_32x32(s_addr + c + 32, v1);
              tmem_ld_wait();
              uint32_t pk[32];
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float a0 = fast_exp2(__uint_as_float(v0[i]) * p.scale_log2 - lse_l2);
                float a1 = fast_exp2(__uint_as_float(v0[i + 1]) * p.scale_log2 - lse_l2);
                float b0 = fast_exp2(__uint_as_float(v1[i]) * p.scale_log2 - lse_l2);
                float b1 = fast_exp2(__uint_as_float(v1[i + 1]) * p.scale_log2 - lse_l2);
                if (!FULL) {
                  if (!(c + i < cvalid)) a0 = 0.f;
                  if (!(c + i + 1 < cvalid)) a1 = 0.f;
                  if (!(c + 32 + i < cvalid)) b0 = 0.f;
                  if (!(c + 33 + i < cvalid)) b1 = 0.f;
                }
                if (DROP) {   // the forward pass's dropout mask, regenerated (query row id, key column)
                  const uint32_t h0 = drop_pair(rk, (uint32_t)(it.key0 + c + i) >> 1), h1 = drop_pair(rk, (uint32_t)(it.key0 + c + 32 + i) >> 1);
                  a0 = drop_keep_lo(h0, p.drop_thresh) ? a0 * p.drop_scale : 0.f; a1 = drop_keep_hi(h0, p.drop_thresh) ? a1 * p.drop_scale : 0.f;
                  b0 = drop_keep_lo(h1, p.drop_thresh) ? b0 * p.drop_scale : 0.f; b1 = drop_keep_hi(h1, p.drop_thresh) ? b1 * p.drop_scale : 0.f;
                }
                pk[i >> 1] = pack_t(F16C, a0, a1);
                pk[16 + (i >> 1)] = pack_t(F16C, b0, b1);
              }
              if (!waited_p) {
                mbar_wait(bp_empty(pb), pe_ph[pb] ^ 1);
                pe_ph[pb] ^= 1;
                waited_p = true;
              }
              store_p64(c, pk);
            }
          };
          // compile-time variants of the loop: operand type x "every row and key column of this thread's slice exists"
          const bool full = __all_sync(0xffffffffu, cvalid >= c_hi);
          if (p.dtype == CSN_F16) {
            if (full) dv_tile(std::true_type{}, std::true_type{}); else dv_tile(std::true_type{}, std::false_type{});
          } else {
            if (full) dv_tile(std::false_type{}, std::true_type{}); else dv_tile(std::false_type{}, std::false_type{});
          }
        } else {
          if constexpr (PRIV) {
          // ---- forward, d_head 64: this thread's 64 scores go into registers at once (the S buffer is released
          //      before the arithmetic starts), their exact maximum decides about the (rare, > 2^8) re-scale of the
          //      thread's own accumulator rows; nothing is shared with the partner warp until the epilogue.
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(s_addr + c_lo, v0);
          tmem_ld_32x32(s_addr + c_lo + 32, v1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(s_empty(b));
          const int nv = nvalid - c_lo;   // valid columns among this thread's 64 (<= 0: none)
          float mx = -INFINITY;
          if (nv >= 64) {   // (warp-uniform) no padding in this thread's columns: no per-element predicates
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < nv) mx = fmaxf(mx, __uint_as_float(v0[i]));
              if (32 + i < nv) mx = fmaxf(mx, __uint_as_float(v1[i]));
            }
          }
          float alpha = 1.f;
          bool need = false;
          if (j == 0) {
            m_used = (mx == -INFINITY) ? 0.f : mx;   // (no valid column at all in this half: P = exp2(-inf) = 0, l stays 0)
          } else if ((mx - m_used) * p.scale_log2 > 8.f) {
            alpha = fast_exp2((m_used - mx) * p.scale_log2);
            m_used = mx;
            need = true;
          }
          if (__any_sync(0xffffffffu, need)) {
            // the accumulator may still be receiving P V of the previous tile (P is double-buffered): wait for it
            if (j > 0) mbar_wait(bp_empty(pb ^ 1), pe_ph[pb ^ 1] ^ 1);
            tc_fence_after();
            const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + half * 64;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
              uint32_t v[32];
              tmem_ld_32x32(o_addr + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_32x32(o_addr + c, v);
            }
            tmem_st_wait();
            l *= alpha;
          }
          const float moff = m_used * p.scale_log2;
          const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.lse_off + r));
          float lsum = 0.f;
          uint32_t pk[32];
          auto exp_regs = [&](auto F16C, auto FULLC) {
            constexpr bool FULL = decltype(FULLC)::value;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a0 = fast_exp2(__uint_as_float(v0[i]) * p.scale_log2 - moff);
              float a1 = fast_exp2(__uint_as_float(v0[i + 1]) * p.scale_log2 - moff);
              float b0 = fast_exp2(__uint_as_float(v1[i]) * p.scale_log2 - moff);
              float b1 = fast_exp2(__uint_as_float(v1[i + 1]) * p.scale_log2 - moff);
              if (!FULL) {
                if (!(i < nv)) a0 = 0.f;
                if (!(i + 1 < nv)) a1 = 0.f;
                if (!(32 + i < nv)) b0 = 0.f;
                if (!(33 + i < nv)) b1 = 0.f;
              }
              lsum += (a0 + a1) + (b0 + b1);   // the softmax denominator sees every key; dropout acts on the result
              if (DROP) {
                const uint32_t h0 = drop_pair(rk, (uint32_t)(j * 128 + c_lo + i) >> 1), h1 = drop_pair(rk, (uint32_t)(j * 128 + c_lo + 32 + i) >> 1);
                a0 = drop_keep_lo(h0, p.drop_thresh) ? a0 * p.drop_scale : 0.f; a1 = drop_keep_hi(h0, p.drop_thresh) ? a1 * p.drop_scale : 0.f;
                b0 = drop_keep_lo(h1, p.drop_thresh) ? b0 * p.drop_scale : 0.f; b1 = drop_keep_hi(h1, p.drop_thresh) ? b1 * p.drop_scale : 0.f;
              }
              pk[i >> 1] = pack_t(F16C, a0, a1);
              pk[16 + (i >> 1)] = pack_t(F16C, b0, b1);
            }
          };
          if (p.dtype == CSN_F16) {
            if (nv >= 64) exp_regs(std::true_type{}, std::true_type{}); else exp_regs(std::true_type{}, std::false_type{});
          } else {
            if (nv >= 64) exp_regs(std::false_type{}, std::true_type{}); else exp_regs(std::false_type{}, std::false_type{});
          }
          mbar_wait(bp_empty(pb), pe_ph[pb] ^ 1);   // the P buffer may be overwritten once the P V MMAs that read it have completed
          pe_ph[pb] ^= 1;
          tc_fence_after();
          tmem_st_32x32(tmem_base + lane_addr + Cfg::P_COL + pb * 64 + (c_lo >> 1), pk);
          l += lsum;
          } else {
          // ---- forward: online softmax. Tile 0 takes the exact row max first; later tiles are
          //      exponentiated optimistically against the running reference max and only redone
          //      (two-pass + rescale of O) if some score exceeds it by more than 2^8.
          bool two_pass = (j == 0);
          float alpha = 1.f;
          bool need = false;
          float lsum;
          for (;;) {
            if (two_pass) {
              float mx = -INFINITY;
#pragma unroll 1
              for (int c = c_lo; c < c_hi; c += 64) {
                uint32_t v0[32], v1[32];
                tmem_ld_32x32(s_addr + c, v0);
                tmem_ld_32x32(s_addr + c + 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  if (c + i < nvalid) mx = fmaxf(mx, __uint_as_float(v0[i]));
                  if (c + 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(v1[i]));
                }
              }
              if (SPLIT) mx = row_exchange(mx, true);
              if (j == 0) {
                m_used = mx;
              } else if ((mx - m_used) * p.scale_log2 > 8.f) {
                alpha = fast_exp2((m_used - mx) * p.scale_log2);
                m_used = mx;
                need = true;
              }
            }
            const float moff = m_used * p.scale_log2;
            const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.lse_off + r));
            lsum = 0.f;
            float cmax = -INFINITY;
            auto exp_tile = [&](auto F16C, auto FULLC) {
              constexpr bool FULL = decltype(FULLC)::value;
#pragma unroll 1
              for (int c = c_lo; c < c_hi; c += 64) {
                uint32_t v0[32], v1[32];
                tmem_ld_32x32(s_addr + c, v0);
                tmem_ld_32x32(s_addr + c + 32, v1);
                tmem_ld_wait();
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float s0 = (FULL || c + i < nvalid) ? __uint_as_float(v0[i]) : -INFINITY;
                  const float s1 = (FULL || c + i + 1 < nvalid) ? __uint_as_float(v0[i + 1]) : -INFINITY;
                  const float t0 = (FULL || c + 32 + i < nvalid) ? __uint_as_float(v1[i]) : -INFINITY;
                  const float t1 = (FULL || c + 33 + i < nvalid) ? __uint_as_float(v1[i + 1]) : -INFINITY;
                  cmax = fmaxf(cmax, fmaxf(fmaxf(s0, s1), fmaxf(t0, t1)));
                  float a0 = fast_exp2(s0 * p.scale_log2 - moff), a1 = fast_exp2(s1 * p.scale_log2 - moff);
                  float b0 = fast_exp2(t0 * p.scale_log2 - moff), b1 = fast_exp2(t1 * p.scale_log2 - moff);
                  lsum += (a0 + a1) + (b0 + b1);   // the softmax denominator sees every key; dropout acts on the result
                  if (DROP) {
                    const uint32_t h0 = drop_pair(rk, (uint32_t)(j * 128 + c + i) >> 1), h1 = drop_pair(rk, (uint32_t)(j * 128 + c + 32 + i) >> 1);
                    a0 = drop_keep_lo(h0, p.drop_thresh) ? a0 * p.drop_scale : 0.f; a1 = drop_keep_hi(h0, p.drop_thresh) ? a1 * p.drop_scale : 0.f;
                    b0 = drop_keep_lo(h1, p.drop_thresh) ? b0 * p.drop_scale : 0.f; b1 = drop_keep_hi(h1, p.drop_thresh) ? b1 * p.drop_scale : 0.f;
                  }
                  pk[i >> 1] = pack_t(F16C, a0, a1);
                  pk[16 + (i >> 1)] = pack_t(F16C, b0, b1);
                }
                if (!waited_p) {
                  // the P buffer may be overwritten once the P V MMAs that read it have completed
                  mbar_wait(bp_empty(pb), pe_ph[pb] ^ 1);
                  pe_ph[pb] ^= 1;
                  waited_p = true;
                  if (TSP) tc_fence_after();
                }
                if constexpr (TSP) tmem_st_32x32(tmem_base + lane_addr + Cfg::P_COL + pb * 64 + (c >> 1), pk);
                else store_p64(c, pk);
              }
            };
            if (p.dtype == CSN_F16) {
              if (nvalid == 128) exp_tile(std::true_type{}, std::true_type{}); else exp_tile(std::true_type{}, std::false_type{});
            } else {
              if (nvalid == 128) exp_tile(std::false_type{}, std::true_type{}); else exp_tile(std::false_type{}, std::false_type{});
            }
            const bool exceeded = !two_pass && (cmax - m_used) * p.scale_log2 > 8.f;
            // (SPLIT: tile 0 is two-pass for everybody, so the barrier inside any_thread is reached by all or none)
            if (two_pass || !any_thread(exceeded)) break;
            two_pass = true;  // rare: redo this tile against the new maximum
          }
          tc_fence_after();
          if (half == 0 && __any_sync(0xffffffffu, need)) {
            // rescale this warp's 32 rows of O in TMEM (rare); P V of the previous tile has completed
            const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL;
#pragma unroll 1
            for (int c = 0; c < DH; c += 32) {
              uint32_t v[32];
              tmem_ld_32x32(o_addr + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_32x32(o_addr + c, v);
            }
            tmem_st_wait();
          }
          if (need) l *= alpha;
          l += lsum;
          }
        }
        // S_j fully consumed; P_j visible to the tensor-core (async) proxy
        if (TSP) tmem_st_wait();
        tc_fence_before();
        if (!PRIV) mbar_arrive(s_empty(b));
        if (!TSP) fence_proxy_async_smem();
        mbar_arrive(bp_full(pb));
      }
      // ---- epilogue: O / l -> 16-bit, LSE   (dV mode: the accumulator as is)
      mbar_wait(bo_full, of_ph);
      of_ph ^= 1;
      tc_fence_after();
      float a_me = 1.f, a_ot = 0.f;   // PRIV: weights of this thread's and the partner's partial results
      if (SPLIT) {
        if (PRIV) {
          // merge the two column halves: reference max m = max over the halves that saw a key, weights 2^((m_h - m) scale)
          const float l_ot = row_exchange_raw(l);
          const float m_me = l > 0.f ? m_used : -INFINITY;
          const float m_ot = row_exchange_raw(m_me);
          const float m = fmaxf(m_me, m_ot);
          a_me = l > 0.f ? fast_exp2((m_me - m) * p.scale_log2) : 0.f;
          a_ot = l_ot > 0.f ? fast_exp2((m_ot - m) * p.scale_log2) : 0.f;
          l = l * a_me + l_ot * a_ot;
          m_used = m;
        } else if (MODE == 0) {
          l = row_exchange(l, false);
        }
        if (half != 0) {
          // warps 4-7 write the output through the P tile: nobody may refill it before their slabs have left
          asm volatile("bar.sync 3, %0;" ::"n"(Cfg::SM_THREADS) : "memory");
          continue;
        }
      }
      const float inv_l = (MODE == 1) ? 1.f : 1.f / l;
      const bool valid = r < it.q_valid;
      const bool want_lo = (MODE == 0) && p.Olo != nullptr;
      const float lo_scale = p.dtype == CSN_F16 ? 2048.f : 256.f;
      const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL;
      // Coalesced output: every warp stages [its 32 rows x 64 columns] slabs in ITS OWN rows of the (now idle)
      // P tile — 128B-swizzled, conflict-free 16-byte writes — and hands them to the TMA store engine; two
      // slabs per warp are in flight. Rows >= q_valid are written as zeros (outputs live in padded layouts).
      int slab = 0;
#pragma unroll 1
      for (int c = 0; c < ((p.debug & 2) ? 0 : DH); c += 64) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(o_addr + c, v0);
        tmem_ld_32x32(o_addr + c + 32, v1);
        tmem_ld_wait();
        if constexpr (PRIV) {   // O = O_0 a_0 + O_1 a_1 (this thread is in half 0: O_0 is its own)
          uint32_t w0[32], w1[32];
          tmem_ld_32x32(o_addr + 64, w0);
          tmem_ld_32x32(o_addr + 96, w1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = __float_as_uint(__uint_as_float(v0[i]) * a_me + __uint_as_float(w0[i]) * a_ot);
            v1[i] = __float_as_uint(__uint_as_float(v1[i]) * a_me + __uint_as_float(w1[i]) * a_ot);
          }
        }
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = valid ? __uint_as_float(v0[i]) * inv_l : 0.f, a1 = valid ? __uint_as_float(v0[i + 1]) * inv_l : 0.f;
          const float b0 = valid ? __uint_as_float(v1[i]) * inv_l : 0.f, b1 = valid ? __uint_as_float(v1[i + 1]) * inv_l : 0.f;
          hi[i >> 1] = pack_pair(a0, a1);
          hi[16 + (i >> 1)] = pack_pair(b0, b1);
          if (want_lo) {
            float2 ha, hb;
            if (p.dtype == CSN_F16) {
              ha = __half22float2(*reinterpret_cast<__half2*>(&hi[i >> 1]));
              hb = __half22float2(*reinterpret_cast<__half2*>(&hi[16 + (i >> 1)]));
            } else {
              ha = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&hi[i >> 1]));
              hb = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&hi[16 + (i >> 1)]));
            }
            lo[i >> 1] = pack_pair((a0 - ha.x) * lo_scale, (a1 - ha.y) * lo_scale);
            lo[16 + (i >> 1)] = pack_pair((b0 - hb.x) * lo_scale, (b1 - hb.y) * lo_scale);
          }
        }
#pragma unroll 1
        for (int which = 0; which < (want_lo ? 2 : 1); ++which) {
          uint32_t buf;
          if (Cfg::XSTG_BYTES > 0) {   // four slabs per warp: two in its rows of the P tile, two in the extra region
            const int bi = slab & 3;
            buf = (bi < 2 ? sP + bi * 16384 : sX + (bi - 2) * 16384) + q * 4096;
            if (lane == 0) tma_store_wait_read<3>();
          } else {
            buf = sP + (slab & 1) * 16384 + q * 4096;
            if (lane == 0) tma_store_wait_read<1>();
          }
          __syncwarp();
          const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const uint32_t a = rowaddr + (((uint32_t)t ^ ((uint32_t)lane & 7u)) << 4);
            if (which == 0)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hi[4 * t]), "r"(hi[4 * t + 1]), "r"(hi[4 * t + 2]), "r"(hi[4 * t + 3]) : "memory");
            else
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(lo[4 * t]), "r"(lo[4 * t + 1]), "r"(lo[4 * t + 2]), "r"(lo[4 * t + 3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_2d(which == 0 ? &tmO : &tmOlo, buf, it.col0 + c, it.o_row0 + q * 32);
            tma_store_commit();
          }
          ++slab;
        }
      }
      if (lane == 0) tma_store_wait_read<0>();   // this warp's rows of the P tile are written again by the next item
      __syncwarp();
      if (MODE == 0 && p.lse) p.lse[it.lse_off + r] = valid ? (m_used * p.scale + __logf(l)) : 0.f;
      tc_fence_before();
      mbar_arrive(bo_empty);
      if (SPLIT) asm volatile("bar.sync 3, %0;" ::"n"(Cfg::SM_THREADS) : "memory");
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA exits while its peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DH, int MODE, int CL, bool DROP>
static int launch_attn_fwd_d(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                           const CUtensorMap& tmO, const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  using Cfg = AttnCfg<DH>;
  auto kern = attn_fwd_kernel<DH, MODE, CL, DROP>;
  CSN_SET_MAX_SMEM(kern, Cfg::SMEM_BYTES);
  const int n_work = a.n_items / CL;
  const int workers = num_sms() / CL;
  const int grid = (n_work < workers ? n_work : workers) * CL;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmQ, tmK, tmV, tmO, tmOlo, a));
  CSN_LAUNCH_OK("attn_fwd_kernel");
  return 0;
}

// the dropout code is a compile-time variant: the eval-mode kernels carry none of it
template <int DH, int MODE, int CL>
static int launch_attn_fwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                           const CUtensorMap& tmO, const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  if (a.drop_thresh) return launch_attn_fwd_d<DH, MODE, CL, true>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
  return launch_attn_fwd_d<DH, MODE, CL, false>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
}

// Items 2m / 2m+1 may share a cluster when they stream exactly the same tiles.
static bool items_pairable(int n_items, int d_head, int paired_flag) {
  return paired_flag && (n_items % 2 == 0) && d_head == 256;
}

}  // namespace csn

static int attn_launch_common(int mode, const void* Q, const void* K, const void* V, int64_t q_rows, int64_t kv_rows,
                              int64_t v_rows, int64_t width, int64_t ldq, int64_t ldk, int64_t ldv, int32_t d_head, int32_t dtype,
                              const int32_t* items, int32_t n_items, void* O, int64_t o_rows, int64_t ldo, float* lse,
                              void* Olo, int32_t paired, uint32_t drop_seed, float drop_p, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Q && K && V && items && O, "csn_attn_fwd: null pointer");
  CSN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "csn_attn_fwd: dropout probability %f outside [0, 1)", (double)drop_p);
  CSN_CHECK_ARG(d_head == 256 || d_head == 64, "csn_attn_fwd: d_head=%d not supported (64 or 256)", d_head);
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "csn_attn_fwd: 16-bit operands only");
  CSN_CHECK_ARG((ldo * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "csn_attn_fwd: O not 16B aligned");
  if (n_items == 0) return 0;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_2d(&tmQ, Q, dtype, width, q_rows, ldq, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmK, K, dtype, width, kv_rows, ldk, 64, 128);
  if (rc) return rc;
  const uint32_t vrows = d_head == 256 ? 64 : 128;  // KEYS_PER_VSLOT
  rc = make_tmap_2d(&tmV, V, dtype, width, v_rows, ldv, 64, vrows);
  if (rc) return rc;
  AttnFwdArgs a;
  a.items = reinterpret_cast<const AttnItem*>(items);
  a.n_items = n_items;
  a.O = O;
  a.Olo = Olo;
  a.ldo = ldo;
  a.lse = lse;
  a.scale = 1.0f / sqrtf((float)d_head);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.dtype = dtype;
  a.drop_seed = drop_seed;
  a.drop_thresh = drop_thresh16(drop_p);
  a.drop_scale = drop_scale_of(a.drop_thresh);
  a.drop_epoch = a.drop_thresh ? drop_epoch_ptr() : nullptr;
  {
    const char* dbg = getenv("CSN_ATTN_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
  }
  const uint32_t fmt = dtype == CSN_F16 ? 0u : 1u;
  a.idesc_qk = umma_idesc_f16(fmt, 0, 0, 128);
  a.idesc_pv = umma_idesc_f16(fmt, mode == 1 ? 1u : 0u, 1, (uint32_t)d_head);   // dV mode reads P^T (MN-major A)
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // outputs leave through TMA stores of [32 rows x 64 columns] slabs
  CUtensorMap tmO, tmOlo;
  rc = make_tmap_2d(&tmO, O, dtype, width, o_rows, ldo, 64, 32);
  if (rc) return rc;
  tmOlo = tmO;
  if (Olo) {
    rc = make_tmap_2d(&tmOlo, Olo, dtype, width, o_rows, ldo, 64, 32);
    if (rc) return rc;
  }
  const bool pair = items_pairable(n_items, d_head, paired);
  if (mode == 1) CSN_CHECK_ARG(lse != nullptr, "csn_attn_bwd_dv: lse is required");
  // [128 x 256] score tiles (attn_wide.cu): bit 0 = forward, bit 1 = dV.  Default: dV only (335 vs 363 us on the
  // config-2 step); the wide forward kernel is correct but no faster than this file's (405 vs 406 us: its second
  // output, the rounding residual of O, needs four staging slabs per warp where the P tile offers two).
  // Unset: dV always; the forward pass when it has a single output (V centred: no rounding residual of O).
  static const int wide_env = getenv("CSN_ATTN_WIDE") == nullptr ? -1 : atoi(getenv("CSN_ATTN_WIDE"));
  const int wide = wide_env >= 0 ? wide_env : (2 | (Olo == nullptr ? 1 : 0));
  if (d_head == 256 && ((wide >> mode) & 1)) {
    a.idesc_qk = umma_idesc_f16(fmt, 0, 0, 256);
    a.idesc_pv = umma_idesc_f16(fmt, 0, 1, 256);
    return launch_attn_wide(mode, pair, tmQ, tmK, tmV, tmO, tmOlo, a, s);
  }
  if (mode == 0) {
    if (d_head == 256) return pair ? launch_attn_fwd<256, 0, 2>(tmQ, tmK, tmV, tmO, tmOlo, a, s) : launch_attn_fwd<256, 0, 1>(tmQ, tmK, tmV, tmO, tmOlo, a, s);
    static const int fwd64 = getenv("CSN_FWD64") == nullptr ? 1 : atoi(getenv("CSN_FWD64"));
    if (fwd64) return launch_attn_fwd64(tmQ, tmK, tmV, tmO, tmOlo, a, s);   // 16 softmax warps (attn_fwd64.cu)
    return launch_attn_fwd<64, 0, 1>(tmQ, tmK, tmV, tmO, tmOlo, a, s);
  }
  CSN_CHECK_ARG(lse != nullptr, "csn_attn_bwd_dv: lse is required");
  if (d_head == 256) return pair ? launch_attn_fwd<256, 1, 2>(tmQ, tmK, tmV, tmO, tmOlo, a, s) : launch_attn_fwd<256, 1, 1>(tmQ, tmK, tmV, tmO, tmOlo, a, s);
  return launch_attn_fwd<64, 1, 1>(tmQ, tmK, tmV, tmO, tmOlo, a, s);
}

extern "C" int csn_attn_fwd(const void* Q, const void* K, const void* V, int64_t q_rows, int64_t kv_rows,
                            int64_t width, int64_t ldq, int64_t ldk, int64_t ldv, int32_t d_head, int32_t dtype,
                            const int32_t* items, int32_t n_items, void* O, int64_t o_rows, int64_t ldo, float* lse,
                            void* O_lo, int32_t paired, uint32_t drop_seed, float drop_p, void* stream) {
  return attn_launch_common(0, Q, K, V, q_rows, kv_rows, kv_rows, width, ldq, ldk, ldv, d_head, dtype, items, n_items,
                            O, o_rows, ldo, lse, O_lo, paired, drop_seed, drop_p, stream);
}

// dV = P^T dO with P^T recomputed from K, Q and the forward log-sum-exp. Same pipeline as the forward
// kernel with the roles swapped: resident tile = 128 KEY rows (argument Kres), streamed K-major
// operand = the QUERY rows (Qstr), streamed MN-major operand = dO rows; items[*].lse_off indexes the
// lse of the first streamed (query) row.
extern "C" int csn_attn_bwd_dv(const void* Kres, const void* Qstr, const void* dO, int64_t k_rows, int64_t q_rows,
                               int64_t do_rows, int64_t width, int64_t ldk, int64_t ldq, int64_t lddo, int32_t d_head, int32_t dtype,
                               const int32_t* items, int32_t n_items, void* dV, int64_t dv_rows, int64_t lddv,
                               const float* lse, int32_t paired, uint32_t drop_seed, float drop_p, void* stream) {
  return attn_launch_common(1, Kres, Qstr, dO, k_rows, q_rows, do_rows, width, ldk, ldq, lddo, d_head, dtype, items,
                            n_items, dV, dv_rows, lddv, const_cast<float*>(lse), nullptr, paired, drop_seed, drop_p, stream);
}
