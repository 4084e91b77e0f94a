// Fused attention for d_head = 256 with 256-wide score tiles (forward and the dV half of the backward pass).
//
// An SS-mode M=128 tcgen05.mma costs about the same for N = 128 and N = 256 (profiles/r1_experiments.md), so the
// 128-key score tiles of attn_fwd.cu run the Q K^T half of the work at half rate.  Here the score tile is
// [128 x 256]: TMEM = S (256 columns, single buffer) | accumulator (256 columns).
//
//   MODE 0 (forward, csa_models.py:138-144):  resident tile R = 128 query rows;   per step of 256 keys
//        S = R K^T  ->  online softmax  ->  P (16-bit, SMEM)  ->  O += P V        (V consumed MN-major)
//   MODE 1 (dV = P^T dO):                      resident tile R = 128 KEY rows;     per step of 256 queries
//        S^T = R Q^T  ->  P^T = exp(S^T*scale - lse[query])  ->  dV += P^T dO     (dO consumed MN-major)
//
// CTA = 384 threads, persistent: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11
// softmax / epilogue — the warp pair (q, q+4) shares TMEM lane quadrant q and splits the 256 columns (row
// statistics are exchanged through SMEM + a 64-thread named barrier).
// Schedule over the global sequence of steps g:  the MMA warp issues  S_g, PV_{g-1}, S_{g+1}, PV_g, ...  — S_{g+1}
// only needs S_g to have been read (P_g is kept as 16-bit pairs in registers until P V_{g-1} has released the
// P tile), so the tensor pipe always has the other kind of MMA to run while the softmax warps work; the
// epilogue of an item runs after the first score tile of the NEXT item has been read, under that item's S_1.
// SMEM: R 64 KB | P 64 KB | 3 x 32 KB ring ([256 x 64] K-major slices for S, [64 x 256] MN-major slices for P V).
#include <stdlib.h>
#include <type_traits>

#include "attn_common.cuh"

namespace csn {

struct WideAttnCfg {
  static constexpr int R_BYTES = 128 * 256 * 2;
  static constexpr int P_BYTES = 128 * 256 * 2;
  static constexpr int SLOT_BYTES = 32768;
  static constexpr int NST = 3;
  static constexpr int XCH_BYTES = 128 * 2 * 4;
  static constexpr int BAR_BYTES = 128;
  static constexpr int SMEM_BYTES = R_BYTES + P_BYTES + NST * SLOT_BYTES + XCH_BYTES + BAR_BYTES + 1024;
};

template <int MODE, int CL, bool DROP>
__global__ void __launch_bounds__(384, 1)
attn_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmOlo, const __grid_constant__ AttnFwdArgs p) {
  using Cfg = WideAttnCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sR = smem_u32(smem);
  const uint32_t sP = sR + Cfg::R_BYTES;
  const uint32_t sKV = sP + Cfg::P_BYTES;
  const uint32_t sX = sKV + Cfg::NST * Cfg::SLOT_BYTES;
  const uint32_t bar_base = sX + Cfg::XCH_BYTES;
  float* xch = reinterpret_cast<float*>(smem + Cfg::R_BYTES + Cfg::P_BYTES + Cfg::NST * Cfg::SLOT_BYTES);
  uint8_t* bar_ptr = reinterpret_cast<uint8_t*>(xch) + Cfg::XCH_BYTES;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t r_full = bar_base + 8u * (2 * Cfg::NST + 0), r_empty = bar_base + 8u * (2 * Cfg::NST + 1);
  const uint32_t s_full = bar_base + 8u * (2 * Cfg::NST + 2), s_empty = bar_base + 8u * (2 * Cfg::NST + 3);
  const uint32_t p_full = bar_base + 8u * (2 * Cfg::NST + 4), p_empty = bar_base + 8u * (2 * Cfg::NST + 5);
  const uint32_t o_full = bar_base + 8u * (2 * Cfg::NST + 6), o_empty = bar_base + 8u * (2 * Cfg::NST + 7);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 8));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dseed = DROP ? drop_seed_eff(p.drop_seed, p.drop_epoch) : 0u;   // (one load per thread, train-mode variants only)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::NST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), CL); }
    mbar_init(r_full, 1); mbar_init(r_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 256);
    mbar_init(p_full, 256); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 256);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int worker = (int)blockIdx.x / CL, n_workers = (int)gridDim.x / CL;
  const int n_work = p.n_items / CL;
  constexpr uint16_t MC_MASK = (1u << CL) - 1;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, r_ph = 0;
      auto load_s = [&](int col0, int row0) {   // [256 rows x 256 cols] as 4 K-major slices of [256 x 64]
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(kv_empty(st), ph ^ 1);
          mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
          const uint32_t dst = sKV + st * Cfg::SLOT_BYTES;
          if (CL == 1) {
            tma_load_2d(dst, &tmK, kv_full(st), col0 + kb * 64, row0);
            tma_load_2d(dst + 16384, &tmK, kv_full(st), col0 + kb * 64, row0 + 128);
          } else {
            tma_load_2d_mc(dst + rank * 16384, &tmK, kv_full(st), col0 + kb * 64, row0 + rank * 128, MC_MASK);
          }
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      auto load_pv = [&](int col0, int row0) {   // [256 rows x 256 cols] as 4 MN-major slices of [64 rows x 256 cols]
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          mbar_wait(kv_empty(st), ph ^ 1);
          mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
#pragma unroll
          for (int a = 0; a < 4; ++a) {   // one box of [64 rows x 64 columns] per 64-column atom
            const uint32_t dst = sKV + st * Cfg::SLOT_BYTES + a * 8192;
            if (CL == 1) tma_load_2d(dst, &tmV, kv_full(st), col0 + a * 64, row0 + s * 64);
            else if ((a % CL) == rank) tma_load_2d_mc(dst, &tmV, kv_full(st), col0 + a * 64, row0 + s * 64, MC_MASK);
          }
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      bool pending = false;
      int pv_col0 = 0, pv_row0 = 0;
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnItem it = p.items[wk * CL + rank];
        const int nst = (it.kv_len + 255) >> 8;
        mbar_wait(r_empty, r_ph ^ 1);
        mbar_arrive_expect_tx(r_full, Cfg::R_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sR + kb * 16384, &tmQ, r_full, it.col0 + kb * 64, it.q_row0);
        r_ph ^= 1;
        for (int i = 0; i < nst; ++i) {
          load_s(it.col0, it.kv_row0 + i * 256);
          if (pending) load_pv(pv_col0, pv_row0);
          pending = true;
          pv_col0 = it.col0;
          pv_row0 = it.v_row0 + i * 256;
        }
      }
      if (pending) load_pv(pv_col0, pv_row0);
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, r_ph = 0, s_ph = 0, p_ph = 0, o_ph = 0;
      auto issue_pv = [&](bool first, bool last) {
        mbar_wait(p_full, p_ph);   // P is in SMEM (and the accumulator was rescaled if needed)
        p_ph ^= 1;
        if (first) mbar_wait(o_empty, o_ph ^ 1);   // the previous item's accumulator has been read out
        tc_fence_after();
        const uint32_t o_tmem = tmem_base + 256;
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          const uint32_t v_tile = sKV + st * Cfg::SLOT_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 64 contraction rows in steps of 16
            umma_f16_ss(o_tmem, umma_desc_sw128(sP + s * 16384 + k * 32, 0, 1024),
                        umma_desc_sw128(v_tile + k * 2048, 8192, 1024), p.idesc_pv, (first && s == 0 && k == 0) ? 0u : 1u);
          if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
        umma_commit(p_empty);
        if (last) { umma_commit(o_full); o_ph ^= 1; }
      };
      bool pending = false, pv_first = false, pv_last = false;
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnItem it = p.items[wk * CL + rank];
        const int nst = (it.kv_len + 255) >> 8;
        mbar_wait(r_full, r_ph);
        r_ph ^= 1;
        for (int i = 0; i < nst; ++i) {
          mbar_wait(s_empty, s_ph ^ 1);   // the previous score tile has been read
          tc_fence_after();
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(kv_full(st), ph);
            tc_fence_after();
            const uint32_t a_tile = sR + kb * 16384;
            const uint32_t b_tile = sKV + st * Cfg::SLOT_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem_base, umma_desc_sw128(a_tile + k * 32, 0, 1024), umma_desc_sw128(b_tile + k * 32, 0, 1024),
                          p.idesc_qk, (kb | k) ? 1u : 0u);
            if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
            if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          }
          umma_commit(s_full);
          s_ph ^= 1;
          if (i == nst - 1) umma_commit(r_empty);   // every S MMA of this item has been issued
          if (pending) issue_pv(pv_first, pv_last);
          pending = true;
          pv_first = (i == 0);
          pv_last = (i == nst - 1);
        }
      }
      if (pending) issue_pv(pv_first, pv_last);
    }
  } else if (warp >= 4) {
    // ================================================================== softmax / epilogue
    const int ew = warp - 4;
    const int q = warp & 3, half = ew >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint8_t* sP_ptr = smem + Cfg::R_BYTES;
    uint32_t s_ph = 0, pe_ph = 0, of_ph = 0;
    constexpr float LOG2E = 1.4426950408889634f;
    const bool f16 = p.dtype == CSN_F16;
    auto pack_pair = [&](float a, float b) -> uint32_t {
      if (f16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    // exchange one float with the warp that owns the other 128 columns of the same rows
    auto exchange = [&](float mine) -> float {
      xch[r * 2 + half] = mine;
      pair_sync();
      const float other = xch[r * 2 + (half ^ 1)];
      pair_sync();   // the slot may be rewritten
      return other;
    };
    // ---- epilogue of a finished item: accumulator columns [half*128, half*128+128) of this warp's 32 rows
    auto epilogue = [&](const AttnItem& it, float m_used, float l_half) {
      mbar_wait(o_full, of_ph);
      of_ph ^= 1;
      tc_fence_after();
      float inv_l = 1.f, l_tot = 1.f;
      if (MODE == 0) {
        l_tot = l_half + exchange(l_half);
        inv_l = 1.f / l_tot;
      }
      const bool valid = r < it.q_valid;
      const bool want_lo = (MODE == 0) && p.Olo != nullptr && !(p.debug & 4);
      const float lo_scale = f16 ? 2048.f : 256.f;
      const uint32_t o_addr = tmem_base + lane_addr + 256 + half * 128;
      // staging: this warp's own 32 rows of its two P slices (idle: P V of the item's last step has completed)
      int slab = 0;
#pragma unroll 1
      for (int c = 0; c < 128; c += 64) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t v[32];
          tmem_ld_32x32(o_addr + c + hc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = valid ? __uint_as_float(v[i]) * inv_l : 0.f, a1 = valid ? __uint_as_float(v[i + 1]) * inv_l : 0.f;
            const uint32_t h = pack_pair(a0, a1);
            hi[hc * 16 + (i >> 1)] = h;
            if (want_lo) {
              float2 hf;
              if (f16) hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
              else hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h));
              lo[hc * 16 + (i >> 1)] = pack_pair((a0 - hf.x) * lo_scale, (a1 - hf.y) * lo_scale);
            }
          }
        }
#pragma unroll 1
        for (int which = 0; which < (want_lo ? 2 : 1); ++which) {
          const uint32_t buf = sP + (half * 2 + (slab & 1)) * 16384 + q * 4096;
          if (lane == 0) tma_store_wait_read<1>();
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
          if (lane == 0) {
            tma_store_2d(which == 0 ? &tmO : &tmOlo, buf, it.col0 + half * 128 + c, it.o_row0 + q * 32);
            tma_store_commit();
          }
          ++slab;
        }
      }
      if (lane == 0) tma_store_wait_read<0>();   // these rows of the P tile are written again by the next P
      __syncwarp();
      if (MODE == 0 && p.lse && half == 0) p.lse[it.lse_off + r] = valid ? (m_used * p.scale + __logf(l_tot)) : 0.f;
      tc_fence_before();
      mbar_arrive(o_empty);
    };

    bool have_prev = false;
    AttnItem prev;
    float prev_m = 0.f, prev_l = 0.f;
    float stat_next = 0.f;   // MODE 1: lse of query column (ew * 32 + lane) of the NEXT step, fetched one step ahead
    if (MODE == 1 && worker < n_work) {
      const AttnItem f = p.items[worker * CL + rank];
      if (ew * 32 + lane < f.kv_len) stat_next = __ldg(p.lse + f.lse_off + ew * 32 + lane);
    }
    for (int wk = worker; wk < n_work; wk += n_workers) {
      const AttnItem it = p.items[wk * CL + rank];
      const int nst = (it.kv_len + 255) >> 8;
      float m_used = -INFINITY, l = 0.f;
      for (int i = 0; i < nst; ++i) {
        const int col_first = i * 256 + half * 128;                 // first streamed row (key / query) of this warp's columns
        const int nvalid = min(128, it.kv_len - col_first);         // columns [0, nvalid) exist (may be <= 0)
        mbar_wait(s_full, s_ph);
        s_ph ^= 1;
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr + half * 128;
        uint32_t pk[64];
        float lsum = 0.f, alpha = 1.f;
        bool need = false;
        const bool skip = (p.debug & 1) != 0;   // diagnostics: no softmax arithmetic (P = 0)
        if (MODE == 0 && !skip) {
          // ---- pass 1: exact row maximum over the 256 columns of the step (both halves)
          float mx = -INFINITY;
          {
            uint32_t va[32], vb[32];
            tmem_ld_32x32(s_addr, va);
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
              tmem_ld_wait();
              tmem_ld_32x32(s_addr + (c + 1) * 32, vb);
#pragma unroll
              for (int k = 0; k < 32; ++k) if (c * 32 + k < nvalid) mx = fmaxf(mx, __uint_as_float(va[k]));
              tmem_ld_wait();
              if (c + 2 < 4) tmem_ld_32x32(s_addr + (c + 2) * 32, va);
#pragma unroll
              for (int k = 0; k < 32; ++k) if ((c + 1) * 32 + k < nvalid) mx = fmaxf(mx, __uint_as_float(vb[k]));
            }
          }
          const float mrow = fmaxf(mx, exchange(mx));
          if (i == 0) {
            m_used = mrow;
          } else if ((mrow - m_used) * p.scale_log2 > 8.f) {   // lazy rescale: only when the maximum grows by > 2^8
            alpha = fast_exp2((m_used - mrow) * p.scale_log2);
            m_used = mrow;
            need = true;
          }
        }
        if (MODE == 1) {
          // the step's 256 per-column statistics (log-sum-exp of the streamed query rows, in log2 units) go through
          // SMEM once: thread t of the 256 softmax threads fetches column t, everybody reads them back as broadcasts
          // (this thread's value was fetched one step ahead: its global-load latency is off the critical path)
          const int t = ew * 32 + lane;
          const float mine = stat_next * LOG2E;
          {   // prefetch for the next step: the same item's next 256 queries, or the first 256 of the next item
            stat_next = 0.f;
            if (i + 1 < nst) {
              const int col = (i + 1) * 256 + t;
              if (col < it.kv_len) stat_next = __ldg(p.lse + it.lse_off + col);
            } else if (wk + n_workers < n_work) {
              const AttnItem nx = p.items[(wk + n_workers) * CL + rank];
              if (t < nx.kv_len) stat_next = __ldg(p.lse + nx.lse_off + t);
            }
          }
          asm volatile("bar.sync 5, 256;" ::: "memory");   // the previous step's values have been consumed
          xch[t] = mine;
          asm volatile("bar.sync 5, 256;" ::: "memory");
        }
        // ---- pass 2: probabilities as 16-bit pairs in registers
        if (skip) {
#pragma unroll
          for (int k = 0; k < 64; ++k) pk[k] = 0u;
          m_used = 0.f; lsum = 1.f;
          tc_fence_before();
          mbar_arrive(s_empty);
        } else {
          const float moff = m_used * p.scale_log2;
          const bool rvalid = (MODE == 0) ? true : (r < it.q_valid);
          const uint32_t rk0 = drop_row_key(dseed, (uint32_t)(it.lse_off + r));   // MODE 0: the thread's query row
          const uint32_t lse4 = sX + half * 512;   // shared-window address of this warp's 128 per-column statistics
          uint32_t va[32], vb[32];
          // operand type and "no padding among this warp's rows and columns" are compile-time variants of the
          // conversion (per-element bounds predicates and both 16-bit packings cost ~10 % of the stage otherwise)
          auto conv_t = [&](const uint32_t (&v)[32], int c, auto F16C, auto FULLC) {
            constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
            auto pack_pair = [&](float a, float b) -> uint32_t {   // shadows the run-time version
              if (F16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
              __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              return *reinterpret_cast<uint32_t*>(&h);
            };
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              float4 off = make_float4(moff, moff, moff, moff);
              if (MODE == 1) off = lds_f4(lse4 + (c * 32 + k) * 4);
              const int c0 = c * 32 + k;
              float p0 = fast_exp2(__uint_as_float(v[k]) * p.scale_log2 - off.x);
              float p1 = fast_exp2(__uint_as_float(v[k + 1]) * p.scale_log2 - off.y);
              float p2 = fast_exp2(__uint_as_float(v[k + 2]) * p.scale_log2 - off.z);
              float p3 = fast_exp2(__uint_as_float(v[k + 3]) * p.scale_log2 - off.w);
              if (!FULL) {
                if (!(rvalid && c0 < nvalid)) p0 = 0.f;
                if (!(rvalid && c0 + 1 < nvalid)) p1 = 0.f;
                if (!(rvalid && c0 + 2 < nvalid)) p2 = 0.f;
                if (!(rvalid && c0 + 3 < nvalid)) p3 = 0.f;
              }
              lsum += (p0 + p1) + (p2 + p3);   // the softmax denominator sees every key; dropout acts on the result
              float d0 = p0, d1 = p1, d2 = p2, d3 = p3;
              if (DROP) {
                if (MODE == 0) {   // row = this thread's query, columns = consecutive keys: one hash per two elements
                  const uint32_t kc = (uint32_t)(col_first + c0);
                  const uint32_t h0 = drop_pair(rk0, kc >> 1), h1 = drop_pair(rk0, (kc >> 1) + 1);
                  d0 = drop_keep_lo(h0, p.drop_thresh) ? p0 * p.drop_scale : 0.f; d1 = drop_keep_hi(h0, p.drop_thresh) ? p1 * p.drop_scale : 0.f;
                  d2 = drop_keep_lo(h1, p.drop_thresh) ? p2 * p.drop_scale : 0.f; d3 = drop_keep_hi(h1, p.drop_thresh) ? p3 * p.drop_scale : 0.f;
                } else {           // dV: row = this thread's KEY, columns = consecutive queries: one hash per element
                  const uint32_t kp = (uint32_t)(it.key0 + r) >> 1;
                  const bool hi = ((it.key0 + r) & 1) != 0;
                  const uint32_t q0 = (uint32_t)(it.lse_off + col_first + c0);
                  // lanes 2m / 2m+1 hold the two keys of one pair: they need the SAME four hashes (one takes the low,
                  // the other the high 16 bits), so each computes two and receives the partner's two by shuffle
                  const bool odd = (lane & 1) != 0;
                  const uint32_t qa = q0 + (odd ? 2u : 0u);
                  const uint32_t ha = drop_pair(drop_row_key(dseed, qa), kp), hb = drop_pair(drop_row_key(dseed, qa + 1), kp);
                  const uint32_t oa = __shfl_xor_sync(0xffffffffu, ha, 1), ob = __shfl_xor_sync(0xffffffffu, hb, 1);
                  const uint32_t h0 = odd ? oa : ha, h1 = odd ? ob : hb, h2 = odd ? ha : oa, h3 = odd ? hb : ob;
                  d0 = (hi ? drop_keep_hi(h0, p.drop_thresh) : drop_keep_lo(h0, p.drop_thresh)) ? p0 * p.drop_scale : 0.f;
                  d1 = (hi ? drop_keep_hi(h1, p.drop_thresh) : drop_keep_lo(h1, p.drop_thresh)) ? p1 * p.drop_scale : 0.f;
                  d2 = (hi ? drop_keep_hi(h2, p.drop_thresh) : drop_keep_lo(h2, p.drop_thresh)) ? p2 * p.drop_scale : 0.f;
                  d3 = (hi ? drop_keep_hi(h3, p.drop_thresh) : drop_keep_lo(h3, p.drop_thresh)) ? p3 * p.drop_scale : 0.f;
                }
              }
              pk[c * 16 + (k >> 1)] = pack_pair(d0, d1);
              pk[c * 16 + (k >> 1) + 1] = pack_pair(d2, d3);
            }
          };
          const int variant = (f16 ? 2 : 0) | ((nvalid >= 128 && __all_sync(0xffffffffu, rvalid)) ? 1 : 0);   // warp-uniform
          auto conv = [&](const uint32_t (&v)[32], int c) {
            switch (variant) {
              case 0: conv_t(v, c, std::false_type{}, std::false_type{}); break;
              case 1: conv_t(v, c, std::false_type{}, std::true_type{}); break;
              case 2: conv_t(v, c, std::true_type{}, std::false_type{}); break;
              default: conv_t(v, c, std::true_type{}, std::true_type{}); break;
            }
          };
          tmem_ld_32x32(s_addr, va);
          tmem_ld_wait();
          tmem_ld_32x32(s_addr + 32, vb);
          conv(va, 0);
          tmem_ld_wait();
          tmem_ld_32x32(s_addr + 64, va);
          conv(vb, 1);
          tmem_ld_wait();
          tmem_ld_32x32(s_addr + 96, vb);
          conv(va, 2);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(s_empty);   // the score tile is in registers: the next S MMAs may overwrite it
          conv(vb, 3);
        }
        // the previous item's epilogue runs here, under the S MMAs of this item's second step
        if (i == 0 && have_prev) epilogue(prev, prev_m, prev_l);
        // ---- P V of the previous step has completed: the P tile (and the accumulator) may be touched
        mbar_wait(p_empty, pe_ph ^ 1);
        pe_ph ^= 1;
        tc_fence_after();
        if (MODE == 0 && __any_sync(0xffffffffu, need)) {
          const uint32_t o_addr = tmem_base + lane_addr + 256 + half * 128;
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(o_addr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) * alpha);
            tmem_st_32x32(o_addr + c, v);
          }
          tmem_st_wait();
          l *= alpha;
        }
        l += lsum;
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {   // this warp's two 64-column slices of the P tile
          const uint32_t rowp = sP + (half * 2 + s2) * 16384 + r * 128;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int ch = t ^ (r & 7);
            sts_v4(rowp + ch * 16, pk[s2 * 32 + 4 * t], pk[s2 * 32 + 4 * t + 1], pk[s2 * 32 + 4 * t + 2], pk[s2 * 32 + 4 * t + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(p_full);
      }
      prev = it; prev_m = m_used; prev_l = l; have_prev = true;
    }
    if (have_prev) epilogue(prev, prev_m, prev_l);
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE, int CL, bool DROP>
static int launch_wide_d(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                       const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  auto kern = attn_wide_kernel<MODE, CL, DROP>;
  CSN_SET_MAX_SMEM(kern, WideAttnCfg::SMEM_BYTES);
  const int n_work = a.n_items / CL;
  const int workers = num_sms() / CL;
  const int grid = (n_work < workers ? n_work : workers) * CL;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = WideAttnCfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmQ, tmK, tmV, tmO, tmOlo, a));
  CSN_LAUNCH_OK("attn_wide_kernel");
  return 0;
}

template <int MODE, int CL>
static int launch_wide(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                       const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  if (a.drop_thresh) return launch_wide_d<MODE, CL, true>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
  return launch_wide_d<MODE, CL, false>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
}

int launch_attn_wide(int mode, bool pair, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                     const CUtensorMap& tmO, const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  if (mode == 0) return pair ? launch_wide<0, 2>(tmQ, tmK, tmV, tmO, tmOlo, a, stream) : launch_wide<0, 1>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
  return pair ? launch_wide<1, 2>(tmQ, tmK, tmV, tmO, tmOlo, a, stream) : launch_wide<1, 1>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
}

}  // namespace csn
