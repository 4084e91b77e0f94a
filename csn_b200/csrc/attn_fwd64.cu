// Attention forward for d_head = 64 (MinkowskiNet CSA head, h = 4):  O = softmax(Q K^T / sqrt(d)) V  per 128-query
// tile (reference: ScaledDotProductAttention.forward, MinkowskiNet/models/attention.py:69-75).
//
// At d_head 64 the MMAs are short (K = 64) and the kernel lives or dies by its softmax warps: with two warps per
// scheduler their dependent instruction chains leave 80 % of the issue slots empty (ncu, profiles/r2_experiments.md).
// This kernel therefore runs 16 softmax warps (four per TMEM lane quadrant, 32 key columns of every tile each) and
// removes every interaction between them from the tile loop:
//   * each (row, column quarter) has a PRIVATE running max, row sum and [128 x 64] accumulator O_g (4 x 64 TMEM
//     columns); the four partial results of a row are merged once per item in the epilogue
//     (m = max m_g, O = sum_g O_g 2^((m_g - m) scale), l likewise);
//   * a thread's 32 scores go into registers with one tcgen05.ld and the S buffer is handed back to the MMA warp
//     before any arithmetic; the maximum is exact (no optimistic pass / retry), the re-scale of the thread's own
//     accumulator rows is lazy (only when the max grows by > 2^8) and needs no vote beyond its own warp;
//   * the 16-bit probabilities go back into TMEM (two buffers) and are the A operand of P V from there (TS mode).
// TMEM: S @0 (128 columns, single buffer), O_g @128 + 64 g, P16 @384 / @448.
//
// CTA = 640 threads, one per SM, persistent over items; warp 0 = TMA producer (Q resident, ring of K_j / V_j slots),
// warp 1 = MMA issuer (S_{j+1} = Q K_{j+1}^T goes ahead of O += P_j V_j), warp 2 = TMEM allocator, warps 4-19 softmax.
#include <cstdlib>
#include <type_traits>

#include "attn_common.cuh"

namespace csn {

struct Fwd64Cfg {
  static constexpr int TILE_BYTES = 128 * 64 * 2;
  static constexpr int OUT_BYTES = 2 * 4 * 4096;      // hi / lo output slabs: [32 rows x 64 columns] per lane quadrant
  static constexpr int SLOT_BYTES = 128 * 64 * 2;
  static constexpr int NST = 9;
  static constexpr int SM_WARPS = 16;
  static constexpr int SM_THREADS = 32 * SM_WARPS;
  static constexpr int THREADS = 128 + SM_THREADS;
  static constexpr int XCH_BYTES = 4 * 128 * 2 * 4;   // [group][row][m | l]
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = TILE_BYTES + OUT_BYTES + NST * SLOT_BYTES + XCH_BYTES + BAR_BYTES + 1024;
  static constexpr int O_COL = 128, P_COL = 384;
};

template <bool DROP>
__global__ void __launch_bounds__(Fwd64Cfg::THREADS, 1)
attn_fwd64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                  const __grid_constant__ CUtensorMap tmOlo, const __grid_constant__ AttnFwdArgs p) {
  using Cfg = Fwd64Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sOut = sQ + Cfg::TILE_BYTES;
  const uint32_t sRing = sOut + Cfg::OUT_BYTES;
  float* xch = reinterpret_cast<float*>(smem + Cfg::TILE_BYTES + Cfg::OUT_BYTES + Cfg::NST * Cfg::SLOT_BYTES);
  uint8_t* bar_ptr = reinterpret_cast<uint8_t*>(xch) + Cfg::XCH_BYTES;
  const uint32_t bar_base = smem_u32(bar_ptr);
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t bq_full = bar_base + 8u * (2 * Cfg::NST + 0);
  const uint32_t bq_empty = bar_base + 8u * (2 * Cfg::NST + 1);
  const uint32_t s_full = bar_base + 8u * (2 * Cfg::NST + 2);
  const uint32_t s_empty = bar_base + 8u * (2 * Cfg::NST + 3);
  auto p_full = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 4 + b); };
  auto p_empty = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 6 + b); };
  const uint32_t acc_full = bar_base + 8u * (2 * Cfg::NST + 8);
  const uint32_t acc_empty = bar_base + 8u * (2 * Cfg::NST + 9);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 10));

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
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(bq_full, 1);
    mbar_init(bq_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, Cfg::SM_THREADS);
    for (int b = 0; b < 2; ++b) {
      mbar_init(p_full(b), Cfg::SM_THREADS);
      mbar_init(p_empty(b), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, Cfg::SM_THREADS);
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
      uint32_t ph = 0, q_ph = 0;
      auto load_slot = [&](const CUtensorMap* tm, int col0, int row0) {
        mbar_wait(kv_empty(st), ph ^ 1);
        mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
        tma_load_2d(sRing + st * Cfg::SLOT_BYTES, tm, kv_full(st), col0, row0);
        if (++st == Cfg::NST) { st = 0; ph ^= 1; }
      };
      for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
        const AttnItem it = p.items[wk];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bq_empty, q_ph ^ 1);
        mbar_arrive_expect_tx(bq_full, Cfg::TILE_BYTES);
        tma_load_2d(sQ, &tmQ, bq_full, it.col0, it.q_row0);
        q_ph ^= 1;
        // consumption order of the MMA warp: K0, [K1, V0], [K2, V1], ..., V_last
        load_slot(&tmK, it.col0, it.kv_row0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) load_slot(&tmK, it.col0, it.kv_row0 + (j + 1) * 128);
          load_slot(&tmV, it.col0, it.v_row0 + j * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, q_ph = 0, s_ph = 0, acc_ph = 0;
      uint32_t pf_ph[2] = {0, 0};
      for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
        const AttnItem it = p.items[wk];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bq_full, q_ph);
        q_ph ^= 1;
        tc_fence_after();
        auto issue_scores = [&](int j) {
          mbar_wait(s_empty, s_ph ^ 1);   // the previous score tile is in the softmax warps' registers
          tc_fence_after();
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          const uint32_t b_tile = sRing + st * Cfg::SLOT_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base, umma_desc_sw128(sQ + k * 32, 0, 1024), umma_desc_sw128(b_tile + k * 32, 0, 1024),
                        p.idesc_qk, k ? 1u : 0u);
          umma_commit(kv_empty(st));
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          umma_commit(s_full);
          if (j == nkv - 1) umma_commit(bq_empty);
          s_ph ^= 1;
        };
        issue_scores(0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) issue_scores(j + 1);
          const int pb = j & 1;
          mbar_wait(p_full(pb), pf_ph[pb]);   // P_j is in TMEM (and the accumulators were re-scaled if needed)
          pf_ph[pb] ^= 1;
          if (j == 0) mbar_wait(acc_empty, acc_ph ^ 1);   // previous item's accumulators have been read out
          tc_fence_after();
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          const uint32_t v_tile = sRing + st * Cfg::SLOT_BYTES;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // 16 keys = 8 TMEM columns of 16-bit pairs; keys 32 g .. 32 g + 31 -> O_g
            umma_f16_ts(tmem_base + Cfg::O_COL + (kk >> 1) * 64, tmem_base + Cfg::P_COL + pb * 64 + kk * 8,
                        umma_desc_sw128(v_tile + kk * 2048, 16384, 1024), p.idesc_pv, (j | (kk & 1)) ? 1u : 0u);
          umma_commit(kv_empty(st));
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          umma_commit(p_empty(pb));
        }
        umma_commit(acc_full);
        acc_ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================================================== softmax + epilogue
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;               // key columns [32 g, 32 g + 32) of every tile
    const int r = q * 32 + lane;                 // query row of the tile owned by this thread
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t s_ph = 0, accf_ph = 0;
    uint32_t pe_ph[2] = {0, 0};
    const float lo_scale = p.dtype == CSN_F16 ? 2048.f : 256.f;
    for (int wk = blockIdx.x; wk < p.n_items; wk += gridDim.x) {
      const AttnItem it = p.items[wk];
      const int nkv = (it.kv_len + 127) >> 7;
      float m_used = 0.f, l = 0.f;
      const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.lse_off + r));
      for (int j = 0; j < nkv; ++j) {
        const int pb = j & 1;
        const int nv = min(128, it.kv_len - j * 128) - g * 32;   // valid columns among this thread's 32 (<= 0: none)
        mbar_wait(s_full, s_ph);
        s_ph ^= 1;
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_addr + g * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty);
        float mx = -INFINITY;
        if (nv >= 32) {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[i + 2])); m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_used = (mx == -INFINITY) ? 0.f : mx;   // (no valid column in this quarter: P = exp2(-inf) = 0, l stays 0)
        } else if ((mx - m_used) * p.scale_log2 > 8.f) {
          alpha = fast_exp2((m_used - mx) * p.scale_log2);
          m_used = mx;
          need = true;
        }
        if (__any_sync(0xffffffffu, need)) {
          // the accumulator may still be receiving P V of the previous tile (P is double-buffered): wait for it
          if (j > 0) mbar_wait(p_empty(pb ^ 1), pe_ph[pb ^ 1] ^ 1);
          tc_fence_after();
          const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + g * 64;
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32(o_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(o_addr + c, o);
          }
          tmem_st_wait();
          l *= alpha;
        }
        const float moff = m_used * p.scale_log2;
        float l0 = 0.f, l1 = 0.f;
        uint32_t pk[16];
        auto exp_regs = [&](auto F16C, auto FULLC) {
          constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0 = fast_exp2(__uint_as_float(v[i]) * p.scale_log2 - moff);
            float a1 = fast_exp2(__uint_as_float(v[i + 1]) * p.scale_log2 - moff);
            if (!FULL) {
              if (!(i < nv)) a0 = 0.f;
              if (!(i + 1 < nv)) a1 = 0.f;
            }
            l0 += a0;   // the softmax denominator sees every key; dropout acts on the result
            l1 += a1;
            if (DROP) {
              const uint32_t hh = drop_pair(rk, (uint32_t)(j * 128 + g * 32 + i) >> 1);
              a0 = drop_keep_lo(hh, p.drop_thresh) ? a0 * p.drop_scale : 0.f;
              a1 = drop_keep_hi(hh, p.drop_thresh) ? a1 * p.drop_scale : 0.f;
            }
            if (F16) {
              __half2 h = __floats2half2_rn(a0, a1);
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            } else {
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
        };
        if (p.dtype == CSN_F16) {
          if (nv >= 32) exp_regs(std::true_type{}, std::true_type{}); else exp_regs(std::true_type{}, std::false_type{});
        } else {
          if (nv >= 32) exp_regs(std::false_type{}, std::true_type{}); else exp_regs(std::false_type{}, std::false_type{});
        }
        l += l0 + l1;
        mbar_wait(p_empty(pb), pe_ph[pb] ^ 1);   // the P V MMAs that read this buffer (tile j-2) have completed
        pe_ph[pb] ^= 1;
        tc_fence_after();
        tmem_st_32x16(tmem_base + lane_addr + Cfg::P_COL + pb * 64 + g * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(pb));
      }
      // ---- epilogue: merge the four column quarters of every row, O / l -> 16-bit (+ rounding residual), LSE
      mbar_wait(acc_full, accf_ph);
      accf_ph ^= 1;
      tc_fence_after();
      xch[(g * 128 + r) * 2] = l > 0.f ? m_used : -INFINITY;
      xch[(g * 128 + r) * 2 + 1] = l;
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::SM_THREADS) : "memory");
      float mg[4], lg[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 e = *reinterpret_cast<const float2*>(xch + (k * 128 + r) * 2);
        mg[k] = e.x;
        lg[k] = e.y;
      }
      const float m = fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3]));
      float ag[4], lt = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ag[k] = lg[k] > 0.f ? fast_exp2((mg[k] - m) * p.scale_log2) : 0.f;
        lt += lg[k] * ag[k];
      }
      const float inv_l = 1.f / lt;
      const bool valid = r < it.q_valid;
      // this thread's 16 output columns [16 g, 16 g + 16) from all four accumulators
      float o[16];
      {
        uint32_t t0[16], t1[16], t2[16], t3[16];
        const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + g * 16;
        tmem_ld_32x16(o_addr, t0);
        tmem_ld_32x16(o_addr + 64, t1);
        tmem_ld_32x16(o_addr + 128, t2);
        tmem_ld_32x16(o_addr + 192, t3);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(acc_empty);   // the accumulators are in registers
#pragma unroll
        for (int i = 0; i < 16; ++i)
          o[i] = valid ? (__uint_as_float(t0[i]) * ag[0] + __uint_as_float(t1[i]) * ag[1] + __uint_as_float(t2[i]) * ag[2] +
                          __uint_as_float(t3[i]) * ag[3]) * inv_l
                       : 0.f;
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        if (p.dtype == CSN_F16) {
          __half2 h = __floats2half2_rn(o[i], o[i + 1]);
          hi[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          const float2 b = __half22float2(h);
          __half2 e = __floats2half2_rn((o[i] - b.x) * lo_scale, (o[i + 1] - b.y) * lo_scale);
          lo[i >> 1] = *reinterpret_cast<uint32_t*>(&e);
        } else {
          __nv_bfloat162 h = __floats2bfloat162_rn(o[i], o[i + 1]);
          hi[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          const float2 b = __bfloat1622float2(h);
          __nv_bfloat162 e = __floats2bfloat162_rn((o[i] - b.x) * lo_scale, (o[i + 1] - b.y) * lo_scale);
          lo[i >> 1] = *reinterpret_cast<uint32_t*>(&e);
        }
      }
      // the four warps of a lane quadrant fill one [32 rows x 64 columns] slab each for O and its residual
      // (128B-swizzled rows, 32 bytes per thread); warp g = 0 hands them to the TMA store engine
      const bool want_lo = p.Olo != nullptr;
      if (g == 0 && lane == 0) tma_store_wait_read<0>();   // previous item's slabs have left
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
      {
        const uint32_t row_hi = sOut + q * 4096 + lane * 128, row_lo = row_hi + 16384;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t ch = ((uint32_t)(2 * g + t) ^ ((uint32_t)lane & 7u)) << 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + ch), "r"(hi[4 * t]), "r"(hi[4 * t + 1]), "r"(hi[4 * t + 2]), "r"(hi[4 * t + 3]) : "memory");
          if (want_lo)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_lo + ch), "r"(lo[4 * t]), "r"(lo[4 * t + 1]), "r"(lo[4 * t + 2]), "r"(lo[4 * t + 3]) : "memory");
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
      if (g == 0) {
        if (elect_one()) {
          tma_store_2d(&tmO, sOut + q * 4096, it.col0, it.o_row0 + q * 32);
          if (want_lo) tma_store_2d(&tmOlo, sOut + 16384 + q * 4096, it.col0, it.o_row0 + q * 32);
          tma_store_commit();
        }
        __syncwarp();
        if (p.lse) p.lse[it.lse_off + r] = valid ? (m * p.scale + __logf(lt)) : 0.f;
      }
      // (xch is rewritten only after the next item's acc_full wait + this barrier's successor: every reader of this
      //  item's entries has passed the reads above before it can arrive there)
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
static int launch_fwd64_d(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                          const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  auto kern = attn_fwd64_kernel<DROP>;
  CSN_SET_MAX_SMEM(kern, Fwd64Cfg::SMEM_BYTES);
  const int grid = a.n_items < num_sms() ? a.n_items : num_sms();
  kern<<<grid, Fwd64Cfg::THREADS, Fwd64Cfg::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, tmOlo, a);
  CSN_LAUNCH_OK("attn_fwd64_kernel");
  return 0;
}

int launch_attn_fwd64(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                      const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream) {
  return a.drop_thresh ? launch_fwd64_d<true>(tmQ, tmK, tmV, tmO, tmOlo, a, stream)
                       : launch_fwd64_d<false>(tmQ, tmK, tmV, tmO, tmOlo, a, stream);
}

}  // namespace csn
