// Fused attention backward, query side (autograd of csa_models.py:139-142 / attention.py:69-75):
//   for one 128-row query tile i and every 128-key tile j of its key range
//     S  = Q_i K_j^T,  dP = dO_i V_j^T                         (tcgen05, fp32 in TMEM)
//     P  = exp(S*scale - lse_i),  dS = P o (dP - delta_i) * scale   (registers, lane = query row)
//     dQ_i += dS K_j                                           (dS staged in SMEM as the A operand;
//                                                               K_j re-read MN-major from the same rows)
//   dS is also written to HBM (16-bit) so that dK = dS^T Q runs as one batched csn_gemm.
//   delta_i = rowsum(dO_i o O_i) comes from csn_attn_delta.
// CTA layout as in attn_fwd.cu: warp 0 TMA, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// element-wise stage + epilogue. TMEM: S @0 (128 cols) | dP @128 (128) | dQ @256 (<= 256).
#include <cstdlib>
#include "host_util.h"
#include "ptx.cuh"
#include <type_traits>

#include "attn_common.cuh"

namespace csn {

__device__ __forceinline__ float fast_exp2_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnBwdItem {
  int q_row0;    // first query row in the Q view == first row of dO / dQ (block rows) is given separately
  int q_valid;   // real query rows in the tile
  int kv_row0;   // first key row
  int kv_len;    // keys attended to
  int o_row0;    // first row of this tile in dO (input) and dQ (output)
  int col0;      // head * d
  int stat_off;  // lse[stat_off + r], delta[stat_off + r]
  int ds_row0;   // first row of this tile in the dS buffer
  int ds_col0;   // first column of key tile 0 in the dS buffer
  int flags;     // bit 0: zero-fill rows >= q_valid of dQ
  int pad0, pad1;
};

struct AttnBwdArgs {
  const AttnBwdItem* items;
  int n_items;
  void* dQ; long long lddq;
  void* dS; long long ldds;
  const float* lse; const float* delta;
  float scale_log2, scale;
  int dtype;
  uint32_t idesc_s;   // M=128, N=128, K-major x K-major
  uint32_t idesc_dq;  // M=128, N=128 (DH>=128) or DH, A K-major, B MN-major
  int prefetch;       // wide dS kernel: L2-prefetch the next item's tiles
  // dropout on the probabilities (attn_common.cuh): dP/dP_dropped = mask/(1-p); row id = stat_off + row, column = key
  uint32_t drop_seed, drop_thresh;
  float drop_scale;
  const uint32_t* drop_epoch;
};

#ifndef CSN_DQ_EW_WARPS
#define CSN_DQ_EW_WARPS 16
#endif

template <int DH>
struct BwdCfg {
  static constexpr int KB = DH / 64;
  static constexpr int TILE_BYTES = 128 * DH * 2;            // Q_i, dO_i resident
  static constexpr int DS_BYTES = 128 * 128 * 2;
  static constexpr int SLOT_BYTES = (DH >= 128) ? 32768 : 128 * DH * 2;   // [128 keys x (SLOT_BYTES/256) cols]
  static constexpr int SLOTS_PER_TILE = (128 * DH * 2) / SLOT_BYTES;
  static constexpr int KB_PER_SLOT = KB / SLOTS_PER_TILE;
  static constexpr int COLS_PER_SLOT = 64 * KB_PER_SLOT;     // dQ columns produced from one K slot
  static constexpr int NST = (DH >= 128) ? 2 : 6;
  static constexpr int SMEM_BYTES = 2 * TILE_BYTES + DS_BYTES + NST * SLOT_BYTES + 256 + 1024;
  static constexpr int EW_WARPS = (DH == 64) ? CSN_DQ_EW_WARPS : 4;   // element-wise warps (several per TMEM lane quadrant at d_head 64)
  static constexpr int EW_THREADS = 32 * EW_WARPS;
  static constexpr int THREADS = 128 + EW_THREADS;
};

template <int DH, int CL, bool WITH_DQ, bool DROP>
__global__ void __launch_bounds__(BwdCfg<DH>::THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmDQ,
                   const __grid_constant__ AttnBwdArgs p) {
  using Cfg = BwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sDO = sQ + Cfg::TILE_BYTES;
  const uint32_t sDS = sDO + Cfg::TILE_BYTES;
  const uint32_t sKV = sDS + Cfg::DS_BYTES;
  const uint32_t bar_base = sKV + Cfg::NST * Cfg::SLOT_BYTES;
  uint8_t* bar_ptr = smem + 2 * Cfg::TILE_BYTES + Cfg::DS_BYTES + Cfg::NST * Cfg::SLOT_BYTES;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t bq_full = bar_base + 8u * (2 * Cfg::NST + 0);
  const uint32_t bq_empty = bar_base + 8u * (2 * Cfg::NST + 1);
  // S and dP of tile j are in TMEM (buffer b) / have been read. Without the dQ accumulator the 256 freed
  // TMEM columns double-buffer S|dP, so the element-wise stage of tile j overlaps the MMAs of tile j+1.
  auto sdp_full = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 2 + b); };
  auto sdp_empty = [&](int b) { return bar_base + 8u * (2 * Cfg::NST + 4 + b); };
  const uint32_t ds_full = bar_base + 8u * (2 * Cfg::NST + 6);    // dS_j is in SMEM
  const uint32_t ds_empty = bar_base + 8u * (2 * Cfg::NST + 7);   // dQ MMAs of tile j done reading it
  const uint32_t dq_full = bar_base + 8u * (2 * Cfg::NST + 8);
  const uint32_t dq_empty = bar_base + 8u * (2 * Cfg::NST + 9);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 10));

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
      mbar_init(kv_empty(s), CL);
    }
    mbar_init(bq_full, 1);
    mbar_init(bq_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(sdp_full(b), 1);
      mbar_init(sdp_empty(b), Cfg::EW_THREADS);
    }
    mbar_init(ds_full, Cfg::EW_THREADS);
    mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 128);
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

  // CL == 2: items 2m / 2m+1 (two query tiles of one (block, chunk, head)) stream the same K / V tiles;
  // each CTA of the cluster loads half of every streamed slot and multicasts it to both.
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int worker = (int)blockIdx.x / CL, n_workers = (int)gridDim.x / CL;
  const int n_work = p.n_items / CL;
  constexpr uint16_t MC_MASK = (1u << CL) - 1;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, q_ph = 0;
      auto load_tile = [&](const CUtensorMap* tm, int col0, int row0) {   // 128 rows x DH cols, K-major k-blocks
#pragma unroll 1
        for (int s = 0; s < Cfg::SLOTS_PER_TILE; ++s) {
          mbar_wait(kv_empty(st), ph ^ 1);
          mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
#pragma unroll
          for (int kb = 0; kb < Cfg::KB_PER_SLOT; ++kb) {
            const uint32_t dst = sKV + st * Cfg::SLOT_BYTES + kb * 16384;
            const int c0 = col0 + (s * Cfg::KB_PER_SLOT + kb) * 64;
            if (CL == 1) tma_load_2d(dst, tm, kv_full(st), c0, row0);
            else if ((kb % CL) == rank) tma_load_2d_mc(dst, tm, kv_full(st), c0, row0, MC_MASK);
          }
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnBwdItem it = p.items[wk * CL + rank];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bq_empty, q_ph ^ 1);
        mbar_arrive_expect_tx(bq_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          tma_load_2d(sQ + kb * 16384, &tmQ, bq_full, it.col0 + kb * 64, it.q_row0);
          tma_load_2d(sDO + kb * 16384, &tmDO, bq_full, it.col0 + kb * 64, it.o_row0);
        }
        q_ph ^= 1;
        // consumption order of the MMA warp: [K0, V0], [K1, V1], K0 (dQ), [K2, V2], K1 (dQ), ...
        load_tile(&tmK, it.col0, it.kv_row0);   // for S
        load_tile(&tmV, it.col0, it.kv_row0);   // for dP
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) {
            load_tile(&tmK, it.col0, it.kv_row0 + (j + 1) * 128);
            load_tile(&tmV, it.col0, it.kv_row0 + (j + 1) * 128);
          }
          if (WITH_DQ) load_tile(&tmK, it.col0, it.kv_row0 + j * 128);   // for dQ (same bytes, consumed MN-major)
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, q_ph = 0, ds_ph = 0, dq_ph = 0;
      uint32_t sdp_ph[2] = {0, 0};
      auto mma_kmajor = [&](uint32_t a_base, uint32_t d_tmem) {   // D = A_tile (resident) x slots^T
#pragma unroll 1
        for (int s = 0; s < Cfg::SLOTS_PER_TILE; ++s) {
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < Cfg::KB_PER_SLOT; ++kb) {
            const uint32_t a_tile = a_base + (s * Cfg::KB_PER_SLOT + kb) * 16384;
            const uint32_t b_tile = sKV + st * Cfg::SLOT_BYTES + kb * 16384;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d_tmem, umma_desc_sw128(a_tile + k * 32, 0, 1024), umma_desc_sw128(b_tile + k * 32, 0, 1024),
                          p.idesc_s, (s | kb | k) ? 1u : 0u);
          }
          if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnBwdItem it = p.items[wk * CL + rank];
        const int nkv = (it.kv_len + 127) >> 7;
        mbar_wait(bq_full, q_ph);
        q_ph ^= 1;
        tc_fence_after();
        auto issue_sdp = [&](int j) {
          const int b = WITH_DQ ? 0 : (j & 1);
          mbar_wait(sdp_empty(b), sdp_ph[b] ^ 1);   // S / dP previously held by this buffer have been read
          tc_fence_after();
          mma_kmajor(sQ, tmem_base + b * 256);         // S  = Q  K_j^T
          mma_kmajor(sDO, tmem_base + b * 256 + 128);  // dP = dO V_j^T
          umma_commit(sdp_full(b));
          if (j == nkv - 1) umma_commit(bq_empty);   // Q_i / dO_i are not read by the dQ MMAs
          sdp_ph[b] ^= 1;
        };
        issue_sdp(0);
        for (int j = 0; j < nkv; ++j) {
          // S / dP of the next tile go in FRONT of this tile's dQ MMAs: the element-wise warps (the longer stage)
          // get their next input as soon as they have read the current one, dQ_j runs in their shadow
          if (j + 1 < nkv) issue_sdp(j + 1);
          if (!WITH_DQ) continue;
          mbar_wait(ds_full, ds_ph);          // dS_j staged
          ds_ph ^= 1;
          if (j == 0) mbar_wait(dq_empty, dq_ph ^ 1);   // previous item's dQ has been read out
          tc_fence_after();
#pragma unroll 1
          for (int s = 0; s < Cfg::SLOTS_PER_TILE; ++s) {
            mbar_wait(kv_full(st), ph);
            tc_fence_after();
            const uint32_t k_tile = sKV + st * Cfg::SLOT_BYTES;
            const uint32_t d_tmem = tmem_base + 256 + s * Cfg::COLS_PER_SLOT;
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 128 keys in steps of 16
              const uint32_t a_addr = sDS + (k >> 2) * 16384 + (k & 3) * 32;
              umma_f16_ss(d_tmem, umma_desc_sw128(a_addr, 0, 1024), umma_desc_sw128(k_tile + k * 2048, 16384, 1024),
                          p.idesc_dq, (j | k) ? 1u : 0u);
            }
            if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
            if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          }
          umma_commit(ds_empty);
        }
        if (WITH_DQ) {
          umma_commit(dq_full);
          dq_ph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== element-wise stage + epilogue
    // warps 4-7 own the TMEM lane quadrants; with EW_WARPS == 8 (d_head 64, where this stage and not the MMAs is
    // the bottleneck) warp q + 4 shares quadrant q and takes the upper 64 key columns of every tile
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int COLS = 128 / (Cfg::EW_WARPS / 4);
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t sdp_ph[2] = {0, 0};
    uint32_t dse_ph = 0, dqf_ph = 0;
    uint8_t* sDS_ptr = smem + 2 * Cfg::TILE_BYTES;
    constexpr float LOG2E = 1.4426950408889634f;
    auto pack_pair = [&](float a, float b) -> uint32_t {
      if (p.dtype == CSN_F16) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    };
    for (int wk = worker; wk < n_work; wk += n_workers) {
      const AttnBwdItem it = p.items[wk * CL + rank];
      const int nkv = (it.kv_len + 127) >> 7;
      const bool valid = r < it.q_valid;
      const float lse_l2 = valid ? p.lse[it.stat_off + r] * LOG2E : 0.f;
      const float dlt_s = valid ? p.delta[it.stat_off + r] * p.scale : 0.f;
      const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.stat_off + r));
      for (int j = 0; j < nkv; ++j) {
        const int b = WITH_DQ ? 0 : (j & 1);
        mbar_wait(sdp_full(b), sdp_ph[b]);
        sdp_ph[b] ^= 1;
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_addr + b * 256;
        const int nvalid = min(128, it.kv_len - j * 128);
        // The stage is instruction-bound at d_head 64 (ncu: ~26 SASS instructions per element pair with per-element
        // bounds predicates and both 16-bit packings): the operand type and "every row and key of this tile exists"
        // are compile-time variants of the loop, selected per tile by a warp-uniform branch.
        auto ew_tile = [&](auto F16C, auto FULLC) {
          constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
          bool waited = false;
#pragma unroll 1
          for (int c = half * COLS; c < half * COLS + COLS; c += 32) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(s_addr + c, sv);
            tmem_ld_32x32(s_addr + 128 + c, dv);
            tmem_ld_wait();
            if (c + 32 >= half * COLS + COLS) {   // this thread's scores are in registers: the MMA warp may overwrite them
              tc_fence_before();
              mbar_arrive(sdp_empty(b));
            }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float m0 = p.scale, m1 = p.scale;   // (d P_dropped / d P = mask / (1 - p)) * scale
              if (DROP) {
                const uint32_t hh = drop_pair(rk, (uint32_t)(j * 128 + c + i) >> 1);
                m0 = drop_keep_lo(hh, p.drop_thresh) ? p.drop_scale * p.scale : 0.f;
                m1 = drop_keep_hi(hh, p.drop_thresh) ? p.drop_scale * p.scale : 0.f;
              }
              float d0 = fast_exp2_b(__uint_as_float(sv[i]) * p.scale_log2 - lse_l2) * fmaf(__uint_as_float(dv[i]), m0, -dlt_s);
              float d1 = fast_exp2_b(__uint_as_float(sv[i + 1]) * p.scale_log2 - lse_l2) * fmaf(__uint_as_float(dv[i + 1]), m1, -dlt_s);
              if (!FULL) {
                if (!(valid && c + i < nvalid)) d0 = 0.f;
                if (!(valid && c + i + 1 < nvalid)) d1 = 0.f;
              }
              if (F16) {
                __half2 h = __floats2half2_rn(d0, d1);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
              } else {
                __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
              }
            }
            if (!waited) {
              if (WITH_DQ) {
                mbar_wait(ds_empty, dse_ph ^ 1);   // dQ MMAs of the previous tile no longer read the staging tile
                dse_ph ^= 1;
              }
              if (warp == 4 && lane == 0) tma_store_wait_read<0>();   // ... nor does its TMA store to HBM
              asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EW_THREADS) : "memory");
              waited = true;
            }
            const uint32_t rowp = sDS + (c >> 6) * 16384 + r * 128;
            const int chunk0 = (c & 63) >> 3;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int ch = (chunk0 + t) ^ (r & 7);
              sts_v4(rowp + ch * 16, pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
            }
          }
        };
        const bool full = (nvalid == 128) && __all_sync(0xffffffffu, valid);
        if (p.dtype == CSN_F16) {
          if (full) ew_tile(std::true_type{}, std::true_type{}); else ew_tile(std::true_type{}, std::false_type{});
        } else {
          if (full) ew_tile(std::false_type{}, std::true_type{}); else ew_tile(std::false_type{}, std::false_type{});
        }
        fence_proxy_async_smem();
        if (WITH_DQ) mbar_arrive(ds_full);
        // HBM copy of dS_j for dK = dS^T Q: the staged tile is exactly two TMA boxes [128 rows x 64 keys]
        if (!WITH_DQ || p.dS != nullptr) asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EW_THREADS) : "memory");
        if (warp == 4 && lane == 0 && (!WITH_DQ || p.dS != nullptr)) {   // dS == NULL: dK comes from csn_attn_bwd_dkv
          tma_store_2d(&tmDS, sDS, it.ds_col0 + j * 128, it.ds_row0);
          tma_store_2d(&tmDS, sDS + 16384, it.ds_col0 + j * 128 + 64, it.ds_row0);
          tma_store_commit();
        }
      }
      if (!WITH_DQ) continue;
      // ---- epilogue: dQ tile -> 16-bit
      mbar_wait(dq_full, dqf_ph);
      dqf_ph ^= 1;
      tc_fence_after();
      // dQ tile -> 16-bit through this warp's own rows of the dS staging tile and TMA stores (rows >= q_valid: zeros).
      // The last dS_j of this item has been consumed (dq_full) and its HBM copy is drained first.
      if (warp == 4 && lane == 0) tma_store_wait_read<0>();
      asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EW_THREADS) : "memory");
      // the upper-half warps go on to the next item: its first write to the staging tile sits behind a barrier that
      // warps 4-7 only reach once their dQ slabs below have left it
      if (half != 0) continue;
      const uint32_t o_addr = tmem_base + lane_addr + 256;
      int slab = 0;
#pragma unroll 1
      for (int c = 0; c < DH; c += 64) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(o_addr + c, v0);
        tmem_ld_32x32(o_addr + c + 32, v1);
        tmem_ld_wait();
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          w[i >> 1] = pack_pair(valid ? __uint_as_float(v0[i]) : 0.f, valid ? __uint_as_float(v0[i + 1]) : 0.f);
          w[16 + (i >> 1)] = pack_pair(valid ? __uint_as_float(v1[i]) : 0.f, valid ? __uint_as_float(v1[i + 1]) : 0.f);
        }
        const uint32_t buf = sDS + (slab & 1) * 16384 + q * 4096;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const uint32_t rowaddr = buf + lane * 128;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t a = rowaddr + (((uint32_t)t ^ ((uint32_t)lane & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * t]), "r"(w[4 * t + 1]), "r"(w[4 * t + 2]), "r"(w[4 * t + 3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&tmDQ, buf, it.col0 + c, it.o_row0 + q * 32);
          tma_store_commit();
        }
        ++slab;
      }
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      tc_fence_before();
      mbar_arrive(dq_empty);
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dS-only variant for d_head = 256 with 256-key steps.  An SS-mode M=128 tcgen05.mma costs about the same for
// N = 128 and N = 256 (profiles/r1_experiments.md), so the 128-key tiles of the kernel above leave half of the
// tensor pipe idle.  Here S = Q_i K^T and dP = dO_i V^T are [128 x 256] accumulators (TMEM columns 0-255 / 256-511,
// the whole TMEM), each single-buffered but alternating: S_{j+1} is issued as soon as the element-wise warps have
// turned S_j into P_j (16-bit, kept in registers), dP_{j+1} as soon as dP_j has been read, so the MMA pipe runs
// while the other accumulator is consumed.  8 element-wise warps: warp pair (q, q+4) shares TMEM lane quadrant q
// and splits the 256 key columns.
// Nothing is resident in SMEM: every ring slot carries one 64-column slice of BOTH operands of four MMAs
// ([128 x 64] of Q_i or dO_i next to [256 x 64] of K or V, 48 KB), so that the ring is 3 deep (a 2-deep ring of
// 32 KB K/V slices next to resident Q_i/dO_i tiles stalls the MMA pipe on TMA latency: 336 vs 256 us measured
// without stores) and 32 KB remain for the output slabs.  Q_i/dO_i slices are re-read per key step (L2 hits).
// L2-prefetching the next item's tiles was tried and is harmful (the prefetches queue in front of the ring's
// loads in the TMA unit: 328 -> 437 us).
template <int DBG>
struct WideCfgT {
  static constexpr int A_BYTES = 128 * 64 * 2;       // slice of Q_i / dO_i
  static constexpr int B_BYTES = 256 * 64 * 2;       // slice of K / V
  static constexpr int SLOT_BYTES = A_BYTES + B_BYTES;
  static constexpr int NST = DBG == 2 ? 3 : 4;     // measured: 4-deep ring + single slabs 328 us, 3-deep + double slabs 345 us
  static constexpr int NSLAB = DBG == 2 ? 2 : 1;
  static constexpr int STG_BYTES = 8 * NSLAB * 4096;     // [32 rows x 64 keys] slabs per element-wise warp
  static constexpr int SMEM_BYTES = NST * SLOT_BYTES + STG_BYTES + 256 + 1024;
};
using WideCfg = WideCfgT<0>;

template <int CL, int DBG, bool DROP>
__global__ void __launch_bounds__(384, 1)
attn_bwd_ds_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ AttnBwdArgs p) {
  using Cfg = WideCfgT<DBG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sKV = smem_u32(smem);
  const uint32_t sSTG = sKV + Cfg::NST * Cfg::SLOT_BYTES;
  const uint32_t bar_base = sSTG + Cfg::STG_BYTES;
  uint8_t* bar_ptr = smem + Cfg::NST * Cfg::SLOT_BYTES + Cfg::STG_BYTES;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (Cfg::NST + s); };
  const uint32_t s_full = bar_base + 8u * (2 * Cfg::NST + 0), s_empty = bar_base + 8u * (2 * Cfg::NST + 1);
  const uint32_t dp_full = bar_base + 8u * (2 * Cfg::NST + 2), dp_empty = bar_base + 8u * (2 * Cfg::NST + 3);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * Cfg::NST + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dseed = DROP ? drop_seed_eff(p.drop_seed, p.drop_epoch) : 0u;   // (one load per thread, train-mode variants only)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::NST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), CL); }
    mbar_init(s_full, 1); mbar_init(s_empty, 256);
    mbar_init(dp_full, 1); mbar_init(dp_empty, 256);
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
      uint32_t ph = 0;
      // one accumulator's operands: 4 slots of { A[128 x 64] from (tmA, a_row0), B[256 x 64] from (tmB, b_row0) }
      auto load_step = [&](const CUtensorMap* tmA, int a_row0, const CUtensorMap* tmB, int b_row0, int col0) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(kv_empty(st), ph ^ 1);
          mbar_arrive_expect_tx(kv_full(st), Cfg::SLOT_BYTES);
          const uint32_t dst = sKV + st * Cfg::SLOT_BYTES;
          tma_load_2d(dst, tmA, kv_full(st), col0 + kb * 64, a_row0);
          if (CL == 1) {
            tma_load_2d(dst + Cfg::A_BYTES, tmB, kv_full(st), col0 + kb * 64, b_row0);
            tma_load_2d(dst + Cfg::A_BYTES + 16384, tmB, kv_full(st), col0 + kb * 64, b_row0 + 128);
          } else {   // each CTA of the pair loads one 128-key half and multicasts it to both
            tma_load_2d_mc(dst + Cfg::A_BYTES + rank * 16384, tmB, kv_full(st), col0 + kb * 64, b_row0 + rank * 128, MC_MASK);
          }
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnBwdItem it = p.items[wk * CL + rank];
        const int nst = (it.kv_len + 255) >> 8;
        if (p.prefetch && wk + n_workers < n_work) {
          // first touches of the next item's Q_i / dO_i (and, for the first pair of a chunk, K / V) would reach
          // the ring with HBM latency, which a 4-deep ring does not cover: pull them into L2 one item ahead
          const AttnBwdItem nx = p.items[(wk + n_workers) * CL + rank];
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            tma_prefetch_l2_2d(&tmQ, nx.col0 + kb * 64, nx.q_row0);
            tma_prefetch_l2_2d(&tmDO, nx.col0 + kb * 64, nx.o_row0);
          }
          if (nx.kv_row0 != it.kv_row0 || nx.col0 != it.col0) {
            const int nrows = (nx.kv_len + 127) >> 7;
#pragma unroll 1
            for (int t = rank; t < nrows; t += CL)
#pragma unroll 1
              for (int kb = 0; kb < 4; ++kb) {
                tma_prefetch_l2_2d(&tmK, nx.col0 + kb * 64, nx.kv_row0 + t * 128);
                tma_prefetch_l2_2d(&tmV, nx.col0 + kb * 64, nx.kv_row0 + t * 128);
              }
          }
        }
        for (int j = 0; j < nst; ++j) {
          load_step(&tmQ, it.q_row0, &tmK, it.kv_row0 + j * 256, it.col0);    // for S
          load_step(&tmDO, it.o_row0, &tmV, it.kv_row0 + j * 256, it.col0);   // for dP
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, s_ph = 0, dp_ph = 0;
      auto mma_step = [&](uint32_t d_tmem) {   // D[128 x 256] = sum over 4 slots of A_slice x B_slice^T
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(kv_full(st), ph);
          tc_fence_after();
          const uint32_t a_tile = sKV + st * Cfg::SLOT_BYTES;
          const uint32_t b_tile = a_tile + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(d_tmem, umma_desc_sw128(a_tile + k * 32, 0, 1024), umma_desc_sw128(b_tile + k * 32, 0, 1024),
                        p.idesc_s, (kb | k) ? 1u : 0u);
          if (CL == 1) umma_commit(kv_empty(st)); else umma_commit_mc(kv_empty(st), MC_MASK);
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      };
      for (int wk = worker; wk < n_work; wk += n_workers) {
        const AttnBwdItem it = p.items[wk * CL + rank];
        const int nst = (it.kv_len + 255) >> 8;
        for (int j = 0; j < nst; ++j) {
          mbar_wait(s_empty, s_ph ^ 1);      // S_{j-1} has been turned into P_{j-1}
          tc_fence_after();
          mma_step(tmem_base);               // S = Q K^T
          umma_commit(s_full);
          s_ph ^= 1;
          mbar_wait(dp_empty, dp_ph ^ 1);    // dP_{j-1} has been read
          tc_fence_after();
          mma_step(tmem_base + 256);         // dP = dO V^T
          umma_commit(dp_full);
          dp_ph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== element-wise stage
    const int ew = warp - 4;
    const int q = warp & 3, half = ew >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t slab0 = sSTG + ew * Cfg::NSLAB * 4096;
    int flip = 0;
    uint32_t s_ph = 0, dp_ph = 0;
    constexpr float LOG2E = 1.4426950408889634f;
    const bool f16 = p.dtype == CSN_F16;
    auto pack_pair = [&](float a, float b) -> uint32_t {
      if (f16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    };
    auto unpack_pair = [&](uint32_t w) -> float2 {
      if (f16) return __half22float2(*reinterpret_cast<__half2*>(&w));
      return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
    };
    for (int wk = worker; wk < n_work; wk += n_workers) {
      const AttnBwdItem it = p.items[wk * CL + rank];
      const int nst = (it.kv_len + 255) >> 8;
      const bool valid = r < it.q_valid;
      const float lse_l2 = valid ? p.lse[it.stat_off + r] * LOG2E : 0.f;
      const float dlt = valid ? p.delta[it.stat_off + r] : 0.f;
      const uint32_t rk = drop_row_key(dseed, (uint32_t)(it.stat_off + r));
      for (int j = 0; j < nst; ++j) {
        const int key0 = j * 256 + half * 128;          // first key of this warp's 128 columns
        const int nvalid = valid ? it.kv_len - key0 : 0;   // columns [0, nvalid) of the 128 are real keys
        // ---- pass A: S -> P (16-bit pairs in registers)
        uint32_t pk[64];
        mbar_wait(s_full, s_ph);
        s_ph ^= 1;
        tc_fence_after();
        // operand type and "every row and key of this warp's half exists" are compile-time variants of both passes
        // (per-element bounds predicates and both 16-bit packings / unpackings otherwise: ~10 % of the stage)
        const int variant = (f16 ? 2 : 0) | (__all_sync(0xffffffffu, nvalid >= 128) ? 1 : 0);   // warp-uniform
        auto pass_a = [&](auto F16C, auto FULLC) {
          constexpr bool F16 = decltype(F16C)::value, FULL = decltype(FULLC)::value;
          auto pack_t = [&](float a, float b) -> uint32_t {
            if (F16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            return *reinterpret_cast<uint32_t*>(&h);
          };
          const uint32_t s_addr = tmem_base + lane_addr + half * 128;
          uint32_t va[32], vb[32];
          tmem_ld_32x32(s_addr, va);
#pragma unroll
          for (int c = 0; c < 4; c += 2) {
            tmem_ld_wait();
            tmem_ld_32x32(s_addr + (c + 1) * 32, vb);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = fast_exp2_b(__uint_as_float(va[i]) * p.scale_log2 - lse_l2);
              float p1 = fast_exp2_b(__uint_as_float(va[i + 1]) * p.scale_log2 - lse_l2);
              if (!FULL) {
                if (!(c * 32 + i < nvalid)) p0 = 0.f;
                if (!(c * 32 + i + 1 < nvalid)) p1 = 0.f;
              }
              pk[c * 16 + (i >> 1)] = pack_t(p0, p1);
            }
            tmem_ld_wait();
            if (c + 2 < 4) tmem_ld_32x32(s_addr + (c + 2) * 32, va);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = fast_exp2_b(__uint_as_float(vb[i]) * p.scale_log2 - lse_l2);
              float p1 = fast_exp2_b(__uint_as_float(vb[i + 1]) * p.scale_log2 - lse_l2);
              if (!FULL) {
                if (!((c + 1) * 32 + i < nvalid)) p0 = 0.f;
                if (!((c + 1) * 32 + i + 1 < nvalid)) p1 = 0.f;
              }
              pk[(c + 1) * 16 + (i >> 1)] = pack_t(p0, p1);
            }
          }
        };
        switch (variant) {
          case 0: pass_a(std::false_type{}, std::false_type{}); break;
          case 1: pass_a(std::false_type{}, std::true_type{}); break;
          case 2: pass_a(std::true_type{}, std::false_type{}); break;
          default: pass_a(std::true_type{}, std::true_type{}); break;
        }
        tc_fence_before();
        mbar_arrive(s_empty);
        // ---- pass B: dS = P o (dP - delta) * scale -> slab -> HBM.  (dP - delta)*scale is one FFMA per element, the
        //      product with P is formed in fp32 (packed 16-bit arithmetic here was no faster and cost 1e-4 on dWq).
        //      Masked columns have P = 0; the select below only guards against non-finite garbage in rows past
        //      kv_len and is skipped when the whole half is valid.
        mbar_wait(dp_full, dp_ph);
        dp_ph ^= 1;
        tc_fence_after();
        {
          const uint32_t d_addr = tmem_base + lane_addr + 256 + half * 128;
          const float nds = -dlt * p.scale;
          uint32_t da[32], db[32];
          auto emit_t = [&](const uint32_t (&dv)[32], int c, auto F16C, auto FULLC) {
            constexpr bool F16 = decltype(F16C)::value, all_valid = decltype(FULLC)::value;
            auto pack_pair = [&](float a, float b) -> uint32_t {   // shadow the run-time versions
              if (F16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
              __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              return *reinterpret_cast<uint32_t*>(&h);
            };
            auto unpack_pair = [&](uint32_t w_) -> float2 {
              if (F16) return __half22float2(*reinterpret_cast<__half2*>(&w_));
              return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w_));
            };
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float sc0 = p.scale, sc1 = p.scale;
              if (DROP) {   // d P_dropped / d P = mask / (1 - p)
                const uint32_t hh = drop_pair(rk, (uint32_t)(key0 + c * 32 + i) >> 1);
                sc0 = drop_keep_lo(hh, p.drop_thresh) ? p.scale * p.drop_scale : 0.f;
                sc1 = drop_keep_hi(hh, p.drop_thresh) ? p.scale * p.drop_scale : 0.f;
              }
              float x0 = fmaf(__uint_as_float(dv[i]), sc0, nds), x1 = fmaf(__uint_as_float(dv[i + 1]), sc1, nds);
              if (!all_valid) {
                x0 = (c * 32 + i < nvalid) ? x0 : 0.f;
                x1 = (c * 32 + i + 1 < nvalid) ? x1 : 0.f;
              }
              const float2 pp = unpack_pair(pk[c * 16 + (i >> 1)]);
              w[i >> 1] = pack_pair(pp.x * x0, pp.y * x1);
            }
            if (DBG == 1) {   // diagnostics: all the arithmetic, no staging / stores
              uint32_t x = 0;
#pragma unroll
              for (int t = 0; t < 16; ++t) x ^= w[t];
              if (x == 0x12345678u) atomicAdd(reinterpret_cast<int*>(p.dS), 1);
              return;
            }
            const uint32_t slab = slab0 + flip * 4096;
            if ((c & 1) == 0) {   // the store issued from this slab two slabs ago must have been read
              if (lane == 0) { if (Cfg::NSLAB == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
              __syncwarp();
            }
            const uint32_t rowaddr = slab + lane * 128;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t a = rowaddr + ((((uint32_t)(c & 1) * 4 + t) ^ ((uint32_t)lane & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * t]), "r"(w[4 * t + 1]), "r"(w[4 * t + 2]), "r"(w[4 * t + 3]) : "memory");
            }
            if (c & 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmDS, slab, it.ds_col0 + key0 + (c >> 1) * 64, it.ds_row0 + q * 32);
                tma_store_commit();
              }
              if (Cfg::NSLAB == 2) flip ^= 1;
            }
          };
          auto emit = [&](const uint32_t (&dv)[32], int c) {
            switch (variant) {
              case 0: emit_t(dv, c, std::false_type{}, std::false_type{}); break;
              case 1: emit_t(dv, c, std::false_type{}, std::true_type{}); break;
              case 2: emit_t(dv, c, std::true_type{}, std::false_type{}); break;
              default: emit_t(dv, c, std::true_type{}, std::true_type{}); break;
            }
          };
          tmem_ld_32x32(d_addr, da);
          tmem_ld_wait();
          tmem_ld_32x32(d_addr + 32, db);
          emit(da, 0);
          tmem_ld_wait();
          tmem_ld_32x32(d_addr + 64, da);
          emit(db, 1);
          tmem_ld_wait();
          tmem_ld_32x32(d_addr + 96, db);
          emit(da, 2);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(dp_empty);
          emit(db, 3);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int CL, int DBG, bool DROP>
static int launch_ds_wide_d(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                          const CUtensorMap& tmDS, const AttnBwdArgs& a, cudaStream_t stream) {
  auto kern = attn_bwd_ds_wide_kernel<CL, DBG, DROP>;
  using WideCfg = WideCfgT<DBG>;
  CSN_SET_MAX_SMEM(kern, WideCfg::SMEM_BYTES);
  const int n_work = a.n_items / CL;
  const int workers = num_sms() / CL;
  const int grid = (n_work < workers ? n_work : workers) * CL;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = WideCfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmQ, tmDO, tmK, tmV, tmDS, a));
  CSN_LAUNCH_OK("attn_bwd_ds_wide_kernel");
  return 0;
}

// delta[(blk*h + head)*rows_pad + row] = sum_c dO[blk*rows_pad + row][head*d + c] * O[...]; one warp per (row, head)
__global__ void attn_delta_kernel(const void* __restrict__ dO, const void* __restrict__ O, const void* __restrict__ Olo,
                                  float* __restrict__ delta, long long rows, int rows_pad, int n_head, int d,
                                  long long ld, int dtype) {
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= rows * n_head) return;
  const long long row = wid / n_head;
  const int head = (int)(wid % n_head);
  const uint4* a = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(dO) + row * ld + head * d);
  const uint4* b = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(O) + row * ld + head * d);
  const uint4* bl = Olo ? reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(Olo) + row * ld + head * d) : nullptr;
  const float lo_inv = dtype == CSN_F16 ? 1.f / 2048.f : 1.f / 256.f;
  float s = 0.f, sl = 0.f;
  for (int c = lane; c < d / 8; c += 32) {   // 8 x 16-bit per 16-byte load
    const uint4 x4 = __ldg(a + c), y4 = __ldg(b + c);
    const uint4 z4 = bl ? __ldg(bl + c) : make_uint4(0u, 0u, 0u, 0u);
    const uint32_t xs[4] = {x4.x, x4.y, x4.z, x4.w}, ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 fx, fy, fz;
      if (dtype == CSN_F16) {
        fx = __half22float2(*reinterpret_cast<const __half2*>(&xs[k]));
        fy = __half22float2(*reinterpret_cast<const __half2*>(&ys[k]));
        fz = __half22float2(*reinterpret_cast<const __half2*>(&zs[k]));
      } else {
        fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[k]));
        fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[k]));
        fz = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zs[k]));
      }
      s += fx.x * fy.x + fx.y * fy.y;
      sl += fx.x * fz.x + fx.y * fz.y;
    }
  }
  s = warp_sum(s) + warp_sum(sl) * lo_inv;
  if (lane == 0) {
    const long long blk = row / rows_pad, rin = row % rows_pad;
    delta[(blk * n_head + head) * rows_pad + rin] = s;
  }
}

template <int DH, int CL, bool WITH_DQ, bool DROP>
static int launch_dq_d(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                     const CUtensorMap& tmDS, const CUtensorMap& tmDQ, const AttnBwdArgs& a, cudaStream_t stream) {
  using Cfg = BwdCfg<DH>;
  auto kern = attn_bwd_dq_kernel<DH, CL, WITH_DQ, DROP>;
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
  CSN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a));
  CSN_LAUNCH_OK("attn_bwd_dq_kernel");
  return 0;
}

// the dropout code is a compile-time variant: the eval-mode kernels carry none of it
template <int DH, int CL, bool WITH_DQ>
static int launch_dq(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                     const CUtensorMap& tmDS, const CUtensorMap& tmDQ, const AttnBwdArgs& a, cudaStream_t stream) {
  if (a.drop_thresh) return launch_dq_d<DH, CL, WITH_DQ, true>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, stream);
  return launch_dq_d<DH, CL, WITH_DQ, false>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, stream);
}
template <int CL, int DBG = 0>
static int launch_ds_wide(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                          const CUtensorMap& tmDS, const AttnBwdArgs& a, cudaStream_t stream) {
  if (a.drop_thresh) return launch_ds_wide_d<CL, DBG, true>(tmQ, tmDO, tmK, tmV, tmDS, a, stream);
  return launch_ds_wide_d<CL, DBG, false>(tmQ, tmDO, tmK, tmV, tmDS, a, stream);
}

int launch_dq64(const CUtensorMap& tmQ, const CUtensorMap& tmDO, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDQ, const int32_t* items, int32_t n_items, const float* lse, const float* delta,
                int32_t dtype, uint32_t drop_seed, float drop_p, cudaStream_t stream);   // attn_dkv.cu

}  // namespace csn

extern "C" {

int csn_attn_delta(const void* dO, const void* O, const void* O_lo, float* delta, int64_t rows, int32_t rows_pad, int32_t n_head,
                   int32_t d_head, int64_t ld, int32_t dtype, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(dO && O && delta, "csn_attn_delta: null pointer");
  CSN_CHECK_ARG(d_head % 64 == 0, "csn_attn_delta: d_head must be a multiple of 64");
  const long long nw = rows * n_head;
  if (nw == 0) return 0;
  const int wpb = 8;
  attn_delta_kernel<<<(unsigned)((nw + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(dO, O, O_lo, delta, rows, rows_pad, n_head, d_head, ld, dtype);
  CSN_LAUNCH_OK("attn_delta_kernel");
  return 0;
}

int csn_attn_bwd_dq(const void* Q, const void* dO, const void* K, const void* V, int64_t q_rows, int64_t do_rows,
                    int64_t kv_rows, int64_t width, int64_t ldq, int64_t lddo, int64_t ldk, int64_t ldv,
                    int32_t d_head, int32_t dtype, const int32_t* items, int32_t n_items, void* dQ, int64_t lddq,
                    void* dS, int64_t ds_rows, int64_t ldds, const float* lse, const float* delta, int32_t paired,
                    uint32_t drop_seed, float drop_p, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(Q && dO && K && V && items && (dS || dQ) && lse && delta, "csn_attn_bwd_dq: null pointer");
  CSN_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "csn_attn_bwd_dq: dropout probability %f outside [0, 1)", (double)drop_p);
  CSN_CHECK_ARG(d_head == 256 || d_head == 64, "csn_attn_bwd_dq: d_head=%d not supported (64 or 256)", d_head);
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "csn_attn_bwd_dq: 16-bit operands only");
  CSN_CHECK_ARG((lddq * 2) % 16 == 0 && (ldds * 2) % 16 == 0, "csn_attn_bwd_dq: output strides must be 16B multiples");
  if (n_items == 0) return 0;
  CUtensorMap tmQ, tmDO, tmK, tmV;
  int rc = make_tmap_2d(&tmQ, Q, dtype, width, q_rows, ldq, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmDO, dO, dtype, width, do_rows, lddo, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmK, K, dtype, width, kv_rows, ldk, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmV, V, dtype, width, kv_rows, ldv, 64, 128);
  if (rc) return rc;
  CUtensorMap tmDS, tmDQ;
  if (dS != nullptr) {   // dS == NULL (with dQ): nothing is materialised, dK / dV come from csn_attn_bwd_dkv
    rc = make_tmap_2d(&tmDS, dS, dtype, ldds, ds_rows, ldds, 64, 128);
    if (rc) return rc;
    tmDQ = tmDS;
  }
  if (dQ != nullptr) {   // dQ == NULL: only dS is produced (dQ = dS K then runs as a csn_gemm)
    rc = make_tmap_2d(&tmDQ, dQ, dtype, width, do_rows, lddq, 64, 32);
    if (rc) return rc;
    if (dS == nullptr) tmDS = tmDQ;
  }
  AttnBwdArgs a;
  a.items = reinterpret_cast<const AttnBwdItem*>(items);
  a.n_items = n_items;
  a.dQ = dQ; a.lddq = lddq; a.dS = dS; a.ldds = ldds;
  a.lse = lse; a.delta = delta;
  a.scale = 1.0f / sqrtf((float)d_head);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.dtype = dtype;
  a.drop_seed = drop_seed;
  a.drop_thresh = drop_thresh16(drop_p);
  a.drop_scale = drop_scale_of(a.drop_thresh);
  a.drop_epoch = a.drop_thresh ? drop_epoch_ptr() : nullptr;
  a.prefetch = 0;
  const uint32_t fmt = dtype == CSN_F16 ? 0u : 1u;
  a.idesc_s = umma_idesc_f16(fmt, 0, 0, 128);
  a.idesc_dq = umma_idesc_f16(fmt, 0, 1, d_head == 256 ? 128u : 64u);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool pair = paired && n_items % 2 == 0 && d_head == 256;
  if (dQ != nullptr && dS == nullptr && d_head == 64 && getenv("CSN_DQ64_TS") == nullptr)   // dS lives in TMEM only (attn_dkv.cu)
    return launch_dq64(tmQ, tmDO, tmK, tmV, tmDQ, items, n_items, lse, delta, dtype, drop_seed, drop_p, s);
  if (dQ != nullptr) {
    if (d_head == 256) return pair ? launch_dq<256, 2, true>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, s) : launch_dq<256, 1, true>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, s);
    return launch_dq<64, 1, true>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, s);
  }
  static const int wide = getenv("CSN_DS_WIDE") == nullptr ? 1 : atoi(getenv("CSN_DS_WIDE"));
  if (d_head == 256 && wide) {   // 256-key steps: N = 256 score tiles
    CUtensorMap tmDSw;
    rc = make_tmap_2d(&tmDSw, dS, dtype, ldds, ds_rows, ldds, 64, 32);
    if (rc) return rc;
    a.idesc_s = umma_idesc_f16(fmt, 0, 0, 256);
    static const int pf = getenv("CSN_DS_PREFETCH") == nullptr ? 0 : atoi(getenv("CSN_DS_PREFETCH"));
    a.prefetch = pf;
    if (wide == 2) return pair ? launch_ds_wide<2, 1>(tmQ, tmDO, tmK, tmV, tmDSw, a, s) : launch_ds_wide<1, 1>(tmQ, tmDO, tmK, tmV, tmDSw, a, s);   // diagnostics: no staging / stores
    if (wide == 3) return pair ? launch_ds_wide<2, 2>(tmQ, tmDO, tmK, tmV, tmDSw, a, s) : launch_ds_wide<1, 2>(tmQ, tmDO, tmK, tmV, tmDSw, a, s);   // 3-deep ring, double-buffered slabs
    return pair ? launch_ds_wide<2>(tmQ, tmDO, tmK, tmV, tmDSw, a, s) : launch_ds_wide<1>(tmQ, tmDO, tmK, tmV, tmDSw, a, s);
  }
  if (d_head == 256) return pair ? launch_dq<256, 2, false>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, s) : launch_dq<256, 1, false>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, s);
  return launch_dq<64, 1, false>(tmQ, tmDO, tmK, tmV, tmDS, tmDQ, a, s);
}

}  // extern "C"
