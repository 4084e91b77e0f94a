// Shape-compatibility retrieval measure + top-K (reference: MID-FC/csa_models.py:244-280,360-404,
// MinkowskiNet/models/hrnet.py:472-490, MinkowskiNet/lib/csn_utils.py:91-96).
//
//   score(A, B) = (1/N_A) * sum_p max_q  a^_p . b^_q        (rows L2-normalised)
//
// The reference materialises the N_A x N_B cosine matrix (400 MB at N = 10k) per pair; here it
// only ever exists as 128x256 fp32 tiles in TMEM:
//   csn_normalize_rows : fp32 rows -> unit-norm 16-bit rows (once per shape, not once per pair)
//   csn_knn_scores     : persistent tcgen05 kernel. A work item = 128 query rows (resident in SMEM
//                        for the whole item) x a list of candidate shapes whose rows are streamed by
//                        TMA; epilogue warps keep a running row-max in registers (lane = row),
//                        warp-shuffle the row sum and emit one partial per (item, candidate).
//   csn_knn_reduce     : fixed-order sum of the partials of one query -> score matrix (deterministic)
//   csn_topk_rows      : warp-level top-(K+1), sorted descending, int64 indices.
#include "host_util.h"
#include "ptx.cuh"

namespace csn {

constexpr int KNN_BM = 128;
constexpr int KNN_BN = 256;
constexpr int KNN_D = 256;                      // feature width on this path (d_model)
constexpr int KNN_KB = KNN_D / 64;              // 4 k-blocks of 64
constexpr int KNN_A_BYTES = KNN_BM * KNN_D * 2; // 64 KB resident query tile
constexpr int KNN_B_STAGE = KNN_BN * 64 * 2;    // 32 KB per streamed stage
constexpr int KNN_STAGES = 4;
constexpr int KNN_SMEM = KNN_A_BYTES + KNN_STAGES * KNN_B_STAGE + 1024 + 1024;
constexpr int KNN_THREADS = 256;

struct KnnItem {   // one work item (host/torch builds the table)
  int q_row0;      // first row of the 128-row query tile in the query feature buffer
  int n_valid;     // rows of the tile that belong to the query shape (<= 128)
  int list_begin;  // first entry of this item's candidate list
  int list_count;  // number of candidates
  int out_off;     // partial[out_off + pos] receives sum over valid rows of the row-max
  int pad;
};
struct KnnCand {
  int row0;  // first row of the candidate shape in the candidate feature buffer
  int len;   // number of points
};

struct KnnArgs {
  const KnnItem* items;
  const KnnCand* cands;
  float* partial;
  int n_items;
  uint32_t idesc;
};

__global__ void __launch_bounds__(KNN_THREADS, 1)
knn_score_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ KnnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + KNN_A_BYTES;
  const uint32_t bar_base = sB + KNN_STAGES * KNN_B_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (KNN_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * KNN_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * KNN_STAGES + 2 + a); };
  const uint32_t a_full = bar_base + 8u * (2 * KNN_STAGES + 4);
  const uint32_t a_empty = bar_base + 8u * (2 * KNN_STAGES + 5);
  uint8_t* bar_ptr = smem + KNN_A_BYTES + KNN_STAGES * KNN_B_STAGE;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * KNN_STAGES + 6));
  float* red = reinterpret_cast<float*>(bar_ptr + 256);  // [2][4] warp partial sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < KNN_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
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
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, a_ph = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const KnnItem item = p.items[it];
        mbar_wait(a_empty, a_ph ^ 1);  // previous item's MMAs have finished reading the query tile
        mbar_arrive_expect_tx(a_full, KNN_A_BYTES);
#pragma unroll
        for (int kb = 0; kb < KNN_KB; ++kb)
          tma_load_2d(sA + kb * (KNN_BM * 128), &tmQ, a_full, kb * 64, item.q_row0);
        a_ph ^= 1;
        for (int pos = 0; pos < item.list_count; ++pos) {
          const KnnCand c = p.cands[item.list_begin + pos];
          const int ntiles = (c.len + KNN_BN - 1) / KNN_BN;
          for (int nt = 0; nt < ntiles; ++nt) {
#pragma unroll 1
            for (int kb = 0; kb < KNN_KB; ++kb) {
              mbar_wait(empty_bar(st), ph ^ 1);
              mbar_arrive_expect_tx(full_bar(st), KNN_B_STAGE);
              tma_load_2d(sB + st * KNN_B_STAGE, &tmC, full_bar(st), kb * 64, c.row0 + nt * KNN_BN);
              if (++st == KNN_STAGES) { st = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int st = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0, a_ph = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const KnnItem item = p.items[it];
        mbar_wait(a_full, a_ph);
        a_ph ^= 1;
        tc_fence_after();
        for (int pos = 0; pos < item.list_count; ++pos) {
          const KnnCand c = p.cands[item.list_begin + pos];
          const int ntiles = (c.len + KNN_BN - 1) / KNN_BN;
          for (int nt = 0; nt < ntiles; ++nt) {
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * KNN_BN;
#pragma unroll 1
            for (int kb = 0; kb < KNN_KB; ++kb) {
              mbar_wait(full_bar(st), ph);
              tc_fence_after();
              const uint32_t a_tile = sA + kb * (KNN_BM * 128);
              const uint32_t b_tile = sB + st * KNN_B_STAGE;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16_ss(d_tmem, umma_desc_sw128(a_tile + k * 32, 0, 1024),
                            umma_desc_sw128(b_tile + k * 32, 0, 1024), p.idesc, (kb | k) ? 1u : 0u);
              }
              umma_commit(empty_bar(st));
              if (++st == KNN_STAGES) { st = 0; ph ^= 1; }
            }
            umma_commit(tfull_bar(acc));
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
          }
        }
        umma_commit(a_empty);  // query tile may be overwritten once everything above completed
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_ph = 0;
    int flip = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const KnnItem item = p.items[it];
      const bool row_valid = (q * 32 + lane) < item.n_valid;
      for (int pos = 0; pos < item.list_count; ++pos) {
        const KnnCand c = p.cands[item.list_begin + pos];
        const int ntiles = (c.len + KNN_BN - 1) / KNN_BN;
        float run = -INFINITY;
        for (int nt = 0; nt < ntiles; ++nt) {
          mbar_wait(tfull_bar(acc), acc_ph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * KNN_BN;
          const int ncols = min(KNN_BN, c.len - nt * KNN_BN);  // valid columns in this tile
          if (ncols == KNN_BN) {
            // two 32-column loads in flight per wait
#pragma unroll 1
            for (int cc = 0; cc < KNN_BN; cc += 64) {
              uint32_t r0[32], r1[32];
              tmem_ld_32x32(taddr + cc, r0);
              tmem_ld_32x32(taddr + cc + 32, r1);
              tmem_ld_wait();
              float m0 = run, m1 = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                m0 = fmaxf(m0, __uint_as_float(r0[j]));
                m1 = fmaxf(m1, __uint_as_float(r1[j]));
              }
              run = fmaxf(m0, m1);
            }
          } else {
#pragma unroll 1
            for (int cc = 0; cc < ncols; cc += 32) {
              uint32_t r0[32];
              tmem_ld_32x32(taddr + cc, r0);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cc + j < ncols) run = fmaxf(run, __uint_as_float(r0[j]));
            }
          }
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
          if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        // sum over the valid rows of this tile: warp shuffle, then the 4 epilogue warps in a fixed order
        const float s = warp_sum(row_valid ? run : 0.f);
        if (lane == 0) red[flip * 4 + q] = s;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 4 && lane == 0) {
          const float tot = (red[flip * 4 + 0] + red[flip * 4 + 1]) + (red[flip * 4 + 2] + red[flip * 4 + 3]);
          p.partial[item.out_off + pos] = tot;
        }
        flip ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// --------------------------------------------------------------------------- exact re-scoring
// Same pipeline with split operands: every unit-norm feature x is stored as hi = fp16(x) and
// lo = fp16((x - hi) * 2^11), i.e. a 22-bit significand.  Per tile two accumulators are kept,
//   acc0 = A_hi B_hi^T ,  acc1 = A_hi B_lo^T + A_lo B_hi^T ,   cos = acc0 + acc1 * 2^-11
// (the lo*lo term is < 2^-22 and dropped), which reproduces the reference's fp32 cosines to ~1e-7.
// It re-scores only the (query, candidate) pairs whose coarse score lies in the band around the
// (K+1)-th best, so that top-K index sets match the reference except for ties below 1e-6.
constexpr int KNX_BN = 128;
constexpr int KNX_A_BYTES = 2 * KNN_BM * KNN_D * 2;   // hi + lo query tiles, 128 KB resident
constexpr int KNX_B_STAGE = 2 * KNX_BN * 64 * 2;      // hi + lo k-block of 128 candidate rows, 32 KB
constexpr int KNX_STAGES = 3;
constexpr int KNX_SMEM = KNX_A_BYTES + KNX_STAGES * KNX_B_STAGE + 1024 + 1024;

__global__ void __launch_bounds__(KNN_THREADS, 1)
knn_score_exact_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                       const __grid_constant__ CUtensorMap tmCh, const __grid_constant__ CUtensorMap tmCl,
                       const __grid_constant__ KnnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sAh = smem_u32(smem);
  const uint32_t sAl = sAh + KNN_BM * KNN_D * 2;
  const uint32_t sB = sAh + KNX_A_BYTES;
  const uint32_t bar_base = sB + KNX_STAGES * KNX_B_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (KNX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * KNX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * KNX_STAGES + 2 + a); };
  const uint32_t a_full = bar_base + 8u * (2 * KNX_STAGES + 4);
  const uint32_t a_empty = bar_base + 8u * (2 * KNX_STAGES + 5);
  uint8_t* bar_ptr = smem + KNX_A_BYTES + KNX_STAGES * KNX_B_STAGE;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * KNX_STAGES + 6));
  float* red = reinterpret_cast<float*>(bar_ptr + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < KNX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
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
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0, a_ph = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const KnnItem item = p.items[it];
        mbar_wait(a_empty, a_ph ^ 1);
        mbar_arrive_expect_tx(a_full, KNX_A_BYTES);
#pragma unroll
        for (int kb = 0; kb < KNN_KB; ++kb) {
          tma_load_2d(sAh + kb * (KNN_BM * 128), &tmQh, a_full, kb * 64, item.q_row0);
          tma_load_2d(sAl + kb * (KNN_BM * 128), &tmQl, a_full, kb * 64, item.q_row0);
        }
        a_ph ^= 1;
        for (int pos = 0; pos < item.list_count; ++pos) {
          const KnnCand c = p.cands[item.list_begin + pos];
          const int ntiles = (c.len + KNX_BN - 1) / KNX_BN;
          for (int nt = 0; nt < ntiles; ++nt) {
#pragma unroll 1
            for (int kb = 0; kb < KNN_KB; ++kb) {
              mbar_wait(empty_bar(st), ph ^ 1);
              mbar_arrive_expect_tx(full_bar(st), KNX_B_STAGE);
              tma_load_2d(sB + st * KNX_B_STAGE, &tmCh, full_bar(st), kb * 64, c.row0 + nt * KNX_BN);
              tma_load_2d(sB + st * KNX_B_STAGE + KNX_BN * 128, &tmCl, full_bar(st), kb * 64, c.row0 + nt * KNX_BN);
              if (++st == KNX_STAGES) { st = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int st = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0, a_ph = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const KnnItem item = p.items[it];
        mbar_wait(a_full, a_ph);
        a_ph ^= 1;
        tc_fence_after();
        for (int pos = 0; pos < item.list_count; ++pos) {
          const KnnCand c = p.cands[item.list_begin + pos];
          const int ntiles = (c.len + KNX_BN - 1) / KNX_BN;
          for (int nt = 0; nt < ntiles; ++nt) {
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + acc * 2 * KNX_BN;   // hi*hi
            const uint32_t d1 = d0 + KNX_BN;                    // cross terms (scaled by 2^11)
#pragma unroll 1
            for (int kb = 0; kb < KNN_KB; ++kb) {
              mbar_wait(full_bar(st), ph);
              tc_fence_after();
              const uint32_t ah = sAh + kb * (KNN_BM * 128), al = sAl + kb * (KNN_BM * 128);
              const uint32_t bh = sB + st * KNX_B_STAGE, bl = bh + KNX_BN * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t dah = umma_desc_sw128(ah + k * 32, 0, 1024), dal = umma_desc_sw128(al + k * 32, 0, 1024);
                const uint64_t dbh = umma_desc_sw128(bh + k * 32, 0, 1024), dbl = umma_desc_sw128(bl + k * 32, 0, 1024);
                umma_f16_ss(d0, dah, dbh, p.idesc, (kb | k) ? 1u : 0u);
                umma_f16_ss(d1, dah, dbl, p.idesc, (kb | k) ? 1u : 0u);
                umma_f16_ss(d1, dal, dbh, p.idesc, 1u);
              }
              umma_commit(empty_bar(st));
              if (++st == KNX_STAGES) { st = 0; ph ^= 1; }
            }
            umma_commit(tfull_bar(acc));
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
          }
        }
        umma_commit(a_empty);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_ph = 0;
    int flip = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const KnnItem item = p.items[it];
      const bool row_valid = (q * 32 + lane) < item.n_valid;
      for (int pos = 0; pos < item.list_count; ++pos) {
        const KnnCand c = p.cands[item.list_begin + pos];
        const int ntiles = (c.len + KNX_BN - 1) / KNX_BN;
        float run = -INFINITY;
        for (int nt = 0; nt < ntiles; ++nt) {
          mbar_wait(tfull_bar(acc), acc_ph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * 2 * KNX_BN;
          const int ncols = min(KNX_BN, c.len - nt * KNX_BN);
#pragma unroll 1
          for (int cc = 0; cc < ncols; cc += 32) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr + cc, r0);
            tmem_ld_32x32(taddr + KNX_BN + cc, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cc + j < ncols) run = fmaxf(run, __uint_as_float(r0[j]) + __uint_as_float(r1[j]) * (1.f / 2048.f));
          }
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
          if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        const float s = warp_sum(row_valid ? run : 0.f);
        if (lane == 0) red[flip * 4 + q] = s;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 4 && lane == 0)
          p.partial[item.out_off + pos] = (red[flip * 4 + 0] + red[flip * 4 + 1]) + (red[flip * 4 + 2] + red[flip * 4 + 3]);
        flip ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out_hi = fp16(x / max(|x|, eps)), out_lo = fp16((x/|x| - out_hi) * 2^11): the split unit-norm rows
__global__ void normalize_rows_split_kernel(const float* __restrict__ in, __half* __restrict__ out_hi,
                                            __half* __restrict__ out_lo, long long rows, int D, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = in + row * D;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = __ldg(src + c); ss += v * v; }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int c = lane; c < D; c += 32) {
    const float v = __ldg(src + c) * inv;
    const __half h = __float2half_rn(v);
    out_hi[row * D + c] = h;
    out_lo[row * D + c] = __float2half_rn((v - __half2float(h)) * 2048.f);
  }
}

// --------------------------------------------------------------------------- normalise rows
// One warp per row of D floats: out = v / max(|v|_2, eps) as 16-bit. D must be a multiple of 128.
__global__ void normalize_rows_kernel(const float* __restrict__ in, void* __restrict__ out, long long rows,
                                      int D, float eps, int out_dtype) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(in + row * D);
  float4 v[4];  // D <= 512
  const int nvec = D / 128;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < nvec) {
      v[i] = __ldg(src + i * 32 + lane);
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + row * D);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < nvec) {
      uint2 o;
      if (out_dtype == CSN_F16) {
        __half2 a = __floats2half2_rn(v[i].x * inv, v[i].y * inv), b = __floats2half2_rn(v[i].z * inv, v[i].w * inv);
        o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[i].x * inv, v[i].y * inv), b = __floats2bfloat162_rn(v[i].z * inv, v[i].w * inv);
        o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
      }
      dst[i * 32 + lane] = o;
    }
  }
}

// --------------------------------------------------------------------------- reduce partials
// scores[q][c] = (sum_{t < ntiles} partial[(q*ntiles + t)*n_cand + c]) / n_rows   (fixed order)
__global__ void knn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ scores, int n_q,
                                  int n_cand, int ntiles, float inv_rows, long long ld_scores) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_q * n_cand) return;
  const int q = (int)(idx / n_cand), c = (int)(idx % n_cand);
  const float* src = partial + ((long long)q * ntiles) * n_cand + c;
  float s = 0.f;
  for (int t = 0; t < ntiles; ++t) s += src[(long long)t * n_cand];
  scores[(long long)q * ld_scores + c] = s * inv_rows;
}

// --------------------------------------------------------------------------- top-K per row
// One warp per row. Each lane keeps its own sorted top-KK (KK <= TOPK_MAX) over the columns it strides,
// then the warp extracts the global top-KK by KK rounds of shuffle arg-max. Ties: lower index wins.
// TOPK_MAX is a template parameter (8 / 16 / 32 / 64 list entries per lane): the reference's launcher default
// K = 10 (run_save_knn.py:34) asks for top-11.
constexpr int TOPK_LIMIT = 64;
template <int TOPK_MAX>
__global__ void topk_rows_kernel(const float* __restrict__ scores, long long ld, int n_rows, int n_cols, int kk,
                                 float* __restrict__ out_val, long long* __restrict__ out_idx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float v[TOPK_MAX];
  int ix[TOPK_MAX];
#pragma unroll
  for (int i = 0; i < TOPK_MAX; ++i) { v[i] = -INFINITY; ix[i] = 0x7fffffff; }
  const float* src = scores + (long long)row * ld;
  for (int c = lane; c < n_cols; c += 32) {
    float x = src[c];
    int xi = c;
    // insert into the sorted list (descending), bubbling the displaced entries down
#pragma unroll
    for (int i = 0; i < TOPK_MAX; ++i) {
      if (i < kk) {
        const bool better = (x > v[i]) || (x == v[i] && xi < ix[i]);
        if (better) {
          const float tv = v[i]; const int ti = ix[i];
          v[i] = x; ix[i] = xi; x = tv; xi = ti;
        }
      }
    }
  }
  for (int r = 0; r < kk; ++r) {
    float bv = v[0];
    int bi = ix[0];
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bl = ol; }
    }
    if (elect_one()) {
      out_val[(long long)row * kk + r] = bv;
      out_idx[(long long)row * kk + r] = bi;
    }
    if (lane == bl) {  // pop the winner from its owner's list
#pragma unroll
      for (int i = 0; i < TOPK_MAX - 1; ++i) { v[i] = v[i + 1]; ix[i] = ix[i + 1]; }
      v[TOPK_MAX - 1] = -INFINITY; ix[TOPK_MAX - 1] = 0x7fffffff;
    }
  }
}

// --------------------------------------------------------------------------- top-K boundary band, on the device
// csn_knn_band_select: per query row, the k-th largest coarse score and every candidate within `margin` of it (or
// above) — the pairs whose order the 16-bit scoring pass cannot be trusted on — written directly as the work tables of
// csn_knn_scores_exact: candidate list of query q at cands[q*n_cols ...] (ascending candidate index), one item per
// 128-row tile of the query with list_count = band size.  Nothing returns to the host (the reference's
// `retrieval_measure.topk(K+1)` csa_models.py:278 is evaluated on scores that are exact where it matters).
struct BandArgs {
  const float* scores; long long ld; int n_q, n_cols, k; float margin;
  const int* q_row0; const int* q_len; const int* item0;   // per query: first feature row, length, first item index
  const int* c_row0; const int* c_len;                      // per candidate shape
  int* band_idx; int* counts; KnnCand* cands; KnnItem* items;
};

template <int TOPK_MAX>
__global__ void knn_band_select_kernel(const BandArgs p) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= p.n_q) return;
  float v[TOPK_MAX];
  int ix[TOPK_MAX];
#pragma unroll
  for (int i = 0; i < TOPK_MAX; ++i) { v[i] = -INFINITY; ix[i] = 0x7fffffff; }
  const float* src = p.scores + (long long)q * p.ld;
  for (int c = lane; c < p.n_cols; c += 32) {
    float x = src[c];
    int xi = c;
#pragma unroll
    for (int i = 0; i < TOPK_MAX; ++i) {
      if (i < p.k) {
        const bool better = (x > v[i]) || (x == v[i] && xi < ix[i]);
        if (better) { const float tv = v[i]; const int ti = ix[i]; v[i] = x; ix[i] = xi; x = tv; xi = ti; }
      }
    }
  }
  float kth = 0.f;
  for (int r = 0; r < p.k; ++r) {
    float bv = v[0]; int bi = ix[0]; int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bl = ol; }
    }
    kth = bv;
    if (lane == bl) {
#pragma unroll
      for (int i = 0; i < TOPK_MAX - 1; ++i) { v[i] = v[i + 1]; ix[i] = ix[i + 1]; }
      v[TOPK_MAX - 1] = -INFINITY; ix[TOPK_MAX - 1] = 0x7fffffff;
    }
  }
  const float thr = kth - p.margin;
  int cnt = 0;
  const long long base = (long long)q * p.n_cols;
  for (int c0 = 0; c0 < p.n_cols; c0 += 32) {
    const int c = c0 + lane;
    const bool in = c < p.n_cols && src[c] >= thr;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int pos = cnt + __popc(m & ((1u << lane) - 1u));
      p.band_idx[base + pos] = c;
      p.cands[base + pos] = KnnCand{p.c_row0[c], p.c_len[c]};
    }
    cnt += __popc(m);
  }
  const int len = p.q_len[q], nt = (len + KNN_BM - 1) / KNN_BM, it0 = p.item0[q];
  if (lane == 0) p.counts[q] = cnt;
  for (int t = lane; t < nt; t += 32) {
    KnnItem it;
    it.q_row0 = p.q_row0[q] + t * KNN_BM;
    it.n_valid = min(KNN_BM, len - t * KNN_BM);
    it.list_begin = (int)base;
    it.list_count = cnt;
    it.out_off = (it0 + t) * p.n_cols;
    it.pad = 0;
    p.items[it0 + t] = it;
  }
}

// scores[q][band_idx[q][i]] = (sum over the query's row tiles of partial[(item0[q] + t)*n_cols + i]) / len_q, tiles summed
// in a fixed order in fp64
__global__ void knn_band_patch_kernel(const float* __restrict__ partial, const int* __restrict__ band_idx,
                                      const int* __restrict__ counts, const int* __restrict__ item0,
                                      const int* __restrict__ q_len, int n_q, int n_cols, float* __restrict__ scores,
                                      long long ld) {
  const int q = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_q || i >= counts[q]) return;
  const int len = q_len[q], nt = (len + KNN_BM - 1) / KNN_BM;
  double s = 0.0;
  for (int t = 0; t < nt; ++t) s += (double)partial[(long long)(item0[q] + t) * n_cols + i];
  scores[(long long)q * ld + band_idx[(long long)q * n_cols + i]] = (float)(s / (double)len);
}

}  // namespace csn

extern "C" {

int csn_normalize_rows(const float* in, void* out, int64_t rows, int32_t D, float eps, int32_t out_dtype,
                       void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(in && out, "csn_normalize_rows: null pointer");
  CSN_CHECK_ARG(D % 128 == 0 && D <= 512, "csn_normalize_rows: D=%d must be a multiple of 128, <= 512", D);
  CSN_CHECK_ARG(out_dtype == CSN_F16 || out_dtype == CSN_BF16, "csn_normalize_rows: 16-bit output only");
  if (rows == 0) return 0;
  const int wpb = 8;
  normalize_rows_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(in, out, rows, D, eps, out_dtype);
  CSN_LAUNCH_OK("normalize_rows_kernel");
  return 0;
}

int csn_knn_scores(const void* feat_q, int64_t rows_q, const void* feat_c, int64_t rows_c, int32_t dtype,
                   const int32_t* items, int32_t n_items, const int32_t* cands, float* partial,
                   void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(feat_q && feat_c && items && cands && partial, "csn_knn_scores: null pointer");
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "csn_knn_scores: features must be f16/bf16");
  if (n_items == 0) return 0;
  CUtensorMap tmQ, tmC;
  int rc = make_tmap_2d(&tmQ, feat_q, dtype, KNN_D, rows_q, KNN_D, 64, KNN_BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmC, feat_c, dtype, KNN_D, rows_c, KNN_D, 64, KNN_BN);
  if (rc) return rc;
  KnnArgs a;
  a.items = reinterpret_cast<const KnnItem*>(items);
  a.cands = reinterpret_cast<const KnnCand*>(cands);
  a.partial = partial;
  a.n_items = n_items;
  a.idesc = umma_idesc_f16(dtype == CSN_F16 ? 0u : 1u, 0, 0, KNN_BN);
  CSN_SET_MAX_SMEM(knn_score_kernel, KNN_SMEM);
  const int grid = n_items < num_sms() ? n_items : num_sms();
  knn_score_kernel<<<grid, KNN_THREADS, KNN_SMEM, (cudaStream_t)stream>>>(tmQ, tmC, a);
  CSN_LAUNCH_OK("knn_score_kernel");
  return 0;
}

int csn_normalize_rows_split(const float* in, void* out_hi, void* out_lo, int64_t rows, int32_t D, float eps,
                             void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(in && out_hi && out_lo, "csn_normalize_rows_split: null pointer");
  if (rows == 0) return 0;
  const int wpb = 8;
  normalize_rows_split_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), rows, D, eps);
  CSN_LAUNCH_OK("normalize_rows_split_kernel");
  return 0;
}

int csn_knn_scores_exact(const void* q_hi, const void* q_lo, int64_t rows_q, const void* c_hi, const void* c_lo,
                         int64_t rows_c, const int32_t* items, int32_t n_items, const int32_t* cands, float* partial,
                         void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(q_hi && q_lo && c_hi && c_lo && items && cands && partial, "csn_knn_scores_exact: null pointer");
  if (n_items == 0) return 0;
  CUtensorMap tmQh, tmQl, tmCh, tmCl;
  int rc = make_tmap_2d(&tmQh, q_hi, CSN_F16, KNN_D, rows_q, KNN_D, 64, KNN_BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmQl, q_lo, CSN_F16, KNN_D, rows_q, KNN_D, 64, KNN_BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmCh, c_hi, CSN_F16, KNN_D, rows_c, KNN_D, 64, KNX_BN);
  if (rc) return rc;
  rc = make_tmap_2d(&tmCl, c_lo, CSN_F16, KNN_D, rows_c, KNN_D, 64, KNX_BN);
  if (rc) return rc;
  KnnArgs a;
  a.items = reinterpret_cast<const KnnItem*>(items);
  a.cands = reinterpret_cast<const KnnCand*>(cands);
  a.partial = partial;
  a.n_items = n_items;
  a.idesc = umma_idesc_f16(0u, 0, 0, KNX_BN);
  CSN_SET_MAX_SMEM(knn_score_exact_kernel, KNX_SMEM);
  const int grid = n_items < num_sms() ? n_items : num_sms();
  knn_score_exact_kernel<<<grid, KNN_THREADS, KNX_SMEM, (cudaStream_t)stream>>>(tmQh, tmQl, tmCh, tmCl, a);
  CSN_LAUNCH_OK("knn_score_exact_kernel");
  return 0;
}

int csn_knn_reduce(const float* partial, float* scores, int32_t n_q, int32_t n_cand, int32_t ntiles,
                   int32_t n_rows, int64_t ld_scores, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(partial && scores, "csn_knn_reduce: null pointer");
  const long long n = (long long)n_q * n_cand;
  if (n == 0) return 0;
  knn_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, scores, n_q, n_cand, ntiles, 1.0f / (float)n_rows, ld_scores);
  CSN_LAUNCH_OK("knn_reduce_kernel");
  return 0;
}

int csn_topk_rows(const float* scores, int64_t ld, int32_t n_rows, int32_t n_cols, int32_t k, float* out_val,
                  int64_t* out_idx, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(scores && out_val && out_idx, "csn_topk_rows: null pointer");
  CSN_CHECK_ARG(k >= 1 && k <= TOPK_LIMIT, "csn_topk_rows: k=%d must be in [1,%d]", k, TOPK_LIMIT);
  CSN_CHECK_ARG(k <= n_cols, "csn_topk_rows: k=%d exceeds the number of columns %d", k, n_cols);
  if (n_rows == 0) return 0;
  const int wpb = 4;
  const dim3 grid((n_rows + wpb - 1) / wpb), block(wpb * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 8) topk_rows_kernel<8><<<grid, block, 0, st>>>(scores, ld, n_rows, n_cols, k, out_val, (long long*)out_idx);
  else if (k <= 16) topk_rows_kernel<16><<<grid, block, 0, st>>>(scores, ld, n_rows, n_cols, k, out_val, (long long*)out_idx);
  else if (k <= 32) topk_rows_kernel<32><<<grid, block, 0, st>>>(scores, ld, n_rows, n_cols, k, out_val, (long long*)out_idx);
  else topk_rows_kernel<64><<<grid, block, 0, st>>>(scores, ld, n_rows, n_cols, k, out_val, (long long*)out_idx);
  CSN_LAUNCH_OK("topk_rows_kernel");
  return 0;
}

int csn_knn_band_select(const float* scores, int64_t ld, int32_t n_q, int32_t n_cols, int32_t k, float margin,
                        const int32_t* q_row0, const int32_t* q_len, const int32_t* item0, const int32_t* c_row0,
                        const int32_t* c_len, int32_t* band_idx, int32_t* counts, int32_t* cands, int32_t* items,
                        void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(scores && q_row0 && q_len && item0 && c_row0 && c_len && band_idx && counts && cands && items,
                "csn_knn_band_select: null pointer");
  CSN_CHECK_ARG(k >= 1 && k <= TOPK_LIMIT && k <= n_cols, "csn_knn_band_select: k=%d must be in [1, min(%d, n_cols=%d)]", k, TOPK_LIMIT, n_cols);
  CSN_CHECK_ARG((long long)n_q * n_cols < (1ll << 31), "csn_knn_band_select: query block too large (n_q*n_cols >= 2^31)");
  if (n_q == 0) return 0;
  BandArgs a{scores, ld, n_q, n_cols, k, margin, q_row0, q_len, item0, c_row0, c_len, band_idx, counts,
             reinterpret_cast<KnnCand*>(cands), reinterpret_cast<KnnItem*>(items)};
  const int wpb = 4;
  const dim3 grid((n_q + wpb - 1) / wpb), block(wpb * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 8) knn_band_select_kernel<8><<<grid, block, 0, st>>>(a);
  else if (k <= 16) knn_band_select_kernel<16><<<grid, block, 0, st>>>(a);
  else if (k <= 32) knn_band_select_kernel<32><<<grid, block, 0, st>>>(a);
  else knn_band_select_kernel<64><<<grid, block, 0, st>>>(a);
  CSN_LAUNCH_OK("knn_band_select_kernel");
  return 0;
}

int csn_knn_band_patch(const float* partial, const int32_t* band_idx, const int32_t* counts, const int32_t* item0,
                       const int32_t* q_len, int32_t n_q, int32_t n_cols, float* scores, int64_t ld, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(partial && band_idx && counts && item0 && q_len && scores, "csn_knn_band_patch: null pointer");
  if (n_q == 0 || n_cols == 0) return 0;
  knn_band_patch_kernel<<<dim3((n_cols + 127) / 128, n_q), 128, 0, (cudaStream_t)stream>>>(partial, band_idx, counts, item0, q_len,
                                                                                           n_q, n_cols, scores, ld);
  CSN_LAUNCH_OK("knn_band_patch_kernel");
  return 0;
}

}  // extern "C"
