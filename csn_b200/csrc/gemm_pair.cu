// D[M x N] = alpha * A[M x K] * B[N x K]^T with CTA PAIRS: one tcgen05.mma.cta_group::2 of M = 256 per two SMs.
//
// Each CTA of a 2-CTA cluster owns 128 rows of a 256-row output tile.  Per k-block it loads its own [128 x 64]
// slice of A and HALF of the [BN x 64] slice of B ([BN/2 x 64]) — 32 KB per stage instead of 48 KB, so the ring is
// 6 deep in the same SMEM, and each SM reads half of the B operand per MMA.  The leader CTA (rank 0) issues the
// MMAs for the pair; TMA completion bytes of both CTAs are counted on the leader's full barrier; tcgen05.commit
// (cta_group::2, multicast) releases the stage in both CTAs and publishes the accumulator to both epilogues; the
// peer's epilogue warps free the accumulator with remote mbarrier arrives.
// K-major 16-bit operands, row-major fp32 / 16-bit output through TMA-store slabs.  Used by csn_gemm for plain
// (unbatched, non-split) problems; also the testbed of the pair protocol for the attention kernels.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace csn {

struct PairArgs {
  int M, N, K;
  int tiles_m2, tiles_n;   // tiles of 256 rows x BN columns
  int kb_total;
  float alpha;
  uint32_t idesc;
  int out_dtype;
};

constexpr int PBN = 256, PBK = 64, PSTAGES = 6;
constexpr int PA_BYTES = 128 * PBK * 2, PB_BYTES = (PBN / 2) * PBK * 2, PSTAGE_BYTES = PA_BYTES + PB_BYTES;
constexpr int PSTG_BYTES = 4 * 2 * 4096;
constexpr int PSMEM_BYTES = PSTAGES * PSTAGE_BYTES + PSTG_BYTES + 256 + 1024;

__global__ void __launch_bounds__(256, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ PairArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t stg_base = smem_base + PSTAGES * PSTAGE_BYTES;
  const uint32_t bar_base = stg_base + PSTG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PSTAGES + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + PSTAGES * PSTAGE_BYTES + PSTG_BYTES + 8 * (2 * PSTAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }   // both CTAs' epilogue threads
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const long long total = (long long)p.tiles_m2 * p.tiles_n;
  const long long t_first = blockIdx.x / 2, t_stride = gridDim.x / 2;

  if (warp == 0) {
    // ================================================================== TMA producer (both CTAs)
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (long long t = t_first; t < total; t += t_stride) {
        const int mt2 = (int)(t / p.tiles_n), nt = (int)(t % p.tiles_n);
        const int row_a = mt2 * 256 + rank * 128;
        const int row_b = nt * PBN + rank * (PBN / 2);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(empty_bar(st), ph ^ 1);
          const uint32_t lbar = mapa_shared(full_bar(st), 0);   // the leader's barrier counts both CTAs' bytes
          if (rank == 0) mbar_arrive_expect_tx(full_bar(st), 2 * PSTAGE_BYTES);
          const uint32_t sA = smem_base + st * PSTAGE_BYTES, sB = sA + PA_BYTES;
          tma_load_2d_pair(sA, &tmA, lbar, kb * PBK, row_a);
          tma_load_2d_pair(sB, &tmB, lbar, kb * PBK, row_b);
          if (++st == PSTAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader only)
    if (rank == 0 && elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (long long t = t_first; t < total; t += t_stride) {
        mbar_wait(tempty_bar(acc), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PBN;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(full_bar(st), ph);
          tc_fence_after();
          const uint32_t sA = smem_base + st * PSTAGE_BYTES, sB = sA + PA_BYTES;
#pragma unroll
          for (int k = 0; k < PBK / 16; ++k)
            umma_f16_ss2(d_tmem, umma_desc_sw128(sA + k * 32, 0, 1024), umma_desc_sw128(sB + k * 32, 0, 1024), p.idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit2_mc(empty_bar(st), 0b11);   // the stage is free in both CTAs
          if (++st == PSTAGES) { st = 0; ph ^= 1; }
        }
        umma_commit2_mc(tfull_bar(acc), 0b11);     // accumulator complete in both CTAs' TMEM
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== epilogue (each CTA: its own 128 rows)
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_ph = 0;
    int flip = 0;
    const uint32_t leader_tempty0 = mapa_shared(tempty_bar(0), 0), leader_tempty1 = mapa_shared(tempty_bar(1), 0);
    const bool o32 = p.out_dtype == CSN_F32;
    const int W = o32 ? 32 : 64;
    for (long long t = t_first; t < total; t += t_stride) {
      const int mt2 = (int)(t / p.tiles_n), nt = (int)(t % p.tiles_n);
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const int row0 = mt2 * 256 + rank * 128 + q * 32;
      const int n0 = nt * PBN;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * PBN;
      const int n_slabs = min(PBN / W, (p.N - n0 + W - 1) / W);
#pragma unroll 1
      for (int sl = 0; sl < n_slabs; ++sl) {
        uint32_t w[32];
        if (o32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + sl * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(__uint_as_float(r[j]) * p.alpha);
        } else {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + sl * 64, r0);
          tmem_ld_32x32(taddr + sl * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a0 = __uint_as_float(r0[2 * j]) * p.alpha, a1 = __uint_as_float(r0[2 * j + 1]) * p.alpha;
            const float b0 = __uint_as_float(r1[2 * j]) * p.alpha, b1 = __uint_as_float(r1[2 * j + 1]) * p.alpha;
            if (p.out_dtype == CSN_F16) {
              __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(b0, b1);
              w[j] = *reinterpret_cast<uint32_t*>(&h0); w[16 + j] = *reinterpret_cast<uint32_t*>(&h1);
            } else {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(a0, a1), h1 = __floats2bfloat162_rn(b0, b1);
              w[j] = *reinterpret_cast<uint32_t*>(&h0); w[16 + j] = *reinterpret_cast<uint32_t*>(&h1);
            }
          }
        }
        const uint32_t buf = stg_base + (q * 2 + flip) * 4096;
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
          tma_store_2d(&tmD, buf, n0 + sl * W, row0);
          tma_store_commit();
        }
        flip ^= 1;
      }
      tc_fence_before();
      mbar_arrive_cluster(acc ? leader_tempty1 : leader_tempty0);   // the leader's MMA warp waits for both CTAs
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace csn

// Experimental entry point (also reachable through csn_gemm): K-major A [M x K], B [N x K] (16-bit, leading
// dimensions lda / ldb), row-major D [M x N] (leading dimension ldd) of dtype out_dtype.
extern "C" int csn_gemm_pair(const void* A, const void* B, void* D, int32_t M, int32_t N, int32_t K, int64_t lda,
                             int64_t ldb, int64_t ldd, int32_t dtype, int32_t out_dtype, float alpha, void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(A && B && D, "csn_gemm_pair: null pointer");
  CSN_CHECK_ARG(M > 0 && N > 0 && K > 0, "csn_gemm_pair: empty problem");
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "csn_gemm_pair: operands must be f16/bf16");
  CUtensorMap tmA, tmB, tmD;
  int rc = make_tmap_2d(&tmA, A, dtype, K, M, lda, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, B, dtype, K, N, ldb, 64, PBN / 2);
  if (rc) return rc;
  rc = make_tmap_2d_any(&tmD, D, out_dtype, N, M, ldd, out_dtype == CSN_F32 ? 32 : 64, 32);
  if (rc) return rc;
  PairArgs a;
  a.M = M; a.N = N; a.K = K;
  a.tiles_m2 = (M + 255) / 256;
  a.tiles_n = (N + PBN - 1) / PBN;
  a.kb_total = (K + PBK - 1) / PBK;
  a.alpha = alpha;
  a.idesc = umma_idesc_f16(dtype == CSN_F16 ? 0u : 1u, 0, 0, PBN, 256);
  a.out_dtype = out_dtype;
  CSN_SET_MAX_SMEM(gemm_pair_kernel, PSMEM_BYTES);
  const long long total = (long long)a.tiles_m2 * a.tiles_n;
  const long long pairs = num_sms() / 2;
  const long long grid = (total < pairs ? total : pairs) * 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = PSMEM_BYTES;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSN_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_pair_kernel, tmA, tmB, tmD, a));
  CSN_LAUNCH_OK("gemm_pair_kernel");
  return 0;
}
