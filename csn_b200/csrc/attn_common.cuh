// Types shared by the fused attention kernels (attn_fwd.cu, attn_fwd2.cu).
#pragma once
#include "host_util.h"
#include "ptx.cuh"

namespace csn {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnItem {
  int q_row0;    // first query row (row of the Q view)
  int q_valid;   // rows of the tile that exist (<= 128); the others are written as zeros
  int kv_row0;   // first key/value row
  int kv_len;    // number of keys this tile attends to
  int o_row0;    // first output row (row of the O buffer)
  int col0;      // head * d: column offset in the Q/K/V views and in O
  int lse_off;   // lse[lse_off + r] for row r of the tile
  int flags;     // bit 0: write zeros to the tile rows >= q_valid (padded layouts); else leave them untouched
  int v_row0;    // first row of the MN-major streamed operand (V; dO in dV mode) matching kv_row0
  int key0;      // dV mode: index of the tile's first resident key inside its chunk (dropout mask column); else 0
};

struct AttnFwdArgs {
  const AttnItem* items;
  int n_items;
  void* O;            // 16-bit [rows][ldo]
  void* Olo;          // optional 16-bit residual (o - round16(o)) * 2^11 (fp16) / 2^8 (bf16), same layout as O:
                      // lets the backward pass form delta = rowsum(dO o O) to ~22 bits
  long long ldo;
  float* lse;         // natural-log sum-exp of the scaled scores per query row
  float scale_log2;   // (1/sqrt(d)) * log2(e)
  float scale;        // 1/sqrt(d)
  int dtype;
  uint32_t idesc_qk;  // M=128, N=128, both K-major
  uint32_t idesc_pv;  // M=128, N=d, A K-major, B MN-major
  int debug;          // experiments only (CSN_ATTN_DEBUG): bit 0 = skip the softmax arithmetic, bit 1 = skip the epilogue stores
  // dropout on the probabilities: off when drop_thresh == 0.  Row id of a query = its index into lse
  // (lse_off + row: unique per block, head and row), column = key index inside the chunk.
  uint32_t drop_seed, drop_thresh;
  float drop_scale;
  const uint32_t* drop_epoch;   // device word added (x 0x9E3779B9) to drop_seed: graph replays draw fresh masks
};

// attn_wide.cu: d_head = 256 with [128 x 256] score tiles (mode 0 forward, mode 1 dV)
int launch_attn_wide(int mode, bool pair, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                     const CUtensorMap& tmO, const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream);

// attn_fwd64.cu: d_head = 64 forward (16 softmax warps, private accumulators per column quarter, P through TMEM)
int launch_attn_fwd64(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                      const CUtensorMap& tmOlo, const AttnFwdArgs& a, cudaStream_t stream);

}  // namespace csn
