// Fused segmentation epilogue: the 1x1 `logit` conv 256 -> C (csa_models.py:201, bias-free), the
// label-0 mask and the cross-entropy of the training scripts (csa_training.py:94-108: mean over the points whose
// label is > 0), forward AND backward in one pass over the (B, 256, N) activation:
//   logits[c] = sum_k W[c][k] f[k][n];  p = softmax(logits);  loss += -log p[label]           (valid points)
//   dlogits[c] = (p[c] - [c == label]) / n_valid   (0 for masked points);  dfeat[k][n] = sum_c W[c][k] dlogits[c]
// One thread per point: the channel-major activation makes every load and store of a warp one contiguous 128 bytes;
// W sits in SMEM and is read as broadcasts.  HBM traffic = read f once + write dfeat once (164 MB at B=8) instead of
// the five ATen kernels (conv fwd, log-softmax, nll, their backwards) that re-read logits and features.
// The weight gradient dW = dlogits f^T stays a library GEMM on the (B, C, N) dlogits this kernel also writes.
#include <stdint.h>

#include "host_util.h"

namespace csn {

__global__ void count_valid_kernel(const long long* __restrict__ labels, long long n, int ignore_index, int* __restrict__ count) {
  int c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += labels[i] != ignore_index;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

struct SegLossArgs {
  const float* feat; long long b_stride, ch_stride;
  int n_points, C;
  const float* W;            // [C][256]
  const long long* labels;   // [B][n_points]
  int ignore_index;
  const int* n_valid;        // device scalar written by count_valid_kernel
  const float* grad_scale;   // optional device scalar multiplied into dlogits / dfeat (the upstream gradient of the loss)
  float* loss_part;          // [gridDim.y][gridDim.x] per-CTA partial sums of -log p[label]
  float* dlogits;            // [B][C][n_points]
  float* dfeat;              // [B][256][n_points] (same strides as feat)
};

template <int CMAX>
__global__ void __launch_bounds__(128) seg_loss_kernel(const SegLossArgs p) {
  extern __shared__ float Wsm[];   // [C][256]
  __shared__ float red[4];
  for (int i = threadIdx.x; i < p.C * 256; i += 128) Wsm[i] = p.W[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int n = blockIdx.x * 128 + threadIdx.x;
  const bool in = n < p.n_points;
  const float* f = p.feat + (long long)b * p.b_stride + (in ? n : 0);
  float acc[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
#pragma unroll 4   // 16 independent loads in flight per thread
  for (int k = 0; k < 256; k += 4) {
    const float f0 = in ? __ldg(f + (long long)(k + 0) * p.ch_stride) : 0.f;
    const float f1 = in ? __ldg(f + (long long)(k + 1) * p.ch_stride) : 0.f;
    const float f2 = in ? __ldg(f + (long long)(k + 2) * p.ch_stride) : 0.f;
    const float f3 = in ? __ldg(f + (long long)(k + 3) * p.ch_stride) : 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < p.C) {
        const float4 w = *reinterpret_cast<const float4*>(Wsm + c * 256 + k);
        acc[c] += w.x * f0 + w.y * f1 + w.z * f2 + w.w * f3;
      }
  }
  // softmax / loss / dlogits
  const long long lab = in ? p.labels[(long long)b * p.n_points + n] : (long long)p.ignore_index;
  const bool valid = in && lab != p.ignore_index && lab >= 0 && lab < p.C;   // out-of-range labels are masked (csn_csa_head counts them)
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) if (c < p.C) mx = fmaxf(mx, acc[c]);
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) if (c < p.C) { acc[c] = __expf(acc[c] - mx); se += acc[c]; }
  const float inv = 1.f / se;
  const float scale = valid ? (p.grad_scale ? __ldg(p.grad_scale) : 1.f) / (float)max(*p.n_valid, 1) : 0.f;
  float lterm = 0.f;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < p.C) {
      const float pc = acc[c] * inv;
      if (valid && c == (int)lab) lterm = -__logf(fmaxf(pc, 1e-38f));
      acc[c] = (pc - ((valid && c == (int)lab) ? 1.f : 0.f)) * scale;
      if (in && p.dlogits) p.dlogits[((long long)b * p.C + c) * p.n_points + n] = acc[c];
    }
  // per-CTA partial loss (fixed summation order inside the CTA; the caller sums the partials)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lterm += __shfl_xor_sync(0xffffffffu, lterm, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lterm;
  __syncthreads();
  if (threadIdx.x == 0) p.loss_part[(long long)blockIdx.y * gridDim.x + blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
  // dfeat
  if (p.dfeat == nullptr || !in) return;
  float* df = p.dfeat + (long long)b * p.b_stride + n;
#pragma unroll 2
  for (int k = 0; k < 256; k += 4) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < p.C) {
        const float4 w = *reinterpret_cast<const float4*>(Wsm + c * 256 + k);
        s0 += w.x * acc[c]; s1 += w.y * acc[c]; s2 += w.z * acc[c]; s3 += w.w * acc[c];
      }
    df[(long long)(k + 0) * p.ch_stride] = s0;
    df[(long long)(k + 1) * p.ch_stride] = s1;
    df[(long long)(k + 2) * p.ch_stride] = s2;
    df[(long long)(k + 3) * p.ch_stride] = s3;
  }
}

template <int CMAX>
static int launch_seg(const SegLossArgs& a, int n_batch, cudaStream_t s) {
  auto kern = seg_loss_kernel<CMAX>;
  const int smem = a.C * 256 * 4;
  int dev = 0;
  CSN_CUDA_OK(cudaGetDevice(&dev));
  static bool configured[64] = {false};   // the attribute is per device
  if (dev < 64 && !configured[dev]) {
    CSN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CMAX * 256 * 4));
    configured[dev] = true;
  }
  kern<<<dim3((a.n_points + 127) / 128, n_batch), 128, smem, s>>>(a);
  CSN_LAUNCH_OK("seg_loss_kernel");
  return 0;
}

}  // namespace csn

// loss_part must hold n_batch * ceil(n_points/128) floats; n_valid is a zero-initialised device int that receives
// the number of unmasked points (the caller divides the summed partials by it).  dlogits / dfeat may be NULL;
// grad_scale (optional device scalar) is multiplied into both (the upstream gradient of the loss).
extern "C" int csn_seg_loss(const float* feat, int64_t b_stride, int64_t ch_stride, int32_t n_batch, int32_t n_points,
                            const float* W, int32_t n_classes, const int64_t* labels, int32_t ignore_index,
                            int32_t* n_valid, float* loss_part, float* dlogits, float* dfeat, const float* grad_scale,
                            void* stream) {
  using namespace csn;
  clear_error();
  CSN_CHECK_ARG(feat && W && labels && n_valid && loss_part, "csn_seg_loss: null pointer");
  CSN_CHECK_ARG(n_classes >= 1 && n_classes <= 64, "csn_seg_loss: 1..64 classes supported (got %d)", n_classes);
  if (n_batch == 0 || n_points == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  count_valid_kernel<<<148, 256, 0, s>>>(reinterpret_cast<const long long*>(labels), (long long)n_batch * n_points, ignore_index, n_valid);
  CSN_LAUNCH_OK("count_valid_kernel");
  SegLossArgs a{feat, b_stride, ch_stride, n_points, n_classes, W, reinterpret_cast<const long long*>(labels), ignore_index,
                n_valid, grad_scale, loss_part, dlogits, dfeat};
  if (n_classes <= 16) return launch_seg<16>(a, n_batch, s);
  if (n_classes <= 32) return launch_seg<32>(a, n_batch, s);
  return launch_seg<64>(a, n_batch, s);
}
