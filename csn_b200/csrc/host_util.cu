#include "host_util.h"

#include <stdarg.h>

namespace csn {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void clear_error() { g_err[0] = 0; }

std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> c{0};
  return c;
}

int num_sms() {   // of the current device (cached per device)
  static int n[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (n[dev] == 0 && cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n[dev] = 148;
  return n[dev];
}

static uint32_t* g_drop_epoch[64] = {nullptr};

static uint32_t* drop_epoch_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_drop_epoch[dev]) {
    uint32_t* p = nullptr;
    if (cudaMalloc(&p, sizeof(uint32_t)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(uint32_t));
    g_drop_epoch[dev] = p;
  }
  return g_drop_epoch[dev];
}

const uint32_t* drop_epoch_ptr() { return drop_epoch_slot(); }

__global__ void set_epoch_kernel(uint32_t* p, uint32_t v) { *p = v; }

int set_drop_epoch(uint32_t epoch, void* stream) {
  uint32_t* p = drop_epoch_slot();
  if (!p) { set_error("csn_set_drop_epoch: could not allocate the epoch word"); return 2; }
  set_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p, epoch);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked (it does not exist on the build box); the entry point is resolved through
// the runtime the first time a tensor map is needed.
static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* tm, const void* ptr, int dtype, int64_t inner, int64_t outer,
                 int64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  encode_tiled_fn enc = get_encode();
  CSN_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  CSN_CHECK_ARG(dtype == CSN_F16 || dtype == CSN_BF16, "tensor map: 16-bit element types only");
  CSN_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: pointer not 16B aligned");
  CSN_CHECK_ARG((ld_elems * 2) % 16 == 0, "tensor map: row stride %lld elems not 16B multiple",
                (long long)ld_elems);
  CSN_CHECK_ARG(inner > 0 && outer > 0, "tensor map: empty extent");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dtype == CSN_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSN_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld outer=%lld ld=%lld box=%ux%u)",
                (int)r, (long long)inner, (long long)outer, (long long)ld_elems, box_inner, box_outer);
  return 0;
}

int make_tmap_2d_any(CUtensorMap* tm, const void* ptr, int dtype, int64_t inner, int64_t outer,
                     int64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  if (dtype != CSN_F32) return make_tmap_2d(tm, ptr, dtype, inner, outer, ld_elems, box_inner, box_outer);
  encode_tiled_fn enc = get_encode();
  CSN_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  CSN_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor map: pointer not 16B aligned");
  CSN_CHECK_ARG((ld_elems * 4) % 16 == 0, "tensor map: row stride not a 16B multiple");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSN_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace csn

extern "C" {
const char* csn_last_error(void) { return csn::g_err; }
int csn_abi_version(void) { return 2; }
int csn_set_drop_epoch(uint32_t epoch, void* stream) {
  csn::clear_error();
  return csn::set_drop_epoch(epoch, stream);
}
int64_t csn_launch_count(void) { return csn::launch_counter().load(); }
}
