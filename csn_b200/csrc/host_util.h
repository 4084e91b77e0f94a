// Host-side helpers shared by the C-ABI entry points: error string, launch counter, TMA tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/csn_b200.h"

namespace csn {

void set_error(const char* fmt, ...);
void clear_error();
std::atomic<int64_t>& launch_counter();
int num_sms();
// device word (per device, zero until csn_set_drop_epoch) that every dropout mask's seed is offset by: see ptx.cuh
const uint32_t* drop_epoch_ptr();

// 2-D tiled tensor map over 16-bit elements with 128-byte swizzle.  dims/box are {inner, outer}.
// Out-of-bounds elements of a box are filled with zeros.  Returns 0 on success.
int make_tmap_2d(CUtensorMap* tm, const void* ptr, int dtype, int64_t inner, int64_t outer,
                 int64_t ld_elems, uint32_t box_inner, uint32_t box_outer);
// Same for any element type (CSN_F32 / CSN_F16 / CSN_BF16); box_inner * element size must be 128 B.
int make_tmap_2d_any(CUtensorMap* tm, const void* ptr, int dtype, int64_t inner, int64_t outer,
                     int64_t ld_elems, uint32_t box_inner, uint32_t box_outer);

#define CSN_CHECK_ARG(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      csn::set_error(__VA_ARGS__); \
      return 1;                    \
    }                              \
  } while (0)

#define CSN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      csn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,     \
                     __LINE__);                                                             \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

// cudaFuncSetAttribute applies to the CURRENT device: remember per (call site, device), not per process
#define CSN_SET_MAX_SMEM(kern, bytes)                                                                      \
  do {                                                                                                     \
    static bool cfg__[64] = {false};                                                                       \
    int dev__ = 0;                                                                                         \
    if (cudaGetDevice(&dev__) != cudaSuccess) dev__ = -1;                                                  \
    if (dev__ < 0 || dev__ >= 64 || !cfg__[dev__]) {                                                       \
      CSN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      if (dev__ >= 0 && dev__ < 64) cfg__[dev__] = true;                                                   \
    }                                                                                                      \
  } while (0)

// Call after every kernel launch.
#define CSN_LAUNCH_OK(name)                                                             \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      csn::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));         \
      return 3;                                                                         \
    }                                                                                   \
    csn::launch_counter().fetch_add(1, std::memory_order_relaxed);                      \
  } while (0)

}  // namespace csn
