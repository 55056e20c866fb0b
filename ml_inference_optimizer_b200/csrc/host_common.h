// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor encoding
// through the driver entry point (no link-time dependency on libcuda), device properties.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>

#include "../../include/b200_attn_mlp.h"

namespace b200 {

// thread-local last-error string returned by b200_last_error()
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)                                \
  do {                                                           \
    if (!(cond)) return ::b200::set_error(B200_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

#define B200_CUDA_OK(expr)                                                                           \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::b200::set_error(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                  \
  } while (0)

// Encode a tiled tensor map with 128-byte swizzle for a 16-bit element tensor.
//   rank      : 2..5
//   dims      : extent of each dimension, innermost first (elements)
//   strides_b : byte stride of dimensions 1..rank-1 (dimension 0 is contiguous)
//   box       : box extent per dimension (elements); box[0]*2 bytes must be <= 128
// Returns 0 or a negative B200 error code.
int encode_tmap_sw128_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                          const uint32_t* box);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, so a process
// that drives several GPUs must set it on each of them. `slot` is a per-kernel static array of 64 flags.
cudaError_t set_max_dynamic_smem(const void* func, int bytes, bool (&slot)[64]);

// launch accounting (b200_launch_count / b200_last_gemm_kernel): every kernel launch of the library calls note_launch
void note_launch(const char* kernel_name, bool is_gemm = false);

int gemm_group_rows();  // b200_set_gemm_group_rows value: rows of the activation panel one L2 raster group covers

int sm_count();  // multiprocessor count of the current device (cached per device)
int sm_limit();  // b200_set_sm_limit value (0 = no limit): persistent kernels launch at most this many CTAs

}  // namespace b200
