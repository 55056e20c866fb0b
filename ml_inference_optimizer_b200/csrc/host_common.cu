// Host-side plumbing of the C-ABI: error strings, TMA descriptor encoding, device queries.
#include "host_common.h"

#include <cudaTypedefs.h>

#include <atomic>

namespace b200 {

std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int encode_tmap_sw128_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                          const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  if (rank < 2 || rank > 5) return set_error(B200_ERR_INVALID_ARGUMENT, "tensor map rank %d out of range", rank);
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (dims[i] == 0) return set_error(B200_ERR_INVALID_ARGUMENT, "tensor map dimension %d is empty", i);
    if (box[i] == 0 || box[i] > 256) return set_error(B200_ERR_INVALID_ARGUMENT, "tensor map box %d = %u", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstrides[i] = strides_b[i];
    if (strides_b[i] % 16 != 0)
      return set_error(B200_ERR_INVALID_ARGUMENT, "tensor stride %d (%llu bytes) must be a multiple of 16 bytes", i + 1,
                       (unsigned long long)strides_b[i]);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(B200_ERR_INVALID_ARGUMENT, "tensor base address must be 16-byte aligned");
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims,
                  gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return B200_OK;
}

std::atomic<int> g_sm_limit{0};
std::atomic<int> g_group_rows{0};  // 0 = choose from the L2 budget
std::atomic<long long> g_launches{0};
std::atomic<const char*> g_last_gemm{""};
std::atomic<const char*> g_last_kernel{""};

cudaError_t set_max_dynamic_smem(const void* func, int bytes, bool (&slot)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && slot[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) slot[dev] = true;  // benign race: setting it twice is harmless
  return e;
}

void note_launch(const char* kernel_name, bool is_gemm) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_last_kernel.store(kernel_name, std::memory_order_relaxed);
  if (is_gemm) g_last_gemm.store(kernel_name, std::memory_order_relaxed);
}

int gemm_group_rows() { return g_group_rows.load(std::memory_order_relaxed); }

int sm_limit() { return g_sm_limit.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200

extern "C" {

const char* b200_version(void) { return "b200_attn_mlp 0.1.0 (sm_100a)"; }

const char* b200_last_error(void) { return b200::last_error_ref().c_str(); }

int b200_set_sm_limit(int max_ctas) {
  if (max_ctas < 0) return b200::set_error(B200_ERR_INVALID_ARGUMENT, "sm limit must be >= 0");
  b200::g_sm_limit.store(max_ctas, std::memory_order_relaxed);
  return B200_OK;
}

int b200_set_gemm_group_rows(int rows) {
  if (rows != 0 && rows < 256) return b200::set_error(B200_ERR_INVALID_ARGUMENT, "raster group must cover >= 256 rows (0 = automatic)");
  b200::g_group_rows.store(rows, std::memory_order_relaxed);
  return B200_OK;
}

int64_t b200_launch_count(void) { return b200::g_launches.load(std::memory_order_relaxed); }

const char* b200_last_gemm_kernel(void) { return b200::g_last_gemm.load(std::memory_order_relaxed); }

const char* b200_last_kernel(void) { return b200::g_last_kernel.load(std::memory_order_relaxed); }

int b200_arch_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return b200::set_error(B200_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return b200::set_error(B200_ERR_NO_DEVICE, "cannot query device: %s", cudaGetErrorString(e));
  }
  return major == 10 ? 1 : 0;
}

}  // extern "C"
