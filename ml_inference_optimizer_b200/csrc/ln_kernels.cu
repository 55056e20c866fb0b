// LayerNorm and LayerNorm(x + alpha * residual) for sm_100a — the op immediately before the attention / MLP blocks
// (SURVEY.md §8 f3). Replaces _layernorm_fwd_kernel / _layernorm_residual_fwd_kernel / triton_layernorm
// (kernels/triton/layernorm_kernels.py:36-190, :191-277).
//
// HBM-bound: each element is read once (x, and the residual when present) and written once. One warp (<= 2048 columns)
// or one CTA of 256 threads (<= 8192) per row, both persistent with the next row's 128-bit loads in flight; the row is
// held in fp32 registers between the mean pass and the variance pass (two-pass variance like the reference's
// (x - u)^2 mean, not E[x^2] - u^2).
// Algorithmic bytes per row: cols * 2 * (2 + has_residual).

#include <stdlib.h>

#include "common.cuh"
#include "host_common.h"

namespace b200 {
namespace ln {

constexpr int THREADS = 256;
constexpr int VEC = 8;         // 16-bit elements per 128-bit access
constexpr int MAX_ITERS = 4;   // cols <= THREADS * VEC * MAX_ITERS = 8192

// streaming 128-bit load: the activations are read exactly once, so they bypass L1 allocation
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Block-wide sum; `red` is one of two alternating 8-float buffers, so a single __syncthreads per reduction is enough:
// buffer A is only rewritten after every thread has passed the barrier of the following reduction on buffer B.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int x = 16; x >= 1; x >>= 1) v += __shfl_xor_sync(0xffffffffu, v, x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[lane & (THREADS / 32 - 1)];
#pragma unroll
  for (int x = 4; x >= 1; x >>= 1) t += __shfl_xor_sync(0xffffffffu, t, x);
  return t;
}

// wide rows (2048 < cols <= 8192): one CTA per row at a time, persistent over the rows (grid = resident CTAs), ITERS =
// ceil(cols / 2048) 128-bit loads per thread and operand. The loads of the CTA's NEXT row are issued before the two
// block reductions of the current one, so every CTA keeps a full row in flight while it reduces, normalises and
// stores (a CTA that exits per row has nothing in flight during those phases: 0.60-0.68 of copy bandwidth at 4096
// columns). weight / bias are re-read from L1/L2 per row (8-16 KB, shared by every row).
template <typename T, bool HAS_RES, int ITERS>
__global__ void __launch_bounds__(THREADS)
layernorm_kernel(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ w, const T* __restrict__ b,
                 T* __restrict__ y, int64_t rows, int cols, int64_t ldx, int64_t ldr, int64_t ldy, float eps, float alpha) {
  __shared__ float red[2][THREADS / 32];
  uint4 nx[ITERS], nr[ITERS];
  auto load_row = [&](int64_t r) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * THREADS + threadIdx.x) * VEC;
      if (c < cols) {
        nx[it] = ld_stream(x + r * ldx + c);
        if constexpr (HAS_RES) nr[it] = ld_stream(res + r * ldr + c);
      }
    }
  };
  int64_t row = blockIdx.x;
  if (row < rows) load_row(row);
  for (; row < rows; row += gridDim.x) {
    float v[ITERS][VEC];
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * THREADS + threadIdx.x) * VEC;
      if (c < cols) {
        const uint32_t rw[4] = {nx[it].x, nx[it].y, nx[it].z, nx[it].w};
        uint32_t rs[4] = {0, 0, 0, 0};
        if constexpr (HAS_RES) { rs[0] = nr[it].x; rs[1] = nr[it].y; rs[2] = nr[it].z; rs[3] = nr[it].w; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = Pack2<T>::unpack(rw[q]);
          v[it][2 * q] = f.x;
          v[it][2 * q + 1] = f.y;
          if constexpr (HAS_RES) {
            const float2 g = Pack2<T>::unpack(rs[q]);
            v[it][2 * q] = fmaf(alpha, g.x, v[it][2 * q]);
            v[it][2 * q + 1] = fmaf(alpha, g.y, v[it][2 * q + 1]);
          }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) sum += v[it][e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[it][e] = 0.f;
      }
    }
    if (row + gridDim.x < rows) load_row(row + gridDim.x);
    const float mean = block_sum(sum, red[0]) / static_cast<float>(cols);
    float sq = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      if ((it * THREADS + threadIdx.x) * VEC < cols) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float d = v[it][e] - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
    const float rstd = rsqrtf(block_sum(sq, red[1]) / static_cast<float>(cols) + eps);
    T* yr = y + row * ldy;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * THREADS + threadIdx.x) * VEC;
      if (c < cols) {
        const uint4 wr = __ldg(reinterpret_cast<const uint4*>(w + c));
        uint4 br = make_uint4(0, 0, 0, 0);
        if (b != nullptr) br = __ldg(reinterpret_cast<const uint4*>(b + c));
        const uint32_t ww[4] = {wr.x, wr.y, wr.z, wr.w}, bb[4] = {br.x, br.y, br.z, br.w};
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 wf = Pack2<T>::unpack(ww[q]), bf = Pack2<T>::unpack(bb[q]);
          pk[q] = Pack2<T>::pack(fmaf((v[it][2 * q] - mean) * rstd, wf.x, bf.x),
                                 fmaf((v[it][2 * q + 1] - mean) * rstd, wf.y, bf.y));
        }
        *reinterpret_cast<uint4*>(yr + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

// narrow rows (cols <= 2048): one warp per row, 8 rows per CTA, no shared memory, shuffles only
constexpr int W_MAX_ITERS_LIMIT = 8;  // warp-per-row kernel: cols <= 32 * VEC * 8 = 2048

// ITERS = ceil(cols / 256) is a template parameter: the row lives in ITERS*8 registers per lane, so narrow rows (768
// columns: 24 registers) leave room for 5+ resident CTAs per SM instead of the 2 a worst-case 2048-column row buffer
// allows — the kernel is HBM-bound and needs the bytes in flight. The grid is persistent (warps stride over the rows).
template <typename T, bool HAS_RES, int ITERS, bool PREFETCH>
__global__ void __launch_bounds__(THREADS)
layernorm_warp_kernel(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ w, const T* __restrict__ b,
                      T* __restrict__ y, int64_t rows, int cols, int64_t ldx, int64_t ldr, int64_t ldy, float eps,
                      float alpha) {
  constexpr int W_MAX_ITERS = ITERS;
  // Software pipelining across rows: the 128-bit loads of the warp's NEXT row are issued as soon as the current row has
  // been unpacked to fp32, so they are in flight during the two shuffle reductions and the stores (narrow rows only: the
  // staging registers cost 4 * ITERS (8 with a residual) per lane).
  constexpr bool kPrefetch = PREFETCH && ITERS <= 4;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (THREADS / 32);
  uint4 nx[W_MAX_ITERS], nr[W_MAX_ITERS];
  auto load_row = [&](int64_t r) {
#pragma unroll
    for (int it = 0; it < W_MAX_ITERS; ++it) {
      const int c = (it * 32 + lane) * VEC;
      if (c < cols) {
        nx[it] = *reinterpret_cast<const uint4*>(x + r * ldx + c);
        if constexpr (HAS_RES) nr[it] = *reinterpret_cast<const uint4*>(res + r * ldr + c);
      }
    }
  };
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * (THREADS / 32) + (threadIdx.x >> 5);
  if (kPrefetch && row0 < rows) load_row(row0);
  for (int64_t row = row0; row < rows; row += warps_total) {
  if (!kPrefetch) load_row(row);
  float v[W_MAX_ITERS][VEC];
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < W_MAX_ITERS; ++it) {
    const int c = (it * 32 + lane) * VEC;
    if (c < cols) {
      const uint4 raw = nx[it];
      const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t rs[4] = {0, 0, 0, 0};
      if constexpr (HAS_RES) {
        const uint4 r4 = nr[it];
        rs[0] = r4.x; rs[1] = r4.y; rs[2] = r4.z; rs[3] = r4.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = Pack2<T>::unpack(rw[q]);
        v[it][2 * q] = f.x;
        v[it][2 * q + 1] = f.y;
        if constexpr (HAS_RES) {
          const float2 g = Pack2<T>::unpack(rs[q]);
          v[it][2 * q] = fmaf(alpha, g.x, v[it][2 * q]);
          v[it][2 * q + 1] = fmaf(alpha, g.y, v[it][2 * q + 1]);
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) sum += v[it][e];
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[it][e] = 0.f;
    }
  }
  if (kPrefetch && row + warps_total < rows) load_row(row + warps_total);
#pragma unroll
  for (int xo = 16; xo >= 1; xo >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, xo);
  const float mean = sum / static_cast<float>(cols);
  float sq = 0.f;
#pragma unroll
  for (int it = 0; it < W_MAX_ITERS; ++it) {
    if ((it * 32 + lane) * VEC < cols) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float d = v[it][e] - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
#pragma unroll
  for (int xo = 16; xo >= 1; xo >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, xo);
  const float rstd = rsqrtf(sq / static_cast<float>(cols) + eps);
  T* yr = y + row * ldy;
#pragma unroll
  for (int it = 0; it < W_MAX_ITERS; ++it) {
    const int c = (it * 32 + lane) * VEC;
    if (c < cols) {
      const uint4 wr = __ldg(reinterpret_cast<const uint4*>(w + c));
      uint4 br = make_uint4(0, 0, 0, 0);
      if (b != nullptr) br = __ldg(reinterpret_cast<const uint4*>(b + c));
      const uint32_t ww[4] = {wr.x, wr.y, wr.z, wr.w}, bb[4] = {br.x, br.y, br.z, br.w};
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 wf = Pack2<T>::unpack(ww[q]), bf = Pack2<T>::unpack(bb[q]);
        pk[q] = Pack2<T>::pack(fmaf((v[it][2 * q] - mean) * rstd, wf.x, bf.x),
                               fmaf((v[it][2 * q + 1] - mean) * rstd, wf.y, bf.y));
      }
      *reinterpret_cast<uint4*>(yr + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  }  // row loop
}

}  // namespace ln
}  // namespace b200

extern "C" int b200_layernorm(const void* x, const void* residual, const void* weight, const void* bias, void* y,
                              int64_t rows, int cols, int64_t ldx, int64_t ldr, int64_t ldy, float eps,
                              float residual_alpha, int dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(x && weight && y, "layernorm: NULL pointer argument");
  B200_CHECK_ARG(rows >= 0 && rows <= 0x7fffffffLL, "layernorm: bad row count");
  B200_CHECK_ARG(cols > 0 && cols % 8 == 0 && cols <= ln::THREADS * ln::VEC * ln::MAX_ITERS,
                 "layernorm: cols must be a multiple of 8 and <= 8192 (got %d)", cols);
  B200_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= cols && ldy >= cols && (!residual || (ldr % 8 == 0 && ldr >= cols)),
                 "layernorm: row strides must be multiples of 8 elements and >= cols");
  B200_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)weight & 15) == 0 &&
                     ((uintptr_t)bias & 15) == 0 && ((uintptr_t)residual & 15) == 0,
                 "layernorm: pointers must be 16-byte aligned");
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_FP16, "layernorm: dtype must be bf16 or fp16");
  if (rows == 0) return B200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // B200_LN_NARROW_MAX (developer knob): widest row the warp-per-row kernel takes
  const char* nm_env = getenv("B200_LN_NARROW_MAX");
  // measured (T = 32768 rows, profiles/r2_ln_probe2.log): up to 1024 columns the warp kernel (its next-row prefetch needs
  // ITERS <= 4) wins; at 2048 columns the CTA kernel wins with a residual (5172 vs 4012 GB/s) and loses without one
  // (3857 vs 4443 GB/s)
  const int narrow_dflt = residual ? 1024 : 32 * ln::VEC * ln::W_MAX_ITERS_LIMIT;
  const int narrow_max = (nm_env && atoi(nm_env) > 0) ? min(atoi(nm_env), 32 * ln::VEC * ln::W_MAX_ITERS_LIMIT) : narrow_dflt;
  const bool narrow = cols <= narrow_max;
  const int w_iters = (cols + 32 * ln::VEC - 1) / (32 * ln::VEC);
  const int64_t warp_ctas = (rows + ln::THREADS / 32 - 1) / (ln::THREADS / 32);
  const int64_t resident = static_cast<int64_t>(sm_count()) * 8;  // persistent: at most 8 CTAs of 256 threads per SM
  const char* pf_env = getenv("B200_LN_PREFETCH");
  const bool prefetch = !(pf_env && pf_env[0] == '0');
  const int wide_iters = (cols + ln::THREADS * ln::VEC - 1) / (ln::THREADS * ln::VEC);  // 2..4 for 2048 < cols <= 8192
  // wide rows: persistent as well; B200_LN_WIDE_CTAS_PER_SM (developer knob) overrides the resident CTAs per SM
  const char* wc_env = getenv("B200_LN_WIDE_CTAS_PER_SM");
  const int wide_per_sm = (wc_env && atoi(wc_env) > 0) ? atoi(wc_env) : 0;  // 0: as many as fit
  const unsigned grid = static_cast<unsigned>(warp_ctas < resident ? warp_ctas : resident);  // (narrow kernel)
#define LAUNCH_W(T, HAS, IT)                                                                                       \
  case IT:                                                                                                             \
    if (prefetch)                                                                                                    \
      ln::layernorm_warp_kernel<T, HAS, IT, true><<<grid, ln::THREADS, 0, s>>>(                                        \
          static_cast<const T*>(x), static_cast<const T*>(residual), static_cast<const T*>(weight),                    \
          static_cast<const T*>(bias), static_cast<T*>(y), rows, cols, ldx, ldr, ldy, eps, residual_alpha);            \
    else                                                                                                             \
      ln::layernorm_warp_kernel<T, HAS, IT, false><<<grid, ln::THREADS, 0, s>>>(                                       \
          static_cast<const T*>(x), static_cast<const T*>(residual), static_cast<const T*>(weight),                    \
          static_cast<const T*>(bias), static_cast<T*>(y), rows, cols, ldx, ldr, ldy, eps, residual_alpha);            \
    break;
#define LAUNCH_C(T, HAS, IT)                                                                                       \
  case IT: {                                                                                                           \
    static int occ = 0; /* resident CTAs per SM of this instantiation (register-limited: 2 at 8192 columns + residual) */ \
    if (occ == 0) {                                                                                                    \
      B200_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln::layernorm_kernel<T, HAS, IT>, ln::THREADS, 0)); \
      if (occ < 1) occ = 1;                                                                                            \
    }                                                                                                                  \
    const int64_t res_w = static_cast<int64_t>(sm_count()) * (wide_per_sm > 0 && wide_per_sm < occ ? wide_per_sm : occ); \
    ln::layernorm_kernel<T, HAS, IT><<<static_cast<unsigned>(rows < res_w ? rows : res_w), ln::THREADS, 0, s>>>(       \
        static_cast<const T*>(x), static_cast<const T*>(residual), static_cast<const T*>(weight),                      \
        static_cast<const T*>(bias), static_cast<T*>(y), rows, cols, ldx, ldr, ldy, eps, residual_alpha);              \
  } break;
#define LAUNCH(T, HAS)                                                                                              \
  if (narrow) {                                                                                                        \
    switch (w_iters) {                                                                                                 \
      LAUNCH_W(T, HAS, 1) LAUNCH_W(T, HAS, 2) LAUNCH_W(T, HAS, 3) LAUNCH_W(T, HAS, 4) LAUNCH_W(T, HAS, 5)              \
      LAUNCH_W(T, HAS, 6) LAUNCH_W(T, HAS, 7) LAUNCH_W(T, HAS, 8)                                                      \
      default: break;                                                                                                  \
    }                                                                                                                  \
  } else {                                                                                                               \
    switch (wide_iters) {                                                                                              \
      LAUNCH_C(T, HAS, 1) LAUNCH_C(T, HAS, 2) LAUNCH_C(T, HAS, 3) LAUNCH_C(T, HAS, 4)                                                   \
      default: break;                                                                                                  \
    }                                                                                                                  \
  }
  if (dtype == B200_DTYPE_BF16) {
    if (residual) { LAUNCH(__nv_bfloat16, true); } else { LAUNCH(__nv_bfloat16, false); }
  } else {
    if (residual) { LAUNCH(__half, true); } else { LAUNCH(__half, false); }
  }
#undef LAUNCH
#undef LAUNCH_W
#undef LAUNCH_C
  B200_CUDA_OK(cudaGetLastError());
  note_launch("layernorm_kernel");
  return B200_OK;
}
