// K3 — FusedMLP for sm_100a: persistent, warp-specialised tcgen05 GEMM with the bias + activation (GELU-tanh,
// GELU-erf, ReLU, SwiGLU) fused into the TMEM->register epilogue.
//
// Replaces the reference's Triton kernels _fused_mlp_{gelu,relu,swiglu}_kernel
// (kernels/triton/mlp_kernels.py:27-641) and the eager path FusedMLP._forward_pytorch
// (kernels/mlp/fused_mlp.py:159-178, :223-237, :262-275).
//
// One CTA per SM, 256 threads:
//   warp 0     TMA producer        (A tile 128x64, B tile 256x64 per stage, 128-byte swizzle, 4 stages)
//   warp 1     MMA issuer          (tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, fp32 accum in TMEM)
//   warp 2     TMEM allocator      (512 columns = two 256-column accumulator stages)
//   warps 4-7  epilogue            (tcgen05.ld -> bias/activation -> bf16 -> swizzled smem -> TMA store)
// The accumulator is double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// SwiGLU: the 256 accumulator columns of a tile are [128 gate | 128 up] — the B stage is filled by two TMA
// loads (rows of W_gate, rows of W_up) and one N=256 MMA computes both; the epilogue reads matching gate/up
// columns, applies silu(g)*u and emits a 128x128 output tile. The [T, i] gate and up tensors never exist.

#include <stdlib.h>

#include "common.cuh"
#include "host_common.h"

namespace b200 {

namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;  // accumulator columns per tile
constexpr int BK = 64;   // 64 x 16-bit = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;           // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;           // 32 KB
constexpr int C_BUF_BYTES = BM * 64 * 2;             // 16 KB : 128 rows x 64 cols staging
constexpr int SMEM_A_OFF = 0;
constexpr int SMEM_B_OFF = SMEM_A_OFF + STAGES * A_STAGE_BYTES;
constexpr int SMEM_C_OFF = SMEM_B_OFF + STAGES * B_STAGE_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_C_OFF + 2 * C_BUF_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR_OFF + 256 + 1024;  // barriers + alignment slack
constexpr int NUM_THREADS = 256;
constexpr int GROUP_M = 16;  // rasterisation: 16 row-blocks share a column-block sweep (L2 reuse)

struct Params {
  int M, N_out, K;       // N_out: output columns (SwiGLU: number of gate/up pairs)
  const void* bias0;     // [N_out] bias (SwiGLU: gate bias) or nullptr
  const void* bias1;     // SwiGLU: up bias or nullptr
  int num_m_blocks, num_n_blocks, num_k_blocks, num_tiles;
  // split-K (skinny problems: few output tiles, long K): tile index = mn_tile * k_splits + split; each split accumulates
  // k-blocks [split*kb_per_split, ...) and stores raw fp32 accumulators to `partial` [k_splits][M][num_n_blocks*256]
  int k_splits, kb_per_split;
  float* partial;
  void* y;       // output pointer / row stride for the split-K reduce pass
  int64_t ldy;
  int use_pair;  // CTA-pair kernel (M = 256 tiles): num_m_blocks counts 256-row blocks
};

__device__ __forceinline__ void tile_coords(const Params& p, int tile_in, int& m_blk, int& n_blk, int& kb0, int& kb1) {
  const int split = tile_in % p.k_splits;
  const int tile = tile_in / p.k_splits;
  kb0 = split * p.kb_per_split;
  kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
  const int per_group = GROUP_M * p.num_n_blocks;
  const int group = tile / per_group;
  const int first_m = group * GROUP_M;
  const int rows_in_group = min(GROUP_M, p.num_m_blocks - first_m);
  const int in_group = tile - group * per_group;
  m_blk = first_m + in_group % rows_in_group;
  n_blk = in_group / rows_in_group;
}

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if constexpr (ACT == B200_ACT_GELU_TANH) {
    // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))   — mlp_kernels.py:144-161
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float inner = k0 * x * fmaf(k1 * x, x, 1.0f);
    return 0.5f * x * (1.0f + fast_tanh(inner));
  } else if constexpr (ACT == B200_ACT_GELU_ERF) {
    return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
  } else if constexpr (ACT == B200_ACT_RELU) {
    return fmaxf(x, 0.0f);
  } else {
    return x;
  }
}

__device__ __forceinline__ float silu(float g) {
  // g * sigmoid(g) = g / (1 + exp(-g))
  return g * fast_rcp(1.0f + fast_exp2(-1.4426950408889634f * g));
}

// load 32 bias values (columns col..col+31) as fp32; all lanes read the same addresses (L1 broadcast)
template <typename T>
__device__ __forceinline__ void load_bias32(const void* bias, int col, int n_limit, float (&out)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (bias != nullptr && col + g * 8 + 8 <= n_limit) {
      v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(bias) + col + g * 8));
    }
    float2 f;
    f = Pack2<T>::unpack(v.x); out[g * 8 + 0] = f.x; out[g * 8 + 1] = f.y;
    f = Pack2<T>::unpack(v.y); out[g * 8 + 2] = f.x; out[g * 8 + 3] = f.y;
    f = Pack2<T>::unpack(v.z); out[g * 8 + 4] = f.x; out[g * 8 + 5] = f.y;
    f = Pack2<T>::unpack(v.w); out[g * 8 + 6] = f.x; out[g * 8 + 7] = f.y;
  }
}

template <int ACT, typename T>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_act_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b0,
                const __grid_constant__ CUtensorMap tmap_b1, const __grid_constant__ CUtensorMap tmap_c,
                const Params p) {
  constexpr bool kSwiglu = (ACT == B200_ACT_SWIGLU);
  constexpr int OUT_COLS = kSwiglu ? 128 : 256;  // output columns per tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SMEM_BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b0);
    tma_prefetch_desc(&tmap_b1);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int m_blk, n_blk, kb0, kb1;
        tile_coords(p, tile, m_blk, n_blk, kb0, kb1);
        const int m0 = m_blk * BM;
        const int n0 = n_blk * OUT_COLS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          uint8_t* sa = smem + SMEM_A_OFF + stage * A_STAGE_BYTES;
          uint8_t* sb = smem + SMEM_B_OFF + stage * B_STAGE_BYTES;
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
          if constexpr (kSwiglu) {
            tma_load_2d(sb, &tmap_b0, &full_bar[stage], kb * BK, n0);                       // gate rows
            tma_load_2d(sb + B_STAGE_BYTES / 2, &tmap_b1, &full_bar[stage], kb * BK, n0);   // up rows
          } else {
            tma_load_2d(sb, &tmap_b0, &full_bar[stage], kb * BK, n0);
            tma_load_2d(sb + B_STAGE_BYTES / 2, &tmap_b0, &full_bar[stage], kb * BK, n0 + 128);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform control flow (descriptors stay in uniform registers); one elected lane issues
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, Pack2<T>::kIsBf16, false, false);
    const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem + SMEM_A_OFF), 16, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem + SMEM_B_OFF), 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m_blk_, n_blk_, kb0, kb1;
      tile_coords(p, tile, m_blk_, n_blk_, kb0, kb1);
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (A_STAGE_BYTES >> 4));
        const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (B_STAGE_BYTES >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_ss(d_tmem, ad + static_cast<uint64_t>(k * 2), bd + static_cast<uint64_t>(k * 2), idesc,
                    (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);  // accumulator ready for the epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int ep_warp = warp_idx - 4;            // TMEM lane quadrant
    const int ep_tid = threadIdx.x - 128;        // 0..127
    const int row = ep_warp * 32 + lane;         // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(ep_warp * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m_blk, n_blk, kb0, kb1;
      tile_coords(p, tile, m_blk, n_blk, kb0, kb1);
      const int m0 = m_blk * BM;
      const int n0 = n_blk * OUT_COLS;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_addr + static_cast<uint32_t>(acc * BN);

      if (p.k_splits > 1) {
        // split-K: raw fp32 accumulators of this split -> workspace; bias / activation happen in splitk_reduce_kernel
        const int split = tile % p.k_splits;
        const int64_t ld = static_cast<int64_t>(p.num_n_blocks) * BN;
        float* dst = p.partial + (static_cast<int64_t>(split) * p.M + (m0 + row)) * ld + static_cast<int64_t>(n_blk) * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(t_acc + c * 32, v);
          tmem_wait_ld();
          if (m0 + row < p.M) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(dst + c * 32 + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }

#pragma unroll 1
      for (int chunk = 0; chunk < OUT_COLS / 64; ++chunk) {
        uint8_t* cs = smem + SMEM_C_OFF + cbuf * C_BUF_BYTES;
        // the staging buffer may still be read by the TMA store issued two chunks ago
        if (ep_tid == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = chunk * 64 + half * 32;  // output column inside the tile
          uint32_t v[32];
          float f[32];
          if constexpr (kSwiglu) {
            uint32_t u[32];
            tmem_ld_x32(t_acc + col, v);
            tmem_ld_x32(t_acc + 128 + col, u);
            tmem_wait_ld();
            float bg[32], bu[32];
            load_bias32<T>(p.bias0, n0 + col, p.N_out, bg);
            load_bias32<T>(p.bias1, n0 + col, p.N_out, bu);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float g = __uint_as_float(v[i]) + bg[i];
              const float up = __uint_as_float(u[i]) + bu[i];
              f[i] = silu(g) * up;
            }
          } else {
            tmem_ld_x32(t_acc + col, v);
            tmem_wait_ld();
            float b[32];
            load_bias32<T>(p.bias0, n0 + col, p.N_out, b);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = apply_act<ACT>(__uint_as_float(v[i]) + b[i]);
          }
          if (chunk == OUT_COLS / 64 - 1 && half == 1) {
            // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
          }
          // 32 values -> 64 bytes = four 16-byte chunks of this row, 128-byte swizzle
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            pk.x = Pack2<T>::pack(f[q * 8 + 0], f[q * 8 + 1]);
            pk.y = Pack2<T>::pack(f[q * 8 + 2], f[q * 8 + 3]);
            pk.z = Pack2<T>::pack(f[q * 8 + 4], f[q * 8 + 5]);
            pk.w = Pack2<T>::pack(f[q * 8 + 6], f[q * 8 + 7]);
            const int c16 = half * 4 + q;
            *reinterpret_cast<uint4*>(cs + row * 128 + ((c16 ^ (row & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (ep_tid == 0) {
          tma_store_2d(&tmap_c, cs, n0 + chunk * 64, m0);
          tma_store_commit();
        }
        cbuf ^= 1;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (ep_tid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// split-K second pass: sum the per-split fp32 partials, add bias, apply the activation (SwiGLU pairs gate/up columns
// of the same tile), convert to 16 bit. One thread per 8 output columns.
template <int ACT, typename T>
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int k_splits, int M, int N_out, int num_n_blocks,
                                     const void* __restrict__ bias0, const void* __restrict__ bias1, T* __restrict__ y,
                                     int64_t ldy) {
  constexpr bool kSwiglu = (ACT == B200_ACT_SWIGLU);
  constexpr int OUT_COLS = kSwiglu ? 128 : 256;
  const int vec_per_row = (N_out + 7) / 8;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(M) * vec_per_row) return;
  const int m = static_cast<int>(idx / vec_per_row);
  const int n = static_cast<int>(idx % vec_per_row) * 8;
  const int nb = n / OUT_COLS, c = n % OUT_COLS;
  const int64_t ld = static_cast<int64_t>(num_n_blocks) * BN;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, u[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < k_splits; ++s) {
    const float* src = partial + (static_cast<int64_t>(s) * M + m) * ld + static_cast<int64_t>(nb) * BN + c;
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
    a[0] += v0.x; a[1] += v0.y; a[2] += v0.z; a[3] += v0.w; a[4] += v1.x; a[5] += v1.y; a[6] += v1.z; a[7] += v1.w;
    if constexpr (kSwiglu) {
      const float4 w0 = *reinterpret_cast<const float4*>(src + 128), w1 = *reinterpret_cast<const float4*>(src + 132);
      u[0] += w0.x; u[1] += w0.y; u[2] += w0.z; u[3] += w0.w; u[4] += w1.x; u[5] += w1.y; u[6] += w1.z; u[7] += w1.w;
    }
  }
  float b0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto load_bias = [&](const void* bias, float* out) {
    if (bias == nullptr) return;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(bias) + n));
    float2 f;
    f = Pack2<T>::unpack(v.x); out[0] = f.x; out[1] = f.y;
    f = Pack2<T>::unpack(v.y); out[2] = f.x; out[3] = f.y;
    f = Pack2<T>::unpack(v.z); out[4] = f.x; out[5] = f.y;
    f = Pack2<T>::unpack(v.w); out[6] = f.x; out[7] = f.y;
  };
  load_bias(bias0, b0);
  if constexpr (kSwiglu) load_bias(bias1, b1);
  float r[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if constexpr (kSwiglu) r[e] = silu(a[e] + b0[e]) * (u[e] + b1[e]);
    else r[e] = apply_act<ACT>(a[e] + b0[e]);
  }
  uint4 pk;
  pk.x = Pack2<T>::pack(r[0], r[1]); pk.y = Pack2<T>::pack(r[2], r[3]);
  pk.z = Pack2<T>::pack(r[4], r[5]); pk.w = Pack2<T>::pack(r[6], r[7]);
  *reinterpret_cast<uint4*>(y + static_cast<int64_t>(m) * ldy + n) = pk;
}

template <int ACT, typename T>
int launch(const CUtensorMap& ta, const CUtensorMap& tb0, const CUtensorMap& tb1, const CUtensorMap& tc,
           const Params& p, cudaStream_t stream) {
  auto kern = gemm_act_kernel<ACT, T>;
  static bool attr_set = false;  // benign race: setting the attribute twice is harmless
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  int max_ctas = sm_count();
  if (sm_limit() > 0 && sm_limit() < max_ctas) max_ctas = sm_limit();  // leave SMs to a concurrent collective
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb0, tb1, tc, p);
  B200_CUDA_OK(cudaGetLastError());
  if (p.k_splits > 1) {
    const int64_t total = static_cast<int64_t>(p.M) * ((p.N_out + 7) / 8);
    splitk_reduce_kernel<ACT, T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        p.partial, p.k_splits, p.M, p.N_out, p.num_n_blocks, p.bias0, p.bias1, static_cast<T*>(p.y), p.ldy);
    B200_CUDA_OK(cudaGetLastError());
  }
  return B200_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cluster of 2, tcgen05.mma cta_group::2, M = 256): each CTA stages its own 128 rows of A and HALF of
// the B tile (128 of the 256 accumulator columns' weight rows); one MMA issued by the leader CTA reads both SMs' shared
// memory, so every weight byte is fetched into shared memory once per 256 output rows instead of once per 128. Per CTA
// and k-block that is 32 KB instead of 48 KB of TMA traffic / shared-memory fill, which buys two more pipeline stages
// (6 x 32 KB) and less power under the 1 kW cap. Each CTA keeps its 128 x 256 fp32 accumulator (double buffered) in its
// own TMEM and runs the same fused epilogue. SwiGLU falls out naturally: CTA 0 stages the gate rows, CTA 1 the up rows.
// ---------------------------------------------------------------------------------------------------------------
namespace pair {
constexpr int P_STAGES = 6;
constexpr int P_A_STAGE_BYTES = 128 * BK * 2;  // 16 KB
constexpr int P_B_STAGE_BYTES = 128 * BK * 2;  // 16 KB: this CTA's half of the 256-column B tile
constexpr int P_SMEM_A_OFF = 0;
constexpr int P_SMEM_B_OFF = P_SMEM_A_OFF + P_STAGES * P_A_STAGE_BYTES;
constexpr int P_SMEM_C_OFF = P_SMEM_B_OFF + P_STAGES * P_B_STAGE_BYTES;
constexpr int P_SMEM_BAR_OFF = P_SMEM_C_OFF + 2 * C_BUF_BYTES;
constexpr int P_SMEM_BYTES = P_SMEM_BAR_OFF + 256 + 1024;
}  // namespace pair

template <int ACT, typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_act_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b0,
                     const __grid_constant__ CUtensorMap tmap_b1, const __grid_constant__ CUtensorMap tmap_c,
                     const Params p) {
  using namespace pair;
  constexpr bool kSwiglu = (ACT == B200_ACT_SWIGLU);
  constexpr int OUT_COLS = kSwiglu ? 128 : 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_SMEM_BAR_OFF);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tmem_full_bar = empty_bar + P_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b0);
    tma_prefetch_desc(&tmap_b1);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // only the leader's is used: one arrive.expect_tx, bytes of both CTAs
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit from the leader
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);   // multicast tcgen05.commit from the leader
      mbar_init(&tmem_empty_bar[s], 2);  // leader's: one elected arrival per CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // p.num_m_blocks counts 256-row blocks for this kernel
  auto coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = (GROUP_M / 2) * p.num_n_blocks;
    const int group = tile / per_group;
    const int first_m = group * (GROUP_M / 2);
    const int rows_in_group = min(GROUP_M / 2, p.num_m_blocks - first_m);
    const int in_group = tile - group * per_group;
    m_blk = first_m + in_group % rows_in_group;
    n_blk = in_group / rows_in_group;
  };

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        coords(tile, m_blk, n_blk);
        const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
        const int n0 = n_blk * OUT_COLS;
        const CUtensorMap* tb = (kSwiglu && rank == 1) ? &tmap_b1 : &tmap_b0;     // SwiGLU: CTA0 gate rows, CTA1 up rows
        const int brow = kSwiglu ? n0 : n0 + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (P_A_STAGE_BYTES + P_B_STAGE_BYTES));
          tma_load_2d_2sm(smem + P_SMEM_A_OFF + stage * P_A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BK, m0);
          tma_load_2d_2sm(smem + P_SMEM_B_OFF + stage * P_B_STAGE_BYTES, tb, &full_bar[stage], kb * BK, brow);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, BN, Pack2<T>::kIsBf16, false, false);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem + P_SMEM_A_OFF), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem + P_SMEM_B_OFF), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (P_A_STAGE_BYTES >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (P_B_STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_ss_2sm(d_tmem, ad + static_cast<uint64_t>(k * 2), bd + static_cast<uint64_t>(k * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);  // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(&tmem_full_bar[acc], 0x3);  // both CTAs' epilogues
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ep_warp = warp_idx - 4;
    const int ep_tid = threadIdx.x - 128;
    const int row = ep_warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(ep_warp * 32) << 16;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      coords(tile, m_blk, n_blk);
      const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
      const int n0 = n_blk * OUT_COLS;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_addr + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int chunk = 0; chunk < OUT_COLS / 64; ++chunk) {
        uint8_t* cs = smem + P_SMEM_C_OFF + cbuf * C_BUF_BYTES;
        if (ep_tid == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = chunk * 64 + half * 32;
          uint32_t v[32];
          float f[32];
          if constexpr (kSwiglu) {
            uint32_t u[32];
            tmem_ld_x32(t_acc + col, v);
            tmem_ld_x32(t_acc + 128 + col, u);
            tmem_wait_ld();
            float bg[32], bu[32];
            load_bias32<T>(p.bias0, n0 + col, p.N_out, bg);
            load_bias32<T>(p.bias1, n0 + col, p.N_out, bu);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = silu(__uint_as_float(v[i]) + bg[i]) * (__uint_as_float(u[i]) + bu[i]);
          } else {
            tmem_ld_x32(t_acc + col, v);
            tmem_wait_ld();
            float b[32];
            load_bias32<T>(p.bias0, n0 + col, p.N_out, b);
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = apply_act<ACT>(__uint_as_float(v[i]) + b[i]);
          }
          if (chunk == OUT_COLS / 64 - 1 && half == 1) {
            // all TMEM reads of this accumulator stage are done in this CTA: one elected arrival on the leader's barrier
            tc_fence_before();
            named_bar_sync(2, 128);
            if (ep_tid == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            pk.x = Pack2<T>::pack(f[q * 8 + 0], f[q * 8 + 1]);
            pk.y = Pack2<T>::pack(f[q * 8 + 2], f[q * 8 + 3]);
            pk.z = Pack2<T>::pack(f[q * 8 + 4], f[q * 8 + 5]);
            pk.w = Pack2<T>::pack(f[q * 8 + 6], f[q * 8 + 7]);
            const int c16 = half * 4 + q;
            *reinterpret_cast<uint4*>(cs + row * 128 + ((c16 ^ (row & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (ep_tid == 0) {
          tma_store_2d(&tmap_c, cs, n0 + chunk * 64, m0);
          tma_store_commit();
        }
        cbuf ^= 1;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (ep_tid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory / barriers are referenced until the very end
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int ACT, typename T>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb0, const CUtensorMap& tb1, const CUtensorMap& tc, const Params& p,
                cudaStream_t stream) {
  auto kern = gemm_act_pair_kernel<ACT, T>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::P_SMEM_BYTES));
    attr_set = true;
  }
  int max_ctas = sm_count();
  if (sm_limit() > 0 && sm_limit() < max_ctas) max_ctas = sm_limit();
  int pairs = max_ctas / 2;
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  if (pairs < 1) pairs = 1;
  kern<<<2 * pairs, NUM_THREADS, pair::P_SMEM_BYTES, stream>>>(ta, tb0, tb1, tc, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

template <typename T>
int dispatch(int act, const CUtensorMap& ta, const CUtensorMap& tb0, const CUtensorMap& tb1, const CUtensorMap& tc,
             const Params& p, cudaStream_t stream) {
  if (p.use_pair) {
    switch (act) {
      case B200_ACT_NONE: return launch_pair<B200_ACT_NONE, T>(ta, tb0, tb1, tc, p, stream);
      case B200_ACT_GELU_TANH: return launch_pair<B200_ACT_GELU_TANH, T>(ta, tb0, tb1, tc, p, stream);
      case B200_ACT_GELU_ERF: return launch_pair<B200_ACT_GELU_ERF, T>(ta, tb0, tb1, tc, p, stream);
      case B200_ACT_RELU: return launch_pair<B200_ACT_RELU, T>(ta, tb0, tb1, tc, p, stream);
      case B200_ACT_SWIGLU: return launch_pair<B200_ACT_SWIGLU, T>(ta, tb0, tb1, tc, p, stream);
      default: return set_error(B200_ERR_INVALID_ARGUMENT, "unknown activation %d", act);
    }
  }
  switch (act) {
    case B200_ACT_NONE: return launch<B200_ACT_NONE, T>(ta, tb0, tb1, tc, p, stream);
    case B200_ACT_GELU_TANH: return launch<B200_ACT_GELU_TANH, T>(ta, tb0, tb1, tc, p, stream);
    case B200_ACT_GELU_ERF: return launch<B200_ACT_GELU_ERF, T>(ta, tb0, tb1, tc, p, stream);
    case B200_ACT_RELU: return launch<B200_ACT_RELU, T>(ta, tb0, tb1, tc, p, stream);
    case B200_ACT_SWIGLU: return launch<B200_ACT_SWIGLU, T>(ta, tb0, tb1, tc, p, stream);
    default: return set_error(B200_ERR_INVALID_ARGUMENT, "unknown activation %d", act);
  }
}

}  // namespace gemm

static int check_ptr16(const void* p, const char* name) {
  if (p == nullptr) return set_error(B200_ERR_INVALID_ARGUMENT, "%s is NULL", name);
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0)
    return set_error(B200_ERR_INVALID_ARGUMENT, "%s must be 16-byte aligned", name);
  return B200_OK;
}

// y[T,N] = act(x[T,K] w[N,K]^T + b)      or      y = silu(x wg^T + bg) * (x w^T + b)  when act == SWIGLU
// number of K splits for a problem with `tiles` output tiles and `kblocks` 64-wide K blocks: only skinny problems
// (fewer output tiles than SMs) are split, into enough pieces to fill the machine, keeping >= 4 k-blocks per split
static int choose_k_splits(int tiles, int kblocks) {
  const int sms = sm_count();
  if (tiles >= sms || kblocks < 8) return 1;
  int splits = sms / tiles;  // one wave: every CTA resident at once, the fp32 reduce pass stays small
  if (splits > kblocks / 4) splits = kblocks / 4;
  if (splits > 32) splits = 32;
  return splits < 1 ? 1 : splits;
}

int64_t linear_act_workspace_bytes(int64_t T, int K, int N, int act) {
  if (T <= 0 || K <= 0 || N <= 0) return 0;
  const int out_cols = (act == B200_ACT_SWIGLU) ? 128 : 256;
  const int64_t m_blocks = (T + gemm::BM - 1) / gemm::BM, n_blocks = (N + out_cols - 1) / out_cols;
  if (m_blocks * n_blocks > 0x7fffffff) return 0;
  const int splits = choose_k_splits(static_cast<int>(m_blocks * n_blocks), (K + gemm::BK - 1) / gemm::BK);
  if (splits <= 1) return 0;
  return static_cast<int64_t>(splits) * T * n_blocks * gemm::BN * static_cast<int64_t>(sizeof(float));
}

int linear_act_impl(const void* x, int64_t ldx, const void* w, const void* b, const void* w_gate, const void* b_gate,
                    void* y, int64_t ldy, int64_t T, int K, int N, int act, int dtype, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream) {
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_FP16, "dtype must be bf16 or fp16");
  B200_CHECK_ARG(T >= 0 && K > 0 && N > 0, "bad sizes T=%lld K=%d N=%d", (long long)T, K, N);
  B200_CHECK_ARG(T <= 0x7fffffffLL, "T too large");
  B200_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "K and N must be multiples of 8 (got K=%d N=%d)", K, N);
  B200_CHECK_ARG(ldx >= K && ldx % 8 == 0 && ldy >= N && ldy % 8 == 0, "row strides must be >= width and multiples of 8");
  if (T == 0) return B200_OK;
  int rc;
  if ((rc = check_ptr16(x, "x"))) return rc;
  if ((rc = check_ptr16(w, "weight"))) return rc;
  if ((rc = check_ptr16(y, "y"))) return rc;
  const bool swiglu = (act == B200_ACT_SWIGLU);
  if (swiglu) {
    if ((rc = check_ptr16(w_gate, "gate weight"))) return rc;
  } else {
    B200_CHECK_ARG(w_gate == nullptr && b_gate == nullptr, "gate weight/bias given but activation is not SwiGLU");
  }
  if (b != nullptr && (rc = check_ptr16(b, "bias"))) return rc;
  if (b_gate != nullptr && (rc = check_ptr16(b_gate, "gate bias"))) return rc;

  CUtensorMap ta, tb0, tb1, tc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(T)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t box[2] = {gemm::BK, gemm::BM};
    if ((rc = encode_tmap_sw128_16b(&ta, x, 2, dims, strides, box))) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {gemm::BK, 128};
    // SwiGLU: b0 = gate rows, b1 = up rows; otherwise both maps describe the same weight
    if ((rc = encode_tmap_sw128_16b(&tb0, swiglu ? w_gate : w, 2, dims, strides, box))) return rc;
    if ((rc = encode_tmap_sw128_16b(&tb1, w, 2, dims, strides, box))) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(T)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldy) * 2};
    uint32_t box[2] = {64, gemm::BM};
    if ((rc = encode_tmap_sw128_16b(&tc, y, 2, dims, strides, box))) return rc;
  }
  gemm::Params p;
  p.M = static_cast<int>(T);
  p.N_out = N;
  p.K = K;
  p.bias0 = swiglu ? b_gate : b;
  p.bias1 = swiglu ? b : nullptr;
  const int out_cols = swiglu ? 128 : 256;
  p.num_m_blocks = (p.M + gemm::BM - 1) / gemm::BM;
  p.num_n_blocks = (N + out_cols - 1) / out_cols;
  p.num_k_blocks = (K + gemm::BK - 1) / gemm::BK;
  p.k_splits = 1;
  p.partial = nullptr;
  p.y = y;
  p.ldy = ldy;
  const int64_t ws_need = linear_act_workspace_bytes(T, K, N, act);
  if (ws_need > 0 && workspace != nullptr && workspace_bytes >= ws_need &&
      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
    p.k_splits = choose_k_splits(p.num_m_blocks * p.num_n_blocks, p.num_k_blocks);
    p.partial = static_cast<float*>(workspace);
  }
  // large problems: CTA-pair kernel (256-row tiles). B200_GEMM_PAIR=0 in the environment keeps the single-CTA kernel.
  static const bool pair_enabled = [] { const char* e = getenv("B200_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  p.use_pair = (pair_enabled && p.k_splits == 1 && T >= 1024) ? 1 : 0;
  if (p.use_pair) p.num_m_blocks = (p.M + 255) / 256;
  p.kb_per_split = (p.num_k_blocks + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.num_tiles = p.num_m_blocks * p.num_n_blocks * p.k_splits;
  if (dtype == B200_DTYPE_BF16) return gemm::dispatch<__nv_bfloat16>(act, ta, tb0, tb1, tc, p, stream);
  return gemm::dispatch<__half>(act, ta, tb0, tb1, tc, p, stream);
}

}  // namespace b200

extern "C" {

int64_t b200_linear_act_workspace_bytes(int64_t T, int K, int N, int act) {
  return b200::linear_act_workspace_bytes(T, K, N, act);
}

int b200_linear_act(const void* x, int64_t ldx, const void* w, const void* b, const void* w_gate, const void* b_gate,
                    void* y, int64_t ldy, int64_t T, int K, int N, int act, void* workspace, int64_t workspace_bytes,
                    int dtype, void* stream) {
  return b200::linear_act_impl(x, ldx, w, b, w_gate, b_gate, y, ldy, T, K, N, act, dtype, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

static int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

int64_t b200_fused_mlp_workspace_bytes(int64_t T, int h, int i) {
  if (T < 0 || i <= 0) return 0;
  // intermediate [T, i] (16 bit) + split-K partials of the skinnier of the two GEMMs (worst case over activations)
  const int64_t inter = align256(T * static_cast<int64_t>(i) * 2);
  int64_t sk = b200::linear_act_workspace_bytes(T, h, i, B200_ACT_SWIGLU);
  const int64_t sk1 = b200::linear_act_workspace_bytes(T, h, i, B200_ACT_NONE);
  const int64_t sk2 = b200::linear_act_workspace_bytes(T, i, h, B200_ACT_NONE);
  if (sk1 > sk) sk = sk1;
  if (sk2 > sk) sk = sk2;
  return inter + sk;
}

int b200_fused_mlp(const void* x, int64_t ldx, const void* w_up, const void* b_up, const void* w_gate,
                   const void* b_gate, const void* w_down, const void* b_down, void* y, int64_t ldy, int64_t T, int h,
                   int i, int h_out, int act, void* workspace, int64_t workspace_bytes, int dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(act >= B200_ACT_GELU_TANH && act <= B200_ACT_SWIGLU, "activation %d is not a FusedMLP activation", act);
  if (T == 0) return B200_OK;
  if (workspace == nullptr || workspace_bytes < b200_fused_mlp_workspace_bytes(T, h, i))
    return set_error(B200_ERR_WORKSPACE, "fused_mlp workspace too small: need %lld bytes, got %lld",
                     (long long)b200_fused_mlp_workspace_bytes(T, h, i), (long long)workspace_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // GEMM1 + bias + activation (SwiGLU: gate/up pair) -> bf16 intermediate (stays L2-resident per row panel)
  const int64_t inter = align256(T * static_cast<int64_t>(i) * 2);
  void* sk_ws = static_cast<char*>(workspace) + inter;
  const int64_t sk_bytes = workspace_bytes - inter;
  int rc = linear_act_impl(x, ldx, w_up, b_up, w_gate, b_gate, workspace, i, T, h, i, act, dtype, sk_ws, sk_bytes, s);
  if (rc) return rc;
  // GEMM2 + bias
  return linear_act_impl(workspace, i, w_down, b_down, nullptr, nullptr, y, ldy, T, i, h_out, B200_ACT_NONE, dtype, sk_ws,
                         sk_bytes, s);
}

}  // extern "C"
