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
  // L2 rasterisation: tiles are visited in groups of `group_m` row-blocks x all column-blocks, row-block fastest, so the
  // group's activation panel (group_m x BM rows) stays L2-resident while the weight is streamed once per group
  int group_m;
  // L2 eviction hints of the TMA traffic (pair kernel): the activation panel of the raster group is re-read by the next
  // waves (evict-last), a weight slab is shared by the CTAs of ONE wave and then dead until the next group (evict-first
  // / normal), the output is written once
  uint64_t hint_a, hint_b, hint_c;
};

__device__ __forceinline__ void tile_coords(const Params& p, int tile_in, int& m_blk, int& n_blk, int& kb0, int& kb1) {
  const int split = tile_in % p.k_splits;
  const int tile = tile_in / p.k_splits;
  kb0 = split * p.kb_per_split;
  kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
  const int per_group = p.group_m * p.num_n_blocks;
  const int group = tile / per_group;
  const int first_m = group * p.group_m;
  const int rows_in_group = min(p.group_m, p.num_m_blocks - first_m);
  const int in_group = tile - group * per_group;
  m_blk = first_m + in_group % rows_in_group;
  n_blk = in_group / rows_in_group;
}

// Row-blocks per raster group. All resident tiles sweep K in lockstep, so what must fit in L2 for the activation panel to
// be re-used by the NEXT wave is one wave's footprint: (group_m + tiles_per_wave / group_m) slabs of block_rows x K x 2
// bytes. Measured on B200 (bench.py --group-rows, C3): K=4096 2048/4096 rows 8.10/8.01 ms per step, 8192 rows 8.73 ms,
// 32768 rows 9.42 ms (DRAM reads of the up+gate GEMM 3.2 GB at 2048 rows, 6.9 GB at 8192): the usable budget is ~48 MB of
// the 126 MB L2 (two partitions, lines replicated across them), not the nominal size.
inline int choose_group_m(int K, int block_rows, int num_m_blocks) {
  int g;
  if (gemm_group_rows() > 0) {
    g = gemm_group_rows() / block_rows;
  } else {
    const double slab = static_cast<double>(block_rows) * K * 2.0;
    const double tiles_per_wave = sm_count() / (block_rows == 256 ? 2.0 : 1.0);
    g = 8 * 256 / block_rows;
    for (int cand = 32 * 256 / block_rows; cand > g; cand /= 2) {
      if ((cand + tiles_per_wave / cand) * slab <= 48.0e6) { g = cand; break; }
    }
  }
  if (g < 1) g = 1;
  if (g > num_m_blocks) g = num_m_blocks;
  return g;
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

// The same activations on a pair of columns with packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2: two elements per issue
// slot; the operation order is the scalar one, so the results are bit-identical). A K = 768 GEMM (GPT-2 fc1) spends as
// long in the GELU epilogue of a 256-column tile as in its 12 k-blocks of MMAs: 1217 TFLOP/s with the scalar epilogue
// against 1466 without an activation (tests/gemm_epilogue_probe.py).
template <int ACT>
__device__ __forceinline__ float2 apply_act2(float2 x) {
  if constexpr (ACT == B200_ACT_GELU_TANH) {
    const float2 k0 = make_float2(0.7978845608028654f, 0.7978845608028654f), k1 = make_float2(0.044715f, 0.044715f);
    const float2 one = make_float2(1.0f, 1.0f), half = make_float2(0.5f, 0.5f);
    const float2 f = __ffma2_rn(__fmul2_rn(k1, x), x, one);
    const float2 inner = __fmul2_rn(__fmul2_rn(k0, x), f);
    const float2 t = make_float2(fast_tanh(inner.x), fast_tanh(inner.y));
    return __fmul2_rn(__fmul2_rn(half, x), __fadd2_rn(one, t));
  } else {
    return make_float2(apply_act<ACT>(x.x), apply_act<ACT>(x.y));
  }
}
// f[i] = act(v[i] + b[i]) for 32 columns, two at a time
template <int ACT>
__device__ __forceinline__ void bias_act32(const uint32_t (&v)[32], const float (&b)[32], float (&f)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float2 x = __fadd2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(b[i], b[i + 1]));
    const float2 r = apply_act2<ACT>(x);
    f[i] = r.x;
    f[i + 1] = r.y;
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
            bias_act32<ACT>(v, b, f);
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
  static bool attr_set[64] = {};  // per device
  B200_CUDA_OK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), SMEM_BYTES, attr_set));
  int max_ctas = sm_count();
  if (sm_limit() > 0 && sm_limit() < max_ctas) max_ctas = sm_limit();  // leave SMs to a concurrent collective
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb0, tb1, tc, p);
  B200_CUDA_OK(cudaGetLastError());
  static const char* const names[5] = {"gemm_act_kernel<NONE>", "gemm_act_kernel<GELU_TANH>", "gemm_act_kernel<GELU_ERF>",
                                       "gemm_act_kernel<RELU>", "gemm_act_kernel<SWIGLU>"};
  note_launch(names[ACT], true);
  if (p.k_splits > 1) {
    const int64_t total = static_cast<int64_t>(p.M) * ((p.N_out + 7) / 8);
    splitk_reduce_kernel<ACT, T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        p.partial, p.k_splits, p.M, p.N_out, p.num_n_blocks, p.bias0, p.bias1, static_cast<T*>(p.y), p.ldy);
    B200_CUDA_OK(cudaGetLastError());
    note_launch("splitk_reduce_kernel");
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

// EPI = number of epilogue warpgroups (1 or 2). With two, warpgroup e takes the 64-column chunks e, e + 2 of every tile
// and owns one of the two staging buffers: for short K with a GELU (GPT-2 fc1: 12 k-blocks, 6144 MMA cycles per tile) the
// epilogue of a 256-column tile is as long as the tile's MMAs, so one warpgroup makes the kernel epilogue-bound
// (tests/gemm_epilogue_probe.py: 1259 TFLOP/s with GELU vs 1470 without).
template <int ACT, typename T, int EPI = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 128 * EPI, 1)
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
      mbar_init(&tmem_empty_bar[s], 2 * EPI);  // leader's: one elected arrival per epilogue warpgroup of either CTA
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
    const int per_group = p.group_m * p.num_n_blocks;
    const int group = tile / per_group;
    const int first_m = group * p.group_m;
    const int rows_in_group = min(p.group_m, p.num_m_blocks - first_m);
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
          tma_load_2d_2sm_hint(smem + P_SMEM_A_OFF + stage * P_A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BK, m0, p.hint_a);
          tma_load_2d_2sm_hint(smem + P_SMEM_B_OFF + stage * P_B_STAGE_BYTES, tb, &full_bar[stage], kb * BK, brow, p.hint_b);
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
    // ===================== epilogue (both CTAs, own 128 rows; EPI warpgroups) =====================
    static_assert(EPI == 1 || EPI == 2, "one or two epilogue warpgroups");
    const int wg = EPI == 1 ? 0 : ((warp_idx - 4) >> 2);   // epilogue warpgroup (a compile-time 0 with one: keeps that build at 209 registers)
    const int ep_warp = (warp_idx - 4) & 3;        // TMEM lane quadrant
    const int ep_tid = threadIdx.x - 128 - wg * 128;
    const int row = ep_warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(ep_warp * 32) << 16;
    // named barriers of this warpgroup, ids as immediates (1, 2 with one warpgroup as before; 1-4 with two)
    auto sync_chunk = [&]() {
      if (EPI == 1 || wg == 0) named_bar_sync_imm<1>(128);
      else named_bar_sync_imm<3>(128);
    };
    auto sync_done = [&]() {
      if (EPI == 1 || wg == 0) named_bar_sync_imm<2>(128);
      else named_bar_sync_imm<4>(128);
    };
    constexpr int kChunks = OUT_COLS / 64;
    constexpr int kLastChunk0 = kChunks - EPI;     // the last chunk of warpgroup 0 (warpgroup e: + e)
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = EPI == 2 ? wg : 0;                  // two warpgroups: one staging buffer each; one: both, alternating
    for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      coords(tile, m_blk, n_blk);
      const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
      const int n0 = n_blk * OUT_COLS;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_addr + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int chunk = wg; chunk < kChunks; chunk += EPI) {
        uint8_t* cs = smem + P_SMEM_C_OFF + cbuf * C_BUF_BYTES;
        if (ep_tid == 0) {
          if constexpr (EPI == 2) tma_store_wait_read<0>();   // this warpgroup's only buffer: its last store has been read
          else tma_store_wait_read<1>();
        }
        sync_chunk();
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
            bias_act32<ACT>(v, b, f);
          }
          if (chunk == kLastChunk0 + wg && half == 1) {
            // all TMEM reads of this accumulator stage by this warpgroup are done: one elected arrival on the leader's barrier
            tc_fence_before();
            sync_done();
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
        sync_chunk();
        if (ep_tid == 0) {
          tma_store_2d_hint(&tmap_c, cs, n0 + chunk * 64, m0, p.hint_c);
          tma_store_commit();
        }
        if constexpr (EPI == 1) cbuf ^= 1;
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

constexpr int kShortKPlain = 12;  // k-blocks up to which a GEMM WITHOUT a GELU gets two epilogue warpgroups (K=512: 0.093 -> 0.086 ms, K=768: 0.1055 -> 0.1034, K=1536: equal)
template <int ACT, typename T, int EPI = 1>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb0, const CUtensorMap& tb1, const CUtensorMap& tc, const Params& p,
                cudaStream_t stream) {
  if constexpr (EPI == 1 && ACT != B200_ACT_SWIGLU) {
    // short K: the epilogue of a tile is as long as its MMAs -> two epilogue warpgroups (B200_GEMM_EPI_WGS=1/2 forces)
    constexpr bool kGelu = ACT == B200_ACT_GELU_TANH || ACT == B200_ACT_GELU_ERF;
    const char* e = getenv("B200_GEMM_EPI_WGS");
    const bool two = e != nullptr ? e[0] == '2' : p.num_k_blocks <= (kGelu ? 16 : kShortKPlain);
    if (two) return launch_pair<ACT, T, 2>(ta, tb0, tb1, tc, p, stream);
  }
  auto kern = gemm_act_pair_kernel<ACT, T, EPI>;
  static bool attr_set[64] = {};  // per device
  B200_CUDA_OK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), pair::P_SMEM_BYTES, attr_set));
  int max_ctas = sm_count();
  if (sm_limit() > 0 && sm_limit() < max_ctas) max_ctas = sm_limit();
  int pairs = max_ctas / 2;
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  if (pairs < 1) pairs = 1;
  kern<<<2 * pairs, 128 + 128 * EPI, pair::P_SMEM_BYTES, stream>>>(ta, tb0, tb1, tc, p);
  B200_CUDA_OK(cudaGetLastError());
  static const char* const names[5] = {"gemm_act_pair_kernel<NONE>", "gemm_act_pair_kernel<GELU_TANH>", "gemm_act_pair_kernel<GELU_ERF>",
                                       "gemm_act_pair_kernel<RELU>", "gemm_act_pair_kernel<SWIGLU>"};
  static const char* const names2[5] = {"gemm_act_pair_kernel<NONE,2wg>", "gemm_act_pair_kernel<GELU_TANH,2wg>", "gemm_act_pair_kernel<GELU_ERF,2wg>",
                                        "gemm_act_pair_kernel<RELU,2wg>", "gemm_act_pair_kernel<SWIGLU,2wg>"};
  note_launch(EPI == 2 ? names2[ACT] : names[ACT], true);
  return B200_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// FusedMLP as ONE launch (CTA pairs, same pipeline as gemm_act_pair_kernel): the tile list interleaves the up(+gate)
// projection and the down projection by row groups —
//     P1(g0) P1(g1) P2(g0) P1(g2) P2(g1) ... P1(gN-1) P2(gN-2) P2(gN-1)
// — so the bf16 intermediate of group g is consumed one group later, while it is still L2-resident (group size chosen on
// the host from the L2 budget: reference intent kernels/mlp/fused_mlp.py:262-275, kernels/triton/mlp_kernels.py:87,:126
// "the intermediate never round-trips HBM"). A P2 tile needs every P1 tile of its 128-row half: the P1 epilogue publishes
// a per-half counter (TMA stores complete -> fence.proxy.async -> red.release.gpu), the P2 producer acquires it before the
// first TMA load of the intermediate. Tiles are assigned round-robin in list order and every dependency points backwards
// in that order. Tiles are handed out dynamically (one global counter, claimed in list order by the leader CTA's scheduler
// warp and passed to both CTAs of the pair through a 2-slot shared-memory ring), so P1 and P2 tiles of different cost stay
// balanced and the schedule cannot deadlock whatever part of the grid is resident; spins are bounded anyway.
// ---------------------------------------------------------------------------------------------------------------
struct FusedParams {
  int M;                      // tokens
  int N1, N2;                 // intermediate width (SwiGLU: gate/up pairs), output width
  const void* bias_p1_0;      // phase 1 bias (SwiGLU: gate bias)
  const void* bias_p1_1;      // SwiGLU: up bias
  const void* bias_p2;        // down-projection bias
  int num_m_blocks;           // 256-row blocks
  int n1_blocks, n2_blocks, k1_blocks, k2_blocks;
  int group_m, num_groups, rows_last;  // row-blocks per group, groups, row-blocks in the last group
  int lag;                    // P2 of group g is listed after P1 of group g + lag
  int debug_flags;            // bit 0: skip the publication fences (timing experiment only: results may be stale)
  int num_tiles;
  unsigned* ready;            // [num_m_blocks * 2] P1 tiles completed per 128-row half (zeroed before the launch)
};

struct FusedTile {
  int phase, m_blk, n_blk;
};

// Tile list: for s = 0 .. num_groups + lag - 1: [P1 tiles of group s (if s < num_groups)] [P2 tiles of group s - lag (if
// s >= lag)]. `seg` (shared memory, built once per CTA) holds the first tile index of every step s and of its P2 part.
constexpr int FUSED_MAX_STEPS = 192;
struct FusedSeg {
  int start[FUSED_MAX_STEPS + 1];   // first tile of step s
};
__device__ __forceinline__ int fused_rows(const FusedParams& p, int g) { return g == p.num_groups - 1 ? p.rows_last : p.group_m; }
__device__ inline void fused_build_segments(const FusedParams& p, FusedSeg* seg) {
  int t = 0;
  const int steps = p.num_groups + p.lag;
  for (int s = 0; s < steps; ++s) {
    seg->start[s] = t;
    if (s < p.num_groups) t += fused_rows(p, s) * p.n1_blocks;
    if (s >= p.lag) t += fused_rows(p, s - p.lag) * p.n2_blocks;
  }
  seg->start[steps] = t;
}
// `cur` is the caller's cursor (tiles arrive in increasing order for every role). Inside a step the P1 and P2 tiles are
// interleaved in proportion (Bresenham): a pair that works through the list sees ~n1 P1 tiles per n2 P2 tiles at any time,
// so the long epilogue of a short-K P1 tile (activation over 256 columns) hides behind the MMAs of a long-K P2 tile
// instead of stalling a run of P1 tiles.
__device__ __forceinline__ FusedTile fused_decode(const FusedParams& p, const FusedSeg* seg, int t, int& cur) {
  while (t >= seg->start[cur + 1]) ++cur;
  const int u = t - seg->start[cur];
  const int a = cur < p.num_groups ? fused_rows(p, cur) * p.n1_blocks : 0;
  const int b = cur >= p.lag ? fused_rows(p, cur - p.lag) * p.n2_blocks : 0;
  const long long tot = a + b;
  const int p2_before = static_cast<int>((static_cast<long long>(u) * b) / tot);
  const int p2_after = static_cast<int>((static_cast<long long>(u + 1) * b) / tot);
  FusedTile ft;
  int group, local;
  if (p2_after > p2_before) {
    ft.phase = 1;
    group = cur - p.lag;
    local = p2_before;
  } else {
    ft.phase = 0;
    group = cur;
    local = u - p2_before;
  }
  const int rows = fused_rows(p, group);
  ft.n_blk = local / rows;
  ft.m_blk = group * p.group_m + (local - ft.n_blk * rows);
  return ft;
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one 128-row x OUT_COLS output tile: TMEM -> registers -> bias/activation -> 16 bit -> swizzled smem -> TMA store
template <int ACT, typename T>
__device__ __forceinline__ void pair_epilogue_tile(uint8_t* smem_c, int& cbuf, uint32_t t_acc, const CUtensorMap* tmap_c,
                                                   const void* bias0, const void* bias1, int n_limit, int m0, int n0,
                                                   int row, int ep_tid, uint64_t* tmem_empty_bar_acc,
                                                   unsigned*& pending_publish, int debug_flags) {
  constexpr bool kSwiglu = (ACT == B200_ACT_SWIGLU);
  constexpr int OUT_COLS = kSwiglu ? 128 : 256;
#pragma unroll 1
  for (int chunk = 0; chunk < OUT_COLS / 64; ++chunk) {
    uint8_t* cs = smem_c + cbuf * C_BUF_BYTES;
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
        load_bias32<T>(bias0, n0 + col, n_limit, bg);
        load_bias32<T>(bias1, n0 + col, n_limit, bu);
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = silu(__uint_as_float(v[i]) + bg[i]) * (__uint_as_float(u[i]) + bu[i]);
      } else {
        tmem_ld_x32(t_acc + col, v);
        tmem_wait_ld();
        float b[32];
        load_bias32<T>(bias0, n0 + col, n_limit, b);
        bias_act32<ACT>(v, b, f);
      }
      if (chunk == OUT_COLS / 64 - 1 && half == 1) {
        tc_fence_before();
        named_bar_sync(2, 128);
        if (ep_tid == 0) mbar_arrive_cluster(tmem_empty_bar_acc, 0);
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
      tma_store_2d(tmap_c, cs, n0 + chunk * 64, m0);
      tma_store_commit();
      if (chunk == 1 && pending_publish != nullptr) {
        // deferred publication of the PREVIOUS P1 tile: all bulk groups but the two just committed are complete, i.e. all
        // of its rows are in global memory; waiting here (two chunks later) costs nothing, waiting right after the
        // tile's last store would put the store latency on the epilogue's critical path
        if (!(debug_flags & 1)) {
          tma_store_wait_all<2>();
          fence_proxy_async_all();
        }
        red_release_gpu_add(pending_publish, 1u);  // release at gpu scope: orders the completed stores before the counter
        pending_publish = nullptr;
      }
    }
    cbuf ^= 1;
  }
}

// cluster-scope mbarrier wait (the tile ring is written from the leader CTA into the peer's shared memory)
__device__ __forceinline__ void mbar_wait_cluster_long(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 24)) {
      printf("b200: fused_mlp scheduler wait timeout block=%d thread=%d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void st_shared_cluster_u32(const void* local_ptr, uint32_t rank, uint32_t v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

constexpr int SCHED_SLOTS = 4;

template <int ACT, typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
fused_mlp_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1a,
                      const __grid_constant__ CUtensorMap tmap_w1b, const __grid_constant__ CUtensorMap tmap_mid_st,
                      const __grid_constant__ CUtensorMap tmap_mid_ld, const __grid_constant__ CUtensorMap tmap_w2,
                      const __grid_constant__ CUtensorMap tmap_y, const FusedParams p) {
  using namespace pair;
  constexpr bool kSwiglu = (ACT == B200_ACT_SWIGLU);
  constexpr int OUT1 = kSwiglu ? 128 : 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_SMEM_BAR_OFF);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tmem_full_bar = empty_bar + P_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* sched_full = tmem_empty_bar + 2;           // per CTA: the leader's scheduler published a tile index
  uint64_t* sched_empty = sched_full + SCHED_SLOTS;    // leader's only: every consumer of both CTAs has read the slot
  volatile uint32_t* sched_tile = reinterpret_cast<volatile uint32_t*>(sched_empty + SCHED_SLOTS);
  uint32_t* tmem_ptr_smem = const_cast<uint32_t*>(sched_tile) + SCHED_SLOTS;
  FusedSeg* seg = reinterpret_cast<FusedSeg*>(smem + P_SMEM_BAR_OFF + 256);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp_idx == 3 && lane == 0) fused_build_segments(p, seg);
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_w1a); tma_prefetch_desc(&tmap_w1b); tma_prefetch_desc(&tmap_mid_st);
    tma_prefetch_desc(&tmap_mid_ld); tma_prefetch_desc(&tmap_w2); tma_prefetch_desc(&tmap_y);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2);
    }
    for (int s = 0; s < SCHED_SLOTS; ++s) {
      mbar_init(&sched_full[s], 1);
      mbar_init(&sched_empty[s], 5);  // CTA0: producer, MMA, epilogue; CTA1: producer, epilogue
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

  // every consumer walks the same ring of tile indices
  int sslot = 0;
  uint32_t sphase = 0;
  int seg_cur = 0;
  auto fetch_tile = [&]() -> int {
    mbar_wait_cluster_long(&sched_full[sslot], sphase);
    return static_cast<int>(sched_tile[sslot]);
  };
  auto release_tile = [&]() {  // call from ONE thread per consumer role, after every thread of the role has read the slot
    mbar_arrive_cluster(&sched_empty[sslot], 0);
  };
  auto advance_tile = [&]() {
    if (++sslot == SCHED_SLOTS) { sslot = 0; sphase ^= 1; }
  };

  if (warp_idx == 3) {
    // ===================== tile scheduler (leader CTA) =====================
    // Tiles are claimed in list order from one global counter, so the pairs stay balanced although P1 and P2 tiles cost
    // different amounts, and every dependency of a claimed tile was claimed earlier by a running pair (no deadlock,
    // whatever subset of the grid is resident).
    if (rank == 0 && lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      while (true) {
        mbar_wait_cluster_long(&sched_empty[slot], ph ^ 1);
        int t = static_cast<int>(atomicAdd(p.ready + 2 * p.num_m_blocks, 1u));
        if (t > p.num_tiles) t = p.num_tiles;
        sched_tile[slot] = static_cast<uint32_t>(t);
        st_shared_cluster_u32(const_cast<uint32_t*>(&sched_tile[slot]), 1, static_cast<uint32_t>(t));
        mbar_arrive_cluster(&sched_full[slot], 0);
        mbar_arrive_cluster(&sched_full[slot], 1);
        if (t >= p.num_tiles) break;
        if (++slot == SCHED_SLOTS) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      while (true) {
        const int tile = fetch_tile();
        release_tile();
        advance_tile();
        if (tile >= p.num_tiles) break;
        const FusedTile ft = fused_decode(p, seg, tile, seg_cur);
        const int m0 = ft.m_blk * 256 + static_cast<int>(rank) * 128;
        const CUtensorMap *ta, *tb;
        int brow, nk;
        if (ft.phase == 0) {
          const int n0 = ft.n_blk * OUT1;
          ta = &tmap_x;
          tb = (kSwiglu && rank == 1) ? &tmap_w1b : &tmap_w1a;  // SwiGLU: CTA0 gate rows, CTA1 up rows
          brow = kSwiglu ? n0 : n0 + static_cast<int>(rank) * 128;
          nk = p.k1_blocks;
        } else {
          ta = &tmap_mid_ld;
          tb = &tmap_w2;
          brow = ft.n_blk * 256 + static_cast<int>(rank) * 128;
          nk = p.k2_blocks;
          // every P1 tile of this 128-row half has landed in global memory
          const unsigned* flag = p.ready + ft.m_blk * 2 + rank;
          const unsigned need = static_cast<unsigned>(p.n1_blocks);
          if (m0 < p.M) {
            uint32_t spins = 0;
            while (ld_acquire_gpu(flag) < need) {
              __nanosleep(128);
              if (++spins > (1u << 24)) {
                printf("b200: fused_mlp dependency timeout block=%d m_blk=%d have=%u need=%u\n", blockIdx.x, ft.m_blk,
                       ld_acquire_gpu(flag), need);
                __trap();
              }
            }
          }
          fence_proxy_async_all();  // the acquired data is read through the async proxy (TMA) next
        }
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (P_A_STAGE_BYTES + P_B_STAGE_BYTES));
          tma_load_2d_2sm(smem + P_SMEM_A_OFF + stage * P_A_STAGE_BYTES, ta, &full_bar[stage], kb * BK, m0);
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
      while (true) {
        const int tile = fetch_tile();
        __syncwarp();
        if (lane == 0) release_tile();
        advance_tile();
        if (tile >= p.num_tiles) break;
        const FusedTile ft = fused_decode(p, seg, tile, seg_cur);
        const int nk = ft.phase == 0 ? p.k1_blocks : p.k2_blocks;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full_bar[stage], phase);  // (no back-off sleep here: a late poll is a tensor-pipe bubble)
          tc_fence_after();
          const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * (P_A_STAGE_BYTES >> 4));
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (P_B_STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_ss_2sm(d_tmem, ad + static_cast<uint64_t>(k * 2), bd + static_cast<uint64_t>(k * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(&tmem_full_bar[acc], 0x3);
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
    unsigned* pending_publish = nullptr;  // (thread 0) counter of the last P1 tile whose stores are still in flight
    while (true) {
      const int tile = fetch_tile();
      named_bar_sync(3, 128);  // all 128 threads have read the slot
      if (ep_tid == 0) release_tile();
      advance_tile();
      if (tile >= p.num_tiles) break;
      const FusedTile ft = fused_decode(p, seg, tile, seg_cur);
      const int m0 = ft.m_blk * 256 + static_cast<int>(rank) * 128;
      if (ft.phase == 1 && ep_tid == 0 && pending_publish != nullptr) {
        // this pair turns to a P2 tile, which may depend on the very P1 tile whose publication is still deferred (its
        // publication would otherwise wait for this tile's epilogue, which waits for this tile's loads: a cycle)
        tma_store_wait_all<0>();
        fence_proxy_async_all();
        red_release_gpu_add(pending_publish, 1u);  // release at gpu scope: orders the completed stores before the counter
        pending_publish = nullptr;
      }
      mbar_wait_long(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + lane_addr + static_cast<uint32_t>(acc * BN);
      if (ft.phase == 0) {
        pair_epilogue_tile<ACT, T>(smem + P_SMEM_C_OFF, cbuf, t_acc, &tmap_mid_st, p.bias_p1_0, p.bias_p1_1, p.N1, m0,
                                   ft.n_blk * OUT1, row, ep_tid, &tmem_empty_bar[acc], pending_publish, p.debug_flags);
        if (ep_tid == 0) pending_publish = p.ready + ft.m_blk * 2 + rank;  // published two chunks into the next tile
      } else {
        pair_epilogue_tile<B200_ACT_NONE, T>(smem + P_SMEM_C_OFF, cbuf, t_acc, &tmap_y, p.bias_p2, nullptr, p.N2, m0,
                                             ft.n_blk * 256, row, ep_tid, &tmem_empty_bar[acc], pending_publish, p.debug_flags);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (ep_tid == 0) {
      tma_store_wait_all<0>();
      if (pending_publish != nullptr) {
        fence_proxy_async_all();
        red_release_gpu_add(pending_publish, 1u);  // release at gpu scope: orders the completed stores before the counter
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int ACT, typename T>
int launch_fused(const CUtensorMap* maps, const FusedParams& p, cudaStream_t stream) {
  auto kern = fused_mlp_pair_kernel<ACT, T>;
  static bool attr_set[64] = {};
  constexpr int FUSED_SMEM = pair::P_SMEM_BYTES + static_cast<int>(sizeof(FusedSeg));
  static_assert(FUSED_SMEM <= 232448, "shared memory budget");
  B200_CUDA_OK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), FUSED_SMEM, attr_set));
  int max_ctas = sm_count();
  if (sm_limit() > 0 && sm_limit() < max_ctas) max_ctas = sm_limit();
  int pairs = max_ctas / 2;
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  if (pairs < 1) pairs = 1;
  // per-half P1 counters + the tile counter of the scheduler
  B200_CUDA_OK(cudaMemsetAsync(p.ready, 0, sizeof(unsigned) * (2 * p.num_m_blocks + 1), stream));
  kern<<<2 * pairs, NUM_THREADS, FUSED_SMEM, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], p);
  B200_CUDA_OK(cudaGetLastError());
  static const char* const names[5] = {"fused_mlp_pair_kernel<NONE>", "fused_mlp_pair_kernel<GELU_TANH>", "fused_mlp_pair_kernel<GELU_ERF>",
                                       "fused_mlp_pair_kernel<RELU>", "fused_mlp_pair_kernel<SWIGLU>"};
  note_launch(names[ACT], true);
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

template <typename T>
int dispatch_fused(int act, const CUtensorMap* maps, const FusedParams& p, cudaStream_t stream) {
  switch (act) {
    case B200_ACT_GELU_TANH: return launch_fused<B200_ACT_GELU_TANH, T>(maps, p, stream);
    case B200_ACT_GELU_ERF: return launch_fused<B200_ACT_GELU_ERF, T>(maps, p, stream);
    case B200_ACT_RELU: return launch_fused<B200_ACT_RELU, T>(maps, p, stream);
    case B200_ACT_SWIGLU: return launch_fused<B200_ACT_SWIGLU, T>(maps, p, stream);
    default: return set_error(B200_ERR_INVALID_ARGUMENT, "activation %d is not a FusedMLP activation", act);
  }
}

// Row-blocks (256 rows) per group and the P1 -> P2 lag (in groups) of the single-launch FusedMLP.
// The lag must cover the latency from claiming a P1 tile to publishing it (~3 tile times: loads, MMAs, epilogue, deferred
// publication), otherwise the P2 producers wait on the dependency counter: lag * tiles_per_group >= 3 waves.
// "Resident" mode: lag + 1 groups of intermediate (being written ... being consumed), the group's x panel and the weights
// that one wave touches fit the usable L2 budget — the intermediate is then read back from L2 (GPT-2-like widths).
// Otherwise ("streaming" mode, Llama widths: 5.6 MB of intermediate per row-block and 2-5.6 MB per weight slab) the raster
// falls back to the per-GEMM rule; the intermediate round-trips HBM but the MLP is still one launch.
inline void fused_schedule(int h, int i, int h_out, bool swiglu, int num_m_blocks, int* group_m, int* lag, bool* resident) {
  const double inter_blk = 256.0 * i * 2, slab_a1 = 256.0 * h * 2, slab_b = 256.0 * (h > i ? h : i) * 2;
  const double weights = (swiglu ? 2.0 : 1.0) * i * h * 2.0 + static_cast<double>(h_out) * i * 2.0;
  const double tpw = sm_count() / 2.0;
  const int n1 = (i + (swiglu ? 128 : 256) - 1) / (swiglu ? 128 : 256), n2 = (h_out + 255) / 256;
  auto lag_for = [&](int g) {
    int l = static_cast<int>((3.0 * tpw + g * (n1 + n2) - 1) / (g * (n1 + n2)));
    return l < 1 ? 1 : (l > 4 ? 4 : l);
  };
  int g = 0;
  const char* eg = getenv("B200_FUSED_G");
  const char* el = getenv("B200_FUSED_LAG");
  if (eg != nullptr && atoi(eg) > 0) {
    g = atoi(eg);
    *resident = false;
  } else if (gemm_group_rows() > 0) {
    g = gemm_group_rows() / 256;
    *resident = false;
  } else {
    for (int cand = 32; cand >= 2; cand /= 2) {
      double w = (tpw / cand) * slab_b;
      if (w > weights) w = weights;
      if ((lag_for(cand) + 1) * cand * inter_blk + cand * slab_a1 + w <= 56.0e6) { g = cand; break; }
    }
    *resident = g > 0;
    if (g == 0) g = choose_group_m(h > i ? h : i, 256, num_m_blocks);
  }
  if (g < 1) g = 1;
  if (g > num_m_blocks) g = num_m_blocks;
  *group_m = g;
  *lag = (el != nullptr && atoi(el) > 0) ? atoi(el) : lag_for(g);
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
static int choose_k_splits(int tiles, int kblocks, int64_t T, int n_blocks, int K) {
  const int sms = sm_count();
  if (const char* e = getenv("B200_GEMM_K_SPLITS")) {  // developer knob (tests/gemm_split_probe.py): force the split count
    const int s = atoi(e);
    if (s >= 1) return s <= kblocks ? s : kblocks;
  }
  if (tiles >= sms || kblocks < 8) return 1;
  // Skinny problems are weight-streaming bound and an SM can only pull so much: what matters is how many SMs stream at once.
  // Cost model per candidate: bytes moved (weights + the fp32 partials, weighted) divided by the fraction of SM-waves that
  // are full. The weight of the partials (3x their write + read bytes: they also cost the reduce launch and its latency)
  // is fitted to tests/gemm_split_probe.py (profiles/r2_gemm_splits.jsonl, CUDA-graph replay, weights streamed from HBM):
  // 86 SwiGLU tiles of a Llama up-projection want 1 split at T=64 (37.9 us; 3 splits 42.3 us: 17 MB of partials) but 5 at
  // T=8 (37.4 vs 44.8 us: 3.5 MB); 16 tiles of the down projection want 9 (25.4 vs 65 us), 48 tiles of a QKV projection 3.
  const double w_bytes = static_cast<double>(n_blocks) * gemm::BN * K * 2.0;
  const double part_bytes = 3.0 * 2.0 * static_cast<double>(T) * n_blocks * gemm::BN * 4.0;
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= 32 && kblocks / s >= 4; ++s) {
    const int ctas = tiles * s;
    const int waves = (ctas + sms - 1) / sms;
    const double eff = static_cast<double>(ctas) / (static_cast<double>(waves) * sms);
    const double cost = (w_bytes + (s > 1 ? s * part_bytes : 0.0)) / eff;
    if (cost < best_cost * 0.98) {  // prefer fewer splits unless the gain is real
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

int64_t linear_act_workspace_bytes(int64_t T, int K, int N, int act) {
  if (T <= 0 || K <= 0 || N <= 0) return 0;
  const int out_cols = (act == B200_ACT_SWIGLU) ? 128 : 256;
  const int64_t m_blocks = (T + gemm::BM - 1) / gemm::BM, n_blocks = (N + out_cols - 1) / out_cols;
  if (m_blocks * n_blocks > 0x7fffffff) return 0;
  const int splits = choose_k_splits(static_cast<int>(m_blocks * n_blocks), (K + gemm::BK - 1) / gemm::BK, T, static_cast<int>(n_blocks), K);
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
    p.k_splits = choose_k_splits(p.num_m_blocks * p.num_n_blocks, p.num_k_blocks, T, p.num_n_blocks, K);
    p.partial = static_cast<float*>(workspace);
  }
  // large problems: CTA-pair kernel (256-row tiles). B200_GEMM_PAIR=0 in the environment keeps the single-CTA kernel.
  static const bool pair_enabled = [] { const char* e = getenv("B200_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  p.use_pair = (pair_enabled && p.k_splits == 1 && T >= 1024) ? 1 : 0;
  if (p.use_pair) p.num_m_blocks = (p.M + 255) / 256;
  p.group_m = gemm::choose_group_m(K, p.use_pair ? 256 : gemm::BM, p.num_m_blocks);
  {
    // B200_GEMM_L2_HINTS = three digits for (A, B, C): 0 normal, 1 evict-first, 2 evict-last. Default "201".
    // Measured (tests/gemm_l2_probe.py under ncu, C3 up+gate GEMM, 4096-row groups): DRAM reads 3.57 GB with no hints,
    // 3.41 GB with 201, 7.2 GB with evict-first weights (a weight slab IS re-read by the other CTAs of its wave, with skew);
    // SM cycles are the same for all of them (5.05-5.09 M: the kernel is tensor-bound) but under the power cap less DRAM
    // traffic means a higher clock: 3.83 ms at 3.6 GB vs 4.20 ms at 10.3 GB (1024-row groups).
    static const uint64_t enc[3] = {kL2EvictNormal, kL2EvictFirst, kL2EvictLast};
    const char* e = getenv("B200_GEMM_L2_HINTS");
    int d[3] = {2, 0, 1};
    if (e != nullptr && e[0] && e[1] && e[2])
      for (int q = 0; q < 3; ++q) d[q] = (e[q] >= '0' && e[q] <= '2') ? e[q] - '0' : 0;
    p.hint_a = enc[d[0]]; p.hint_b = enc[d[1]]; p.hint_c = enc[d[2]];
  }
  p.kb_per_split = (p.num_k_blocks + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.num_tiles = p.num_m_blocks * p.num_n_blocks * p.k_splits;
  if (dtype == B200_DTYPE_BF16) return gemm::dispatch<__nv_bfloat16>(act, ta, tb0, tb1, tc, p, stream);
  return gemm::dispatch<__half>(act, ta, tb0, tb1, tc, p, stream);
}

int fused_mlp_single_launch(const void* x, int64_t ldx, const void* w_up, const void* b_up, const void* w_gate,
                            const void* b_gate, const void* w_down, const void* b_down, void* y, int64_t ldy, int64_t T,
                            int h, int i, int h_out, int act, int dtype, void* mid, void* flags, int64_t flag_bytes,
                            cudaStream_t stream) {
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_FP16, "dtype must be bf16 or fp16");
  B200_CHECK_ARG(T > 0 && T <= 0x7fffffffLL && h > 0 && i > 0 && h_out > 0, "bad sizes T=%lld h=%d i=%d", (long long)T, h, i);
  B200_CHECK_ARG(h % 8 == 0 && i % 8 == 0 && h_out % 8 == 0, "h, i and h_out must be multiples of 8");
  B200_CHECK_ARG(ldx >= h && ldx % 8 == 0 && ldy >= h_out && ldy % 8 == 0, "row strides must be >= width and multiples of 8");
  int rc;
  if ((rc = check_ptr16(x, "x")) || (rc = check_ptr16(w_up, "w_up")) || (rc = check_ptr16(w_down, "w_down")) ||
      (rc = check_ptr16(y, "y")) || (rc = check_ptr16(mid, "workspace")))
    return rc;
  const bool swiglu = (act == B200_ACT_SWIGLU);
  if (swiglu) {
    if ((rc = check_ptr16(w_gate, "gate weight"))) return rc;
  } else {
    B200_CHECK_ARG(w_gate == nullptr && b_gate == nullptr, "gate weight/bias given but activation is not SwiGLU");
  }
  for (const void* b : {b_up, b_gate, b_down})
    if (b != nullptr && (rc = check_ptr16(b, "bias"))) return rc;
  gemm::FusedParams p;
  p.M = static_cast<int>(T);
  p.N1 = i;
  p.N2 = h_out;
  p.bias_p1_0 = swiglu ? b_gate : b_up;
  p.bias_p1_1 = swiglu ? b_up : nullptr;
  p.bias_p2 = b_down;
  p.num_m_blocks = (p.M + 255) / 256;
  p.n1_blocks = (i + (swiglu ? 128 : 256) - 1) / (swiglu ? 128 : 256);
  p.n2_blocks = (h_out + 255) / 256;
  p.k1_blocks = (h + gemm::BK - 1) / gemm::BK;
  p.k2_blocks = (i + gemm::BK - 1) / gemm::BK;
  bool resident = false;
  gemm::fused_schedule(h, i, h_out, swiglu, p.num_m_blocks, &p.group_m, &p.lag, &resident);
  if ((p.num_m_blocks + p.group_m - 1) / p.group_m + p.lag > gemm::FUSED_MAX_STEPS)  // keep the segment table bounded
    p.group_m = (p.num_m_blocks + (gemm::FUSED_MAX_STEPS - p.lag) - 1) / (gemm::FUSED_MAX_STEPS - p.lag);
  p.num_groups = (p.num_m_blocks + p.group_m - 1) / p.group_m;
  { const char* e = getenv("B200_FUSED_DEBUG"); p.debug_flags = e ? atoi(e) : 0; }
  p.rows_last = p.num_m_blocks - (p.num_groups - 1) * p.group_m;
  const int64_t tiles = static_cast<int64_t>(p.num_m_blocks) * (p.n1_blocks + p.n2_blocks);
  B200_CHECK_ARG(tiles <= 0x7fffffffLL, "too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  const int64_t need = (static_cast<int64_t>(p.num_m_blocks) * 2 + 1) * static_cast<int64_t>(sizeof(unsigned));
  if (flags == nullptr || flag_bytes < need || (reinterpret_cast<uintptr_t>(flags) & 15) != 0)
    return set_error(B200_ERR_WORKSPACE, "fused_mlp workspace too small for the dependency counters");
  p.ready = static_cast<unsigned*>(flags);

  CUtensorMap maps[7];
  auto map2d = [&](CUtensorMap* m, const void* base, int64_t cols, int64_t rows, int64_t ld) {
    uint64_t dims[2] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
    uint32_t box[2] = {64, 128};
    return encode_tmap_sw128_16b(m, base, 2, dims, strides, box);
  };
  if ((rc = map2d(&maps[0], x, h, T, ldx))) return rc;                          // x            [T, h]
  if ((rc = map2d(&maps[1], swiglu ? w_gate : w_up, h, i, h))) return rc;      // gate (or up) [i, h]
  if ((rc = map2d(&maps[2], w_up, h, i, h))) return rc;                         // up           [i, h]
  if ((rc = map2d(&maps[3], mid, i, T, i))) return rc;                          // intermediate [T, i] (store)
  if ((rc = map2d(&maps[4], mid, i, T, i))) return rc;                          // intermediate [T, i] (load)
  if ((rc = map2d(&maps[5], w_down, i, h_out, i))) return rc;                   // down         [h_out, i]
  if ((rc = map2d(&maps[6], y, h_out, T, ldy))) return rc;                      // y            [T, h_out]
  if (dtype == B200_DTYPE_BF16) return gemm::dispatch_fused<__nv_bfloat16>(act, maps, p, stream);
  return gemm::dispatch_fused<__half>(act, maps, p, stream);
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
  const int64_t flags = align256((((T + 255) / 256) * 2 + 1) * static_cast<int64_t>(sizeof(unsigned)));  // fused kernel: counters
  const int64_t sk1 = b200::linear_act_workspace_bytes(T, h, i, B200_ACT_NONE);
  const int64_t sk2 = b200::linear_act_workspace_bytes(T, i, h, B200_ACT_NONE);
  if (sk1 > sk) sk = sk1;
  if (sk2 > sk) sk = sk2;
  if (flags > sk) sk = flags;
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
  const int64_t inter = align256(T * static_cast<int64_t>(i) * 2);
  void* sk_ws = static_cast<char*>(workspace) + inter;
  const int64_t sk_bytes = workspace_bytes - inter;
  // Prefill-sized inputs can run as ONE launch, up(+gate) and down projection interleaved by row groups
  // (fused_mlp_pair_kernel): B200_MLP_FUSED=1. Measured on B200 (tests/fused_mlp_probe.py, interleaved A/B): the single
  // launch keeps the intermediate in L2 at GPT-2 widths (DRAM reads 129 MB vs >= 260 MB, profiles/) and equals the two
  // launches in SM cycles (7.87M vs 7.72M at C3 under ncu), but in the power-capped sustained regime it runs 5-12 % slower
  // (C3 6.3-6.6 vs 5.9-6.0 ms, C2 0.274 vs 0.250 ms), so two launches stay the default.
  const char* fused_env = getenv("B200_MLP_FUSED");  // read per call: tests flip it inside one process
  const bool fused_enabled = fused_env != nullptr && fused_env[0] == '1';
  if (fused_enabled && T >= 1024) return b200::fused_mlp_single_launch(x, ldx, w_up, b_up, w_gate, b_gate, w_down, b_down, y, ldy, T, h, i, h_out, act, dtype, workspace, sk_ws, sk_bytes, s);
  // decode-sized inputs: GEMM1 + bias + activation (SwiGLU: gate/up pair) -> 16-bit intermediate -> GEMM2 + bias, each with
  // split-K over the SMs the few output tiles leave idle
  int rc = linear_act_impl(x, ldx, w_up, b_up, w_gate, b_gate, workspace, i, T, h, i, act, dtype, sk_ws, sk_bytes, s);
  if (rc) return rc;
  // GEMM2 + bias
  return linear_act_impl(workspace, i, w_down, b_down, nullptr, nullptr, y, ldy, T, i, h_out, B200_ACT_NONE, dtype, sk_ws,
                         sk_bytes, s);
}

}  // extern "C"
