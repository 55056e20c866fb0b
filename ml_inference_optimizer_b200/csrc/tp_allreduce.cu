// K6 — the tensor-parallel all-reduce over NVSwitch, written for the row-parallel down projection of the FusedMLP.
//
// Replaces `torch.distributed.all_reduce(output_parallel)` of RowParallelLinear.forward
// (reference parallelism/tensor_parallel.py:296-302) and the bias add that follows it (:304-308).
//
// Every rank owns a SYMMETRIC buffer (same size and layout on all ranks) that is mapped into every process: once as
// one unicast address per peer and, when the fabric supports it, once more as a MULTICAST address that names the same
// offset on all ranks at the same time. The down projection writes its partial [T, h] output straight into that buffer.
// This kernel then performs a two-shot all-reduce in place:
//   rank r owns slice r of the region;  multimem.ld_reduce  sums the slice over all ranks INSIDE the switch (fp32
//   accumulation of the bf16/fp16 partials, one 16-byte vector per instruction), the optional bias is added once, and
//   multimem.st  broadcasts the reduced vector back to all ranks through the switch.
// Per GPU that is payload/world bytes in, payload/world out on the reduce side and the same on the broadcast side — the
// minimum NVLink traffic for an all-reduce — and the reduction happens in the switch, so the CTAs only issue loads and
// stores: they are built to share SMs with the GEMMs of the next token chunk (see THREADS below). Without multicast support the same kernel reads the peers' slices with
// plain loads over their unicast mappings, sums in fp32 and stores to every peer.
//
// Synchronisation: per-CTA flag slots inside the symmetric buffer (`flags[cta][src_rank]`, monotonically increasing
// epochs, st.release.sys / ld.acquire.sys): CTA b of every rank waits for CTA b of all ranks before it reads (partials
// complete) and after it wrote (result complete everywhere). Every spin is bounded; on a timeout the kernel raises
// *error_flag and leaves, so a lost peer cannot hang the GPU.
#include "common.cuh"
#include "host_common.h"

namespace b200 {
namespace ar {

constexpr int MAX_WORLD = 8;
// Small CTAs with few registers and no shared memory: a CTA of this kernel fits on an SM NEXT TO a resident CTA of the
// persistent GEMM (209 regs x 256 threads + 227 KB smem leave ~12K registers and the 1 KB CTA reservation free), so the
// all-reduce of token chunk c runs underneath the GEMMs of chunk c+1 without taking SMs away from them. Its warps mostly
// wait on NVLink round trips; bytes in flight = CTAs x 256 threads x 4 x 16 B (2.4 MB at one CTA per SM).
constexpr int MAX_CTAS = 160;
constexpr int THREADS = 256;
constexpr int UNROLL = 4;
constexpr int MIN_CTAS_PER_SM = 6;  // caps the kernel at 42 registers per thread
constexpr long long SPIN_TIMEOUT_CYCLES = 6000000000LL;  // ~3 s at 2 GHz

struct Params {
  char* mc_base;                 // multicast mapping of the symmetric buffer (nullptr: unicast path)
  char* peer_base[MAX_WORLD];    // unicast mapping of every rank's buffer in this process
  long long data_off, nbytes;    // region to reduce (16-byte aligned, multiple of 16 bytes)
  long long flag_off;            // offset of the flag block (MAX_CTAS x MAX_WORLD x uint32) in the symmetric buffer
  int rank, world;
  unsigned epoch;                // this call uses epoch (entry barrier) and epoch + 1 (exit barrier)
  const void* bias;              // optional [ncols] bias added to every reduced row
  int ncols;                     // row width in elements (only used with bias)
  int* error_flag;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// CTA b of this rank meets CTA b of every other rank
__device__ __forceinline__ void cross_rank_barrier(const Params& p, unsigned value) {
  __syncthreads();
  if (threadIdx.x < static_cast<unsigned>(p.world)) {
    unsigned* remote = reinterpret_cast<unsigned*>(p.peer_base[threadIdx.x] + p.flag_off) + blockIdx.x * MAX_WORLD + p.rank;
    st_release_sys(remote, value);
    const unsigned* mine = reinterpret_cast<const unsigned*>(p.peer_base[p.rank] + p.flag_off) + blockIdx.x * MAX_WORLD + threadIdx.x;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(mine) - value) < 0) {
      if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) {
        atomicExch(p.error_flag, 1);
        break;
      }
    }
  }
  __syncthreads();
}

template <typename T>
__device__ __forceinline__ uint4 mc_ld_reduce(const char* addr);
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const char* addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
  return v;
}
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<__half>(const char* addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(char* addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_peer(const char* addr) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(char* addr, const uint4& v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T>
__device__ __forceinline__ void add8(float (&acc)[8], const uint4& v) {
  float2 f;
  f = Pack2<T>::unpack(v.x); acc[0] += f.x; acc[1] += f.y;
  f = Pack2<T>::unpack(v.y); acc[2] += f.x; acc[3] += f.y;
  f = Pack2<T>::unpack(v.z); acc[4] += f.x; acc[5] += f.y;
  f = Pack2<T>::unpack(v.w); acc[6] += f.x; acc[7] += f.y;
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  uint4 v;
  v.x = Pack2<T>::pack(a[0], a[1]); v.y = Pack2<T>::pack(a[2], a[3]);
  v.z = Pack2<T>::pack(a[4], a[5]); v.w = Pack2<T>::pack(a[6], a[7]);
  return v;
}

template <typename T, bool MULTICAST>
__global__ void __launch_bounds__(THREADS, MIN_CTAS_PER_SM) tp_allreduce_kernel(const Params p) {
  cross_rank_barrier(p, p.epoch);  // every rank's partial rows are complete and visible

  const long long total_vec = p.nbytes >> 4;
  const long long per_rank = (total_vec + p.world - 1) / p.world;
  const long long v0 = per_rank * p.rank;
  long long v1 = v0 + per_rank;
  if (v1 > total_vec) v1 = total_vec;
  const long long stride = static_cast<long long>(gridDim.x) * THREADS;
  const int vec_per_row = p.ncols >> 3;

  for (long long base = v0 + static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; base < v1; base += stride * UNROLL) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * stride;
      if (i < v1) {
        const long long off = p.data_off + (i << 4);
        if constexpr (MULTICAST) {
          v[u] = mc_ld_reduce<T>(p.mc_base + off);
        } else {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int r = 0; r < p.world; ++r) add8<T>(acc, ld_peer(p.peer_base[r] + off));
          v[u] = pack8<T>(acc);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * stride;
      if (i < v1) {
        const long long off = p.data_off + (i << 4);
        if (p.bias != nullptr) {
          const int col = static_cast<int>(i % vec_per_row) << 3;
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          add8<T>(acc, v[u]);
          add8<T>(acc, __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(p.bias) + col)));
          v[u] = pack8<T>(acc);
        }
        if constexpr (MULTICAST) {
          mc_st(p.mc_base + off, v[u]);
        } else {
          for (int r = 0; r < p.world; ++r) st_peer(p.peer_base[(p.rank + r) % p.world] + off, v[u]);
        }
      }
    }
  }
  __threadfence_system();
  cross_rank_barrier(p, p.epoch + 1);  // every rank's slice has landed everywhere
}

}  // namespace ar
}  // namespace b200

extern "C" {

int64_t b200_tp_allreduce_flag_bytes(void) { return b200::ar::MAX_CTAS * b200::ar::MAX_WORLD * static_cast<int64_t>(sizeof(unsigned)); }

int b200_tp_allreduce(void* multicast_base, void* const* peer_bases, int world, int rank, int64_t data_offset, int64_t nbytes,
                      int64_t flag_offset, uint32_t epoch, const void* bias, int ncols, int dtype, int max_ctas,
                      int* error_flag, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(world >= 2 && world <= ar::MAX_WORLD && rank >= 0 && rank < world, "tp_allreduce: bad world %d / rank %d", world, rank);
  B200_CHECK_ARG(peer_bases != nullptr && error_flag != nullptr, "tp_allreduce: NULL pointer argument");
  B200_CHECK_ARG(nbytes >= 0 && nbytes % 16 == 0 && data_offset % 16 == 0 && flag_offset % 16 == 0,
                 "tp_allreduce: offsets and length must be multiples of 16 bytes");
  B200_CHECK_ARG(dtype == B200_DTYPE_BF16 || dtype == B200_DTYPE_FP16, "tp_allreduce: dtype must be bf16 or fp16");
  B200_CHECK_ARG(bias == nullptr || (ncols > 0 && ncols % 8 == 0 && (nbytes / 2) % ncols == 0 && (data_offset / 2) % ncols == 0),
                 "tp_allreduce: with a bias the region must consist of whole rows of ncols (multiple of 8) elements");
  if (nbytes == 0) return B200_OK;
  ar::Params p;
  p.mc_base = static_cast<char*>(multicast_base);
  for (int r = 0; r < ar::MAX_WORLD; ++r) p.peer_base[r] = r < world ? static_cast<char*>(peer_bases[r]) : nullptr;
  for (int r = 0; r < world; ++r) B200_CHECK_ARG(p.peer_base[r] != nullptr, "tp_allreduce: peer %d has no mapping", r);
  p.data_off = data_offset;
  p.nbytes = nbytes;
  p.flag_off = flag_offset;
  p.rank = rank;
  p.world = world;
  p.epoch = epoch;
  p.bias = bias;
  p.ncols = bias != nullptr ? ncols : 8;
  p.error_flag = error_flag;
  int ctas = max_ctas > 0 ? max_ctas : sm_count();
  if (ctas > ar::MAX_CTAS) ctas = ar::MAX_CTAS;
  const long long per_rank_vec = ((nbytes >> 4) + world - 1) / world;
  const long long need = (per_rank_vec + ar::THREADS * ar::UNROLL - 1) / (ar::THREADS * ar::UNROLL);
  if (need < ctas) ctas = need < 1 ? 1 : static_cast<int>(need);
  // NOTE: every rank must launch the same number of CTAs (the flag slots are per CTA): ctas depends only on arguments
  // that are identical on all ranks.
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool mc = multicast_base != nullptr;
  if (dtype == B200_DTYPE_BF16) {
    if (mc) ar::tp_allreduce_kernel<__nv_bfloat16, true><<<ctas, ar::THREADS, 0, s>>>(p);
    else ar::tp_allreduce_kernel<__nv_bfloat16, false><<<ctas, ar::THREADS, 0, s>>>(p);
  } else {
    if (mc) ar::tp_allreduce_kernel<__half, true><<<ctas, ar::THREADS, 0, s>>>(p);
    else ar::tp_allreduce_kernel<__half, false><<<ctas, ar::THREADS, 0, s>>>(p);
  }
  B200_CUDA_OK(cudaGetLastError());
  note_launch(mc ? "tp_allreduce_kernel<multicast>" : "tp_allreduce_kernel<peer>");
  return B200_OK;
}

}  // extern "C"
