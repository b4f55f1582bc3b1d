// NVLink 5 / NVSwitch collectives over symmetric (peer-mapped) memory for one 8xB200 box.
//
// Every rank maps every peer's buffer (P2P pointers) and, when NVLS is available, one multicast
// address that aliases all replicas:
//   * all-gather      : each rank `multimem.st`s its shard once; NVSwitch replicates it to all peers
//   * reduce-scatter  : each rank `multimem.ld_reduce`s its shard; NVSwitch adds the replicas in fp32
//   * all-reduce      : ld_reduce own slice + multimem.st the result (two-shot in ONE kernel)
// Cross-GPU ordering: epoch barrier on per-rank flag arrays (st.release.sys / ld.acquire.sys).  Epochs
// only grow, so flags are never reset; every rank issues the same collective sequence per slot, so
// the host passes the epoch as an argument.  Independent slots (own flags + ctrl block) let collectives
// on different streams run concurrently.  P2P fallbacks cover boxes without multicast support.
// Reference analogue: inference/communication/torch_symm_triton/{multimem_asm,barrier,collectives}.py.
#include "common.cuh"

namespace mb200 {

constexpr int MAX_RANKS = 16;

struct Peers {
  void* ptr[MAX_RANKS];        // peer-mapped data buffers (same layout on every rank)
  uint32_t* flags[MAX_RANKS];  // peer-mapped flag arrays [slots][MAX_RANKS]; entry [slot][src] is written by rank src
};
struct Ctrl {
  uint32_t arrive, release, pad0, pad1;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// One warp: tell every peer "I reached `epoch`" and wait until every peer did.
__device__ __forceinline__ void cross_gpu_barrier_warp(const Peers& pr, int rank, int world, uint32_t epoch, int slot) {
  const int lane = threadIdx.x & 31;
  __threadfence_system();
  if (lane < world) {
    st_release_sys(pr.flags[lane] + slot * MAX_RANKS + rank, epoch);
    const uint32_t* mine = pr.flags[rank] + slot * MAX_RANKS + lane;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncwarp();
}

// Leading barrier: block 0 synchronises with the peers, all other blocks wait for its local release.
__device__ __forceinline__ void leading_barrier(const Peers& pr, int rank, int world, uint32_t epoch, int slot, Ctrl* ctrl) {
  if (blockIdx.x == 0) {
    if (threadIdx.x < 32) {
      cross_gpu_barrier_warp(pr, rank, world, epoch, slot);
      if (threadIdx.x == 0) st_release_gpu(&ctrl->release, epoch);
    }
  } else if (threadIdx.x == 0) {
    while ((int32_t)(ld_acquire_gpu(&ctrl->release) - epoch) < 0) {
    }
  }
  __syncthreads();
}

// Trailing barrier: the last block to finish its part synchronises with the peers.
__device__ __forceinline__ void trailing_barrier(const Peers& pr, int rank, int world, uint32_t epoch, int slot, Ctrl* ctrl, bool cross_gpu) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&ctrl->arrive, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < 32) {
    if (threadIdx.x == 0) ctrl->arrive = 0;
    if (cross_gpu) cross_gpu_barrier_warp(pr, rank, world, epoch, slot);
  }
}

__device__ __forceinline__ void multimem_st_v4(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ uint4 multimem_ld_reduce_f32x4(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------------
// all-gather: local `src` shard → bytes [dst_off + rank*shard, …) of EVERY rank's symmetric buffer.
// Trailing barrier (epoch) ⇒ when the kernel retires, all shards have landed locally.
__global__ void __launch_bounds__(512) allgather_kernel(Peers pr, void* mc_base, const uint4* __restrict__ src, size_t dst_off_bytes, size_t shard_bytes,
                                                          int rank, int world, uint32_t epoch, Ctrl* ctrl, int slot) {
  const size_t n16 = shard_bytes / 16;
  const size_t off16 = (dst_off_bytes + (size_t)rank * shard_bytes) / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (mc_base != nullptr) {
    uint4* mc = reinterpret_cast<uint4*>(mc_base) + off16;
    // 8 independent loads in flight per thread before the first (fire-and-forget) multicast store
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) multimem_st_v4(mc + i + u * stride, v[u]);
    }
    for (; i < n16; i += stride) multimem_st_v4(mc + i, __ldg(src + i));
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) {
      const uint4 v = src[i];
#pragma unroll 1
      for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(pr.ptr[(rank + p) % world])[off16 + i] = v;
    }
  }
  trailing_barrier(pr, rank, world, epoch, slot, ctrl, true);
}

// reduce-scatter of a symmetric input: out[i] = scale * sum_r in_r[src_off + rank*shard + i].
// Leading barrier (epoch): all ranks' inputs are complete.  Optional trailing barrier (epoch+1) for
// in-place use (nobody may overwrite its input while a peer is still reading it).
template <typename T>
__global__ void __launch_bounds__(512) reducescatter_kernel(Peers pr, const void* mc_base, size_t src_off_bytes, T* __restrict__ out, size_t shard_elems, float scale,
                                                              int rank, int world, uint32_t epoch, Ctrl* ctrl, int slot, int trailing) {
  leading_barrier(pr, rank, world, epoch, slot, ctrl);
  constexpr int VN = 16 / sizeof(T);
  const size_t n16 = shard_elems / VN;
  const size_t off16 = src_off_bytes / 16 + (size_t)rank * n16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (mc_base != nullptr) {
    // in-switch reduction: the round trip is several microseconds, so keep 8 ld_reduce per thread in flight
    const uint4* mc = reinterpret_cast<const uint4*>(mc_base) + off16;
    auto finish = [&](size_t i, const uint4& v) {
      const T* h = reinterpret_cast<const T*>(&v);
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = from_f<T>(to_f(h[k]) * scale);
      st16(out + i * VN, o);
    };
    for (; i0 + 7 * stride < n16; i0 += 8 * stride) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = sizeof(T) == 2 ? multimem_ld_reduce_bf16x8(mc + i0 + u * stride) : multimem_ld_reduce_f32x4(mc + i0 + u * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) finish(i0 + u * stride, v[u]);
    }
    for (; i0 < n16; i0 += stride) finish(i0, sizeof(T) == 2 ? multimem_ld_reduce_bf16x8(mc + i0) : multimem_ld_reduce_f32x4(mc + i0));
  } else {
    for (size_t i = i0; i < n16; i += stride) {
      float acc[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] = 0.f;
#pragma unroll 1
      for (int p = 0; p < world; ++p) {
        const uint4 v = reinterpret_cast<const uint4*>(pr.ptr[(rank + p) % world])[off16 + i];
        const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[k] += to_f(h[k]);
      }
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = from_f<T>(acc[k] * scale);
      st16(out + i * VN, o);
    }
  }
  trailing_barrier(pr, rank, world, epoch + 1, slot, ctrl, trailing != 0);
}

// all-reduce in place on a symmetric buffer: rank r reduces slice r and broadcasts it (epochs e, e+1).
template <typename T>
__global__ void __launch_bounds__(512) allreduce_kernel(Peers pr, void* mc_base, size_t off_bytes, size_t elems, float scale, int rank, int world, uint32_t epoch,
                                                          Ctrl* ctrl, int slot) {
  leading_barrier(pr, rank, world, epoch, slot, ctrl);
  constexpr int VN = 16 / sizeof(T);
  const size_t n16 = elems / VN;
  const size_t per = (n16 + world - 1) / world;
  const size_t lo = (size_t)rank * per, hi = min(n16, lo + per);
  const size_t base16 = off_bytes / 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hi; i += stride) {
    float acc[VN];
    if (mc_base != nullptr) {
      const uint4 v = sizeof(T) == 2 ? multimem_ld_reduce_bf16x8(reinterpret_cast<const uint4*>(mc_base) + base16 + i)
                                     : multimem_ld_reduce_f32x4(reinterpret_cast<const uint4*>(mc_base) + base16 + i);
      const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] = to_f(h[k]) * scale;
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] = 0.f;
      for (int p = 0; p < world; ++p) {
        const uint4 v = reinterpret_cast<const uint4*>(pr.ptr[(rank + p) % world])[base16 + i];
        const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[k] += to_f(h[k]);
      }
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] *= scale;
    }
    Vec<T> o;
#pragma unroll
    for (int k = 0; k < VN; ++k) o.v[k] = from_f<T>(acc[k]);
    const uint4 ov = *reinterpret_cast<const uint4*>(&o);
    if (mc_base != nullptr) {
      multimem_st_v4(reinterpret_cast<uint4*>(mc_base) + base16 + i, ov);
    } else {
      for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(pr.ptr[(rank + p) % world])[base16 + i] = ov;
    }
  }
  trailing_barrier(pr, rank, world, epoch + 1, slot, ctrl, true);
}

__global__ void barrier_kernel(Peers pr, int rank, int world, uint32_t epoch, int slot) { cross_gpu_barrier_warp(pr, rank, world, epoch, slot); }

static Peers make_peers(const int64_t* ptrs, const int64_t* flags, int world) {
  Peers p;
  for (int i = 0; i < MAX_RANKS; ++i) {
    p.ptr[i] = i < world ? reinterpret_cast<void*>(ptrs[i]) : nullptr;
    p.flags[i] = i < world ? reinterpret_cast<uint32_t*>(flags[i]) : nullptr;
  }
  return p;
}

}  // namespace mb200

using namespace mb200;

extern "C" void mb200_nvl_barrier(const int64_t* ptrs, const int64_t* flags, int rank, int world, uint32_t epoch, int slot, cudaStream_t s) {
  barrier_kernel<<<1, 32, 0, s>>>(make_peers(ptrs, flags, world), rank, world, epoch, slot);
}
extern "C" void mb200_nvl_allgather(const int64_t* ptrs, const int64_t* flags, int64_t mc, const void* src, size_t dst_off, size_t shard_bytes, int rank,
                                    int world, uint32_t epoch, void* ctrl, int slot, int nblocks, cudaStream_t s) {
  allgather_kernel<<<nblocks, 512, 0, s>>>(make_peers(ptrs, flags, world), reinterpret_cast<void*>(mc), (const uint4*)src, dst_off, shard_bytes, rank, world,
                                           epoch, (Ctrl*)ctrl + slot, slot);
}
extern "C" void mb200_nvl_reducescatter(const int64_t* ptrs, const int64_t* flags, int64_t mc, size_t src_off, void* out, size_t shard_elems, float scale,
                                        int dtype, int rank, int world, uint32_t epoch, void* ctrl, int slot, int trailing, int nblocks, cudaStream_t s) {
  Peers p = make_peers(ptrs, flags, world);
  if (dtype == kBF16)
    reducescatter_kernel<__nv_bfloat16><<<nblocks, 512, 0, s>>>(p, reinterpret_cast<void*>(mc), src_off, (__nv_bfloat16*)out, shard_elems, scale, rank, world,
                                                               epoch, (Ctrl*)ctrl + slot, slot, trailing);
  else
    reducescatter_kernel<float><<<nblocks, 512, 0, s>>>(p, reinterpret_cast<void*>(mc), src_off, (float*)out, shard_elems, scale, rank, world, epoch,
                                                       (Ctrl*)ctrl + slot, slot, trailing);
}
extern "C" void mb200_nvl_allreduce(const int64_t* ptrs, const int64_t* flags, int64_t mc, size_t off, size_t elems, float scale, int dtype, int rank, int world,
                                    uint32_t epoch, void* ctrl, int slot, int nblocks, cudaStream_t s) {
  Peers p = make_peers(ptrs, flags, world);
  if (dtype == kBF16)
    allreduce_kernel<__nv_bfloat16><<<nblocks, 512, 0, s>>>(p, reinterpret_cast<void*>(mc), off, elems, scale, rank, world, epoch, (Ctrl*)ctrl + slot, slot);
  else
    allreduce_kernel<float><<<nblocks, 512, 0, s>>>(p, reinterpret_cast<void*>(mc), off, elems, scale, rank, world, epoch, (Ctrl*)ctrl + slot, slot);
}
