// Small native runtime pieces that the reference ships as inline C++/CUDA strings:
//  * chunked batched copy kernel   — reference N2 (resharding/nvshmem_copy_service/kernels/chunked_kernel.cu): many (src, dst, bytes)
//    transfers in ONE launch; src/dst may be peer-mapped (symmetric-heap) addresses, so a weight reshard is a single kernel that
//    stores straight into the destination GPUs over NVLink instead of a stream of small NCCL send/recvs.
//  * managed-memory allocator hooks — reference N5 (inference/unified_memory.py): cudaMallocManaged-backed pluggable allocator so
//    KV caches may oversubscribe HBM (pages migrate over NVLink-C2C / PCIe on demand).
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb200 {

struct CopyTask {
  const void* src;
  void* dst;
  unsigned long long bytes;
};

constexpr unsigned long long COPY_CHUNK = 64 * 1024;

// one CTA per (task, chunk); 16-byte vectors when both pointers allow it, bytes otherwise
__global__ void batched_copy_kernel(const CopyTask* __restrict__ tasks, const unsigned long long* __restrict__ chunk_prefix, int ntasks, unsigned long long total_chunks) {
  for (unsigned long long g = blockIdx.x; g < total_chunks; g += gridDim.x) {
    // binary search the task owning global chunk g
    int lo = 0, hi = ntasks - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (chunk_prefix[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const CopyTask t = tasks[lo];
    const unsigned long long off = (g - chunk_prefix[lo]) * COPY_CHUNK;
    const unsigned long long n = min(COPY_CHUNK, t.bytes - off);
    const char* s = reinterpret_cast<const char*>(t.src) + off;
    char* d = reinterpret_cast<char*>(t.dst) + off;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
      const unsigned long long nv = n / 16;
      const uint4* sv = reinterpret_cast<const uint4*>(s);
      uint4* dv = reinterpret_cast<uint4*>(d);
      for (unsigned long long i = threadIdx.x; i < nv; i += blockDim.x) dv[i] = sv[i];
      for (unsigned long long i = nv * 16 + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    } else {
      for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
  }
}

}  // namespace mb200

extern "C" void mb200_batched_copy(const void* tasks_dev, const void* chunk_prefix_dev, int ntasks, unsigned long long total_chunks, int nblocks, cudaStream_t s) {
  if (ntasks == 0 || total_chunks == 0) return;
  const unsigned long long grid = total_chunks < (unsigned long long)nblocks ? total_chunks : (unsigned long long)nblocks;
  mb200::batched_copy_kernel<<<(unsigned)grid, 256, 0, s>>>(reinterpret_cast<const mb200::CopyTask*>(tasks_dev),
                                                           reinterpret_cast<const unsigned long long*>(chunk_prefix_dev), ntasks, total_chunks);
}

// ---- torch.cuda.memory.CUDAPluggableAllocator entry points ---------------------------------------------------------------
extern "C" void* mb200_managed_malloc(size_t size, int device, cudaStream_t stream) {
  void* p = nullptr;
  if (size == 0) return nullptr;
  if (cudaMallocManaged(&p, size, cudaMemAttachGlobal) != cudaSuccess) return nullptr;
  cudaMemLocation loc;
  loc.type = cudaMemLocationTypeDevice;
  loc.id = device;
  cudaMemAdvise(p, size, cudaMemAdviseSetPreferredLocation, loc);   // stay in HBM while it fits, spill to host otherwise
  cudaMemAdvise(p, size, cudaMemAdviseSetAccessedBy, loc);
  return p;
}

extern "C" void mb200_managed_free(void* ptr, size_t size, int device, cudaStream_t stream) {
  if (ptr != nullptr) cudaFree(ptr);
}
