// Multi-tensor optimizer kernels for sm_100a: l2-norm, scale, fused Adam(W).
//
// A launch covers an arbitrary list of tensors.  Work is cut into CHUNK-element pieces;
// `chunk_prefix[i]` = number of chunks before tensor i, so a CTA maps a global chunk id to
// (tensor, offset) with a binary search over <= a few thousand entries held in L2.
// One persistent grid (multiple of 148 CTAs) walks all chunks — a single launch updates the
// fp32 master, both moments and the bf16 model copy of every parameter of a param group
// (reference: TE/Apex multi_tensor_applier + FusedAdam).
#include "common.cuh"

namespace mb200 {

constexpr int CHUNK = 8192;  // elements per work item

struct Meta {
  const long* chunk_prefix;  // [n+1]
  const long* sizes;         // [n]
  int n;
};

__device__ __forceinline__ int find_tensor(const long* prefix, int n, long chunk) {
  int lo = 0, hi = n;  // prefix[lo] <= chunk < prefix[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float load_as_float(const void* p, int dtype, long i) {
  if (dtype == kF32) return ((const float*)p)[i];
  if (dtype == kBF16) return __bfloat162float(((const __nv_bfloat16*)p)[i]);
  return __half2float(((const __half*)p)[i]);
}
__device__ __forceinline__ void store_from_float(void* p, int dtype, long i, float v) {
  if (dtype == kF32) ((float*)p)[i] = v;
  else if (dtype == kBF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((__half*)p)[i] = __float2half_rn(v);
}

__global__ void __launch_bounds__(512) l2norm_kernel(const void* const* __restrict__ ptrs, const long* __restrict__ sizes, const int* __restrict__ dtypes,
                                                       const long* __restrict__ prefix, int n, float* __restrict__ partial) {
  __shared__ float red[32];
  const long total_chunks = prefix[n];
  float acc = 0.f;
  for (long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int t = find_tensor(prefix, n, c);
    const long off = (c - prefix[t]) * CHUNK;
    const long len = min((long)CHUNK, sizes[t] - off);
    const int dt = dtypes[t];
    const char* base = (const char*)ptrs[t];
    if (dt == kF32) {
      const float* p = (const float*)base + off;
      if ((((uintptr_t)p) & 15) == 0) {
        for (long i = threadIdx.x * 4L; i + 3 < len; i += blockDim.x * 4L) {
          const float4 v = *reinterpret_cast<const float4*>(p + i);
          acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (long i = (len & ~3L) + threadIdx.x; i < len; i += blockDim.x) acc += p[i] * p[i];
      } else {
        for (long i = threadIdx.x; i < len; i += blockDim.x) acc += p[i] * p[i];
      }
    } else {
      const __nv_bfloat16* p = (const __nv_bfloat16*)base + off;  // same width as half
      if ((((uintptr_t)p) & 15) == 0) {
        for (long i = threadIdx.x * 8L; i + 7 < len; i += blockDim.x * 8L) {
          if (dt == kBF16) {
            Vec<__nv_bfloat16> v = ld16(p + i);
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float f = to_f(v.v[k]); acc += f * f; }
          } else {
            Vec<__half> v = ld16((const __half*)p + i);
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float f = to_f(v.v[k]); acc += f * f; }
          }
        }
        for (long i = (len & ~7L) + threadIdx.x; i < len; i += blockDim.x) { const float f = load_as_float(p, dt, i); acc += f * f; }
      } else {
        for (long i = threadIdx.x; i < len; i += blockDim.x) { const float f = load_as_float(p, dt, i); acc += f * f; }
      }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void l2norm_final_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) out[0] = sqrtf(a);
}

__global__ void __launch_bounds__(512) scale_kernel(void* const* __restrict__ ptrs, const long* __restrict__ sizes, const int* __restrict__ dtypes,
                                                      const long* __restrict__ prefix, int n, const float* __restrict__ scale) {
  const long total_chunks = prefix[n];
  const float sc = scale[0];
  for (long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int t = find_tensor(prefix, n, c);
    const long off = (c - prefix[t]) * CHUNK;
    const long len = min((long)CHUNK, sizes[t] - off);
    const int dt = dtypes[t];
    for (long i = threadIdx.x; i < len; i += blockDim.x) store_from_float(ptrs[t], dt, off + i, load_as_float(ptrs[t], dt, off + i) * sc);
  }
}

// Adam(W): g' = g * grad_scale; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2;
//          p -= lr * ( (m/bc1) / (sqrt(v/bc2) + eps) + wd * p )        [adamw]
// L2 mode (adamw=0) adds wd*p to g' first.  lowp (bf16/fp16) copy written in the same pass.
template <typename G, typename L>
__device__ __forceinline__ void adam_chunk(float* __restrict__ p, const G* __restrict__ g, float* __restrict__ m, float* __restrict__ v, L* __restrict__ lp,
                                           long len, float lr, float b1, float b2, float eps, float wd, float rbc1, float rbc2, int adamw, float gs) {
  // 4 elements per thread per iteration (fp32 state vectors are 16 B)
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0 && ((((uintptr_t)g) & (sizeof(G) * 4 - 1)) == 0) &&
                       (lp == nullptr || (((uintptr_t)lp) & (sizeof(L) * 4 - 1)) == 0);
  long i0 = 0;
  if (aligned) {
    for (long i = threadIdx.x * 4L; i + 3 < len; i += blockDim.x * 4L) {
      float4 pv = *reinterpret_cast<const float4*>(p + i), mv = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
      float gv[4];
      if (sizeof(G) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(g + i);
        gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(g + i);
        const G* h = reinterpret_cast<const G*>(&t);
#pragma unroll
        for (int k = 0; k < 4; ++k) gv[k] = to_f(h[k]);
      }
      float pp[4] = {pv.x, pv.y, pv.z, pv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gg = gv[k] * gs;
        if (!adamw) gg += wd * pp[k];
        mm[k] = b1 * mm[k] + (1.f - b1) * gg;
        vq[k] = b2 * vq[k] + (1.f - b2) * gg * gg;
        float upd = (mm[k] * rbc1) / (sqrtf(vq[k] * rbc2) + eps);
        if (adamw) upd += wd * pp[k];
        pp[k] -= lr * upd;
      }
      *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vq[0], vq[1], vq[2], vq[3]);
      if (lp != nullptr) {
        if (sizeof(L) == 4) {
          *reinterpret_cast<float4*>(lp + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        } else {
          L h[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) h[k] = from_f<L>(pp[k]);
          *reinterpret_cast<uint2*>(lp + i) = *reinterpret_cast<const uint2*>(h);
        }
      }
    }
    i0 = len & ~3L;
  }
  for (long i = i0 + threadIdx.x; i < len; i += blockDim.x) {
    float gg = to_f(g[i]) * gs, pp = p[i];
    if (!adamw) gg += wd * pp;
    const float mm = b1 * m[i] + (1.f - b1) * gg, vq = b2 * v[i] + (1.f - b2) * gg * gg;
    float upd = (mm * rbc1) / (sqrtf(vq * rbc2) + eps);
    if (adamw) upd += wd * pp;
    pp -= lr * upd;
    p[i] = pp; m[i] = mm; v[i] = vq;
    if (lp != nullptr) lp[i] = from_f<L>(pp);
  }
}

__global__ void __launch_bounds__(512) adam_kernel(float* const* __restrict__ p32, const void* const* __restrict__ grads, float* const* __restrict__ m,
                                                     float* const* __restrict__ v, void* const* __restrict__ lowp, const long* __restrict__ sizes,
                                                     const int* __restrict__ gdt, const int* __restrict__ ldt, const long* __restrict__ prefix, int n,
                                                     float lr, float b1, float b2, float eps, float wd, float rbc1, float rbc2, int adamw,
                                                     const float* __restrict__ grad_scale) {
  const long total_chunks = prefix[n];
  const float gs = grad_scale ? grad_scale[0] : 1.f;
  for (long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int t = find_tensor(prefix, n, c);
    const long off = (c - prefix[t]) * CHUNK;
    const long len = min((long)CHUNK, sizes[t] - off);
    float* pp = p32[t] + off;
    float* mm = m[t] + off;
    float* vv = v[t] + off;
    const int gd = gdt[t], ld = ldt[t];
    // lowp == p32 pointer means "no low-precision copy"
    void* lpv = (lowp[t] == (void*)p32[t]) ? nullptr : lowp[t];
#define RUN(G, L) adam_chunk<G, L>(pp, (const G*)grads[t] + off, mm, vv, lpv ? (L*)lpv + off : (L*)nullptr, len, lr, b1, b2, eps, wd, rbc1, rbc2, adamw, gs)
    if (gd == kF32) {
      if (ld == kBF16) RUN(float, __nv_bfloat16); else if (ld == kF16) RUN(float, __half); else RUN(float, float);
    } else if (gd == kBF16) {
      if (ld == kBF16) RUN(__nv_bfloat16, __nv_bfloat16); else if (ld == kF16) RUN(__nv_bfloat16, __half); else RUN(__nv_bfloat16, float);
    } else {
      if (ld == kBF16) RUN(__half, __nv_bfloat16); else if (ld == kF16) RUN(__half, __half); else RUN(__half, float);
    }
#undef RUN
  }
}

}  // namespace mb200

using namespace mb200;

// For the three entry points `sizes` points at [sizes(n) | prefix(n+1)] in device memory.
extern "C" void mb200_multi_l2norm(const void* const* ptrs, const long* sizes, const int* dtypes, int n, float* partial, float* out, int nblocks,
                                   cudaStream_t s) {
  l2norm_kernel<<<nblocks, 512, 0, s>>>(ptrs, sizes, dtypes, sizes + n, n, partial);
  l2norm_final_kernel<<<1, 256, 0, s>>>(partial, nblocks, out);
}
extern "C" void mb200_multi_scale(void* const* ptrs, const long* sizes, const int* dtypes, int n, const float* scale, int nblocks, cudaStream_t s) {
  scale_kernel<<<nblocks, 512, 0, s>>>(ptrs, sizes, dtypes, sizes + n, n, scale);
}
extern "C" void mb200_multi_adam(float* const* p32, const void* const* grads, float* const* m, float* const* v, void* const* lowp, const long* sizes,
                                 const int* gdtypes, const int* ldtypes, int n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                                 int adamw, const float* grad_scale, int nblocks, cudaStream_t s) {
  adam_kernel<<<nblocks, 512, 0, s>>>(p32, grads, m, v, lowp, sizes, gdtypes, ldtypes, sizes + n, n, lr, b1, b2, eps, wd, 1.f / bc1, 1.f / bc2, adamw,
                                      grad_scale);
}
