// Fused vocab-parallel cross entropy for sm_100a.
//
// ce_stats:  ONE pass over a logits row produces (row max, sum exp(x - max), target logit
//            if the target falls in this rank's vocab range) — online softmax, so the
//            [rows, V] tensor is read exactly once (reference: max pass + exp/sum pass +
//            gather, `tensor_parallel/cross_entropy.py:13-235`).
// ce_bwd:    overwrites the logits buffer IN PLACE with (softmax - onehot) * dloss, so the
//            backward allocates nothing (logits for Llama-3 vocab are 2 GiB / 8192 tokens).
#include "common.cuh"

namespace mb200 {

// stats layout: [3][rows] fp32
template <typename T>
__global__ void __launch_bounds__(256) ce_stats_kernel(const T* __restrict__ logits, const long* __restrict__ target, float* __restrict__ stats, int rows,
                                                         int V, long vocab_start) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[32];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* lr = logits + (size_t)row * V;
    float m = -INFINITY, s = 0.f;
    const int nvec = V / VN;
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
      Vec<T> t = ld16_stream(lr + vi * VN);
      float lm = to_f(t.v[0]);
#pragma unroll
      for (int k = 1; k < VN; ++k) lm = fmaxf(lm, to_f(t.v[k]));
      if (lm > m) {
        s *= __expf(m - lm);
        m = lm;
      }
#pragma unroll
      for (int k = 0; k < VN; ++k) s += __expf(to_f(t.v[k]) - m);
    }
    for (int c = nvec * VN + threadIdx.x; c < V; c += blockDim.x) {  // ragged tail
      const float xv = to_f(lr[c]);
      if (xv > m) {
        s *= __expf(m - xv);
        m = xv;
      }
      s += __expf(xv - m);
    }
    const float gm = block_max(m, red);
    s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
      const long tl = target[row] - vocab_start;
      stats[row] = gm;
      stats[rows + row] = s;
      stats[2 * (size_t)rows + row] = (tl >= 0 && tl < V) ? to_f(lr[tl]) : 0.f;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(T* __restrict__ logits, const long* __restrict__ target, const float* __restrict__ lse,
                                                       const float* __restrict__ gloss, int rows, int V, long vocab_start) {
  constexpr int VN = Vec<T>::N;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    T* lr = logits + (size_t)row * V;
    const float l = lse[row], g = gloss[row];
    const long tl = target[row] - vocab_start;
    const int nvec = V / VN;
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
      Vec<T> t = ld16(lr + vi * VN);
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float p = __expf(to_f(t.v[k]) - l);
        if ((long)(vi * VN + k) == tl) p -= 1.f;
        o.v[k] = from_f<T>(p * g);
      }
      st16(lr + vi * VN, o);
    }
    for (int c = nvec * VN + threadIdx.x; c < V; c += blockDim.x) {
      float p = __expf(to_f(lr[c]) - l);
      if ((long)c == tl) p -= 1.f;
      lr[c] = from_f<T>(p * g);
    }
  }
}

}  // namespace mb200

using namespace mb200;

#define DISPATCH(dtype, ...)                                   \
  switch (dtype) {                                             \
    case kF32: { using T = float; __VA_ARGS__; break; }        \
    case kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    default: { using T = __half; __VA_ARGS__; break; }         \
  }

extern "C" void mb200_ce_stats(const void* logits, const long* target, float* stats, int rows, int V, long vocab_start, int dtype, cudaStream_t s) {
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
  DISPATCH(dtype, (ce_stats_kernel<T><<<grid, 256, 0, s>>>((const T*)logits, target, stats, rows, V, vocab_start)));
}
extern "C" void mb200_ce_bwd(void* logits, const long* target, const float* lse, const float* gloss, int rows, int V, long vocab_start, int dtype,
                             cudaStream_t s) {
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
  DISPATCH(dtype, (ce_bwd_kernel<T><<<grid, 256, 0, s>>>((T*)logits, target, lse, gloss, rows, V, vocab_start)));
}
