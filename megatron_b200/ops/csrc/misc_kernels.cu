// Memory-saver and serving helpers that the reference writes as Triton kernels, plus the bias-dropout-add fusion:
//   * paged stash copy / pop           (reference transformer/moe/ops/paged_stash.py: paged_stash_copy_kernel / paged_stash_pop_kernel)
//   * batched speculative-decode verify (reference inference/text_generation_controllers/mtp_utils_triton.py: verify / rewind)
//   * residual + dropout(x + bias)      (reference fusions/fused_bias_dropout.py, a jit-scripted elementwise chain)
// All are bandwidth / latency bound: one launch each, 16-byte vector accesses, device-resident control values (no host sync).
#include <curand_kernel.h>

#include "common.cuh"

namespace mb200 {

// ---- paged stash ---------------------------------------------------------------------------------------------------------------------
// src [T_max, H] -> pages[(page_ids[r / ps] * ps + r % ps), :] for r < *num_tokens (rows beyond the valid count are not written);
// pop: the inverse gather, rows beyond the valid count are zero-filled.  One warp per row, 16-byte vectors (H * sizeof(T) % 16 == 0).
template <bool POP>
__global__ void __launch_bounds__(256) paged_stash_kernel(uint4* __restrict__ flat, uint4* __restrict__ pages, const int64_t* __restrict__ page_ids,
                                                          const int64_t* __restrict__ num_tokens, long t_max, int vec_per_row, int page_size) {
  const long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= t_max) return;
  const bool valid = r < *num_tokens;
  uint4* frow = flat + r * vec_per_row;
  if (!valid) {
    if (POP)
      for (int c = lane; c < vec_per_row; c += 32) frow[c] = make_uint4(0, 0, 0, 0);
    return;
  }
  uint4* prow = pages + (page_ids[r / page_size] * page_size + r % page_size) * vec_per_row;
  for (int c = lane; c < vec_per_row; c += 32) {
    if (POP) frow[c] = prow[c];
    else prow[c] = frow[c];
  }
}

// ---- speculative decoding: accept / reject k draft tokens per sequence, pick the token that follows the accepted prefix ------------------
// draft_tokens [B, k]; draft_probs [B, k, V] (nullptr = greedy draft: one-hot on the draft token); target_probs [B, k + 1, V];
// u_accept [B, k], u_sample [B] uniforms in [0, 1).  Outputs n_accepted [B], next_token [B].  One block per sequence.
// Rule (Leviathan et al.): accept token i while u_i < min(1, p_t / p_d); at the first rejection sample from normalise(max(p_t - p_d, 0));
// after k acceptances sample the bonus token from the last target row.  Sampling is an inverse-CDF walk over the vocabulary in index order.
__global__ void __launch_bounds__(256) spec_verify_kernel(const int64_t* __restrict__ draft_tokens, const float* __restrict__ draft_probs, const float* __restrict__ target_probs,
                                                          const float* __restrict__ u_accept, const float* __restrict__ u_sample, int64_t* __restrict__ n_accepted,
                                                          int64_t* __restrict__ next_token, int k, int V) {
  __shared__ float red[32];
  __shared__ float warp_part[8];
  __shared__ int s_n, s_choice;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* tp = target_probs + (size_t)b * (k + 1) * V;
  const float* dp = draft_probs != nullptr ? draft_probs + (size_t)b * k * V : nullptr;
  if (tid == 0) {
    int n = 0;
    for (; n < k; ++n) {
      const int64_t tok = draft_tokens[(size_t)b * k + n];
      const float pt = tp[(size_t)n * V + tok];
      const float pd = dp != nullptr ? fmaxf(dp[(size_t)n * V + tok], 1e-20f) : 1.f;
      if (!(u_accept[(size_t)b * k + n] < fminf(pt / pd, 1.f))) break;
    }
    s_n = n;
    s_choice = -1;
  }
  __syncthreads();
  const int n = s_n;
  const float* trow = tp + (size_t)n * V;
  const float* drow = (n < k && dp != nullptr) ? dp + (size_t)n * V : nullptr;
  const int64_t dtok = n < k ? draft_tokens[(size_t)b * k + n] : -1;
  auto residual = [&](int v) -> float {
    float t = trow[v];
    if (n < k) t = fmaxf(t - (drow != nullptr ? drow[v] : (v == dtok ? 1.f : 0.f)), 0.f);
    return t;
  };
  float acc = 0.f;
  for (int v = tid; v < V; v += 256) acc += residual(v);
  float total = block_sum(acc, red);
  const bool degenerate = !(total > 0.f);                 // residual vanished (p_t <= p_d everywhere): fall back to the target row itself
  if (degenerate) {
    acc = 0.f;
    for (int v = tid; v < V; v += 256) acc += trow[v];
    total = block_sum(acc, red);
  }
  const float threshold = u_sample[b] * total;
  // walk the vocabulary in chunks of 256 in index order; running = mass before the chunk
  float running = 0.f;
  for (int v0 = 0; v0 < V && s_choice < 0; v0 += 256) {
    const int v = v0 + tid;
    const float x = v < V ? (degenerate ? trow[v] : residual(v)) : 0.f;
    float incl = x;                                        // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_part[wid] = incl;
    __syncthreads();
    float before = running, chunk = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      before += w < wid ? warp_part[w] : 0.f;
      chunk += warp_part[w];
    }
    const float hi = before + incl, lo = hi - x;
    if (v < V && x > 0.f && lo <= threshold && threshold < hi) atomicMax(&s_choice, v);
    running += chunk;
    __syncthreads();
  }
  if (tid == 0) {
    int choice = s_choice;
    if (choice < 0) {                                      // rounding left the threshold at/after the total mass: last token with non-zero mass
      for (int v = V - 1; v >= 0; --v)
        if ((degenerate ? trow[v] : residual(v)) > 0.f) { choice = v; break; }
      if (choice < 0) choice = 0;
    }
    n_accepted[b] = n;
    next_token[b] = choice;
  }
}

// ---- y = residual + dropout(x + bias, p) ------------------------------------------------------------------------------------------------
// The keep mask is a pure function of (seed, offset, element index): the backward regenerates it instead of storing it.
// Thread t owns vectors t, t + total, ...; vector i of 8 elements takes Philox subsequence t, draws 2 * iteration and 2 * iteration + 1.
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) bias_dropout_add_kernel(const T* __restrict__ x, const T* __restrict__ bias, const T* __restrict__ residual, T* __restrict__ y, long nvec,
                                                               int H, float p, float scale, unsigned long long seed, unsigned long long offset) {
  constexpr int N = VecTraits<T>::N;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, total = (long)gridDim.x * blockDim.x;
  curandStatePhilox4_32_10_t st;
  if (p > 0.f) curand_init(seed, (unsigned long long)tid, offset, &st);
  for (long i = tid; i < nvec; i += total) {
    float keep[8];
    if (p > 0.f) {
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        const float4 r = curand_uniform4(&st);
        keep[j] = r.x >= p ? scale : 0.f;
        keep[j + 1] = r.y >= p ? scale : 0.f;
        keep[j + 2] = r.z >= p ? scale : 0.f;
        keep[j + 3] = r.w >= p ? scale : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) keep[j] = 1.f;
    }
    const Vec<T> xv = ld16(x + i * N);
    Vec<T> out;
    if (BWD) {                                             // x = grad of y -> grad of x (= grad of bias rows)
#pragma unroll
      for (int j = 0; j < N; ++j) out.v[j] = from_f<T>(to_f(xv.v[j]) * keep[j]);
    } else {
      const Vec<T> rv = ld16(residual + i * N);
      const int col = (int)((i * N) % H);
      Vec<T> bv;
      if (bias != nullptr) bv = ld16(bias + col);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float v = to_f(xv.v[j]) + (bias != nullptr ? to_f(bv.v[j]) : 0.f);
        out.v[j] = from_f<T>(to_f(rv.v[j]) + v * keep[j]);
      }
    }
    st16(y + i * N, out);
  }
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_paged_stash(void* flat, void* pages, const int64_t* page_ids, const int64_t* num_tokens, long t_max, long row_bytes, int page_size, int pop, cudaStream_t s) {
  if (row_bytes % 16 != 0) return -1;
  if (t_max == 0) return 0;
  const unsigned grid = (unsigned)((t_max + 7) / 8);
  if (pop) paged_stash_kernel<true><<<grid, 256, 0, s>>>((uint4*)flat, (uint4*)pages, page_ids, num_tokens, t_max, (int)(row_bytes / 16), page_size);
  else paged_stash_kernel<false><<<grid, 256, 0, s>>>((uint4*)flat, (uint4*)pages, page_ids, num_tokens, t_max, (int)(row_bytes / 16), page_size);
  return 0;
}
extern "C" void mb200_spec_verify(const int64_t* draft_tokens, const float* draft_probs, const float* target_probs, const float* u_accept, const float* u_sample,
                                  int64_t* n_accepted, int64_t* next_token, int B, int k, int V, cudaStream_t s) {
  if (B > 0) spec_verify_kernel<<<B, 256, 0, s>>>(draft_tokens, draft_probs, target_probs, u_accept, u_sample, n_accepted, next_token, k, V);
}
// returns the number of Philox draws per thread (the generator offset must advance by it), or -1
extern "C" long mb200_bias_dropout_add_draws(long numel, int dtype, int* grid_out) {
  const int N = dtype == 0 ? 4 : 8;
  const long nvec = numel / N;
  const long grid = nvec == 0 ? 1 : (nvec + 255) / 256 < 1184 ? (nvec + 255) / 256 : 1184;   // 8 CTAs of 256 threads per SM
  *grid_out = (int)grid;
  const long total = grid * 256;
  return ((nvec + total - 1) / total) * (N / 4) * 4;
}
extern "C" int mb200_bias_dropout_add(const void* x, const void* bias, const void* residual, void* y, long numel, int H, float p, unsigned long long seed,
                                      unsigned long long offset, int backward, int dtype, cudaStream_t s) {
  const int N = dtype == 0 ? 4 : 8;
  if (numel % N != 0 || H % N != 0) return -1;
  int grid;
  mb200_bias_dropout_add_draws(numel, dtype, &grid);
  const long nvec = numel / N;
  const float scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
#define MB200_BDA(T)                                                                                                                                            \
  if (backward) bias_dropout_add_kernel<T, true><<<grid, 256, 0, s>>>((const T*)x, nullptr, nullptr, (T*)y, nvec, H, p, scale, seed, offset);                   \
  else bias_dropout_add_kernel<T, false><<<grid, 256, 0, s>>>((const T*)x, (const T*)bias, (const T*)residual, (T*)y, nvec, H, p, scale, seed, offset);
  if (dtype == 1) { MB200_BDA(__nv_bfloat16) } else if (dtype == 2) { MB200_BDA(__half) } else { MB200_BDA(float) }
#undef MB200_BDA
  return 0;
}
