// MoE data-movement and routing kernels for sm_100a (replace TE ``moe_permute`` / ``moe_unpermute`` / ``fused_topk_with_score_function``,
// SURVEY X14-X15).  All are HBM-bound: 16-byte vector accesses, one CTA per output row, fp32 accumulation where rows are summed.
//
//   moe_gather_rows    out[i, :]  = scale[i] * in[src[i], :]                       permute fwd, unpermute bwd
//   moe_combine_rows   out[t, :]  = Σ_k w[t,k] * in[pos[t,k], :]   (pos < 0 skipped) unpermute fwd (deterministic, no atomics), permute bwd
//   moe_topk_router    logits [T, E] → softmax|sigmoid scores → top-k (k ≤ 8) ids + (optionally renormalised) probs, routing map [T, E],
//                      tokens-per-expert histogram (one atomic per selection); one warp per token, E ≤ 256.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mb200 {

__global__ void __launch_bounds__(256) moe_gather_rows_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, const int64_t* __restrict__ src,
                                                               const float* __restrict__ scale, int row_vecs) {
  const int64_t i = blockIdx.x;
  const int64_t s = src[i];
  const uint4* ip = in + s * row_vecs;
  uint4* op = out + i * row_vecs;
  if (scale == nullptr) {
    for (int v = threadIdx.x; v < row_vecs; v += blockDim.x) op[v] = __ldg(ip + v);
  } else {
    const float w = scale[i];
    for (int v = threadIdx.x; v < row_vecs; v += blockDim.x) {
      uint4 x = __ldg(ip + v);
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = __bfloat1622float2(h[k]);
        h[k] = __floats2bfloat162_rn(f.x * w, f.y * w);
      }
      op[v] = x;
    }
  }
}

__global__ void __launch_bounds__(256) moe_combine_rows_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, const int64_t* __restrict__ pos,
                                                                const float* __restrict__ w, int topk, int row_vecs) {
  const int64_t t = blockIdx.x;
  for (int v = threadIdx.x; v < row_vecs; v += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int j = 0; j < topk; ++j) {
      const int64_t p = pos[t * topk + j];
      if (p < 0) continue;
      const float wj = w != nullptr ? w[t * topk + j] : 1.f;
      const uint4 x = __ldg(in + p * row_vecs + v);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        acc[2 * k] += wj * f.x;
        acc[2 * k + 1] += wj * f.y;
      }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
    out[t * row_vecs + v] = o;
  }
}

constexpr int ROUTER_MAX_E_PER_LANE = 8;   // E <= 256
constexpr int ROUTER_MAX_K = 8;

// score_fn: 0 softmax over all experts (pre-top-k softmax), 1 sigmoid, 2 softmax over the selected k (post-top-k softmax)
__global__ void __launch_bounds__(128) moe_topk_router_kernel(const float* __restrict__ logits, const float* __restrict__ expert_bias, int T, int E, int topk, int score_fn,
                                                               int renormalize, float scaling, float* __restrict__ probs, int64_t* __restrict__ ids,
                                                               uint8_t* __restrict__ routing_map, int* __restrict__ tokens_per_expert) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T) return;
  const float* row = logits + (size_t)warp * E;
  float s[ROUTER_MAX_E_PER_LANE], sel[ROUTER_MAX_E_PER_LANE];   // score, selection key (score + bias)
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j) {
    const int e = j * 32 + lane;
    s[j] = e < E ? row[e] : -CUDART_INF_F;
    mx = fmaxf(mx, s[j]);
  }
  if (score_fn == 0) {
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j) {
      s[j] = j * 32 + lane < E ? __expf(s[j] - mx) : 0.f;
      sum += s[j];
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j) s[j] *= inv;
  } else if (score_fn == 1) {
#pragma unroll
    for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j) s[j] = j * 32 + lane < E ? 1.f / (1.f + __expf(-s[j])) : -CUDART_INF_F;
  }
#pragma unroll
  for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j) {
    const int e = j * 32 + lane;
    sel[j] = e < E ? s[j] + (expert_bias != nullptr ? expert_bias[e] : 0.f) : -CUDART_INF_F;
  }
  float top_v[ROUTER_MAX_K];
  int top_e[ROUTER_MAX_K];
  for (int k = 0; k < topk; ++k) {
    // warp arg-max (ties → lowest expert id, matching torch.topk's stable behaviour on sorted input)
    float bv = -CUDART_INF_F;
    int be = 0x7fffffff, bj = 0;
#pragma unroll
    for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j) {
      const int e = j * 32 + lane;
      if (sel[j] > bv || (sel[j] == bv && e < be)) { bv = sel[j]; be = e; bj = j; }
    }
    float wv = bv;
    int we = be;
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oe = __shfl_xor_sync(0xffffffffu, we, o);
      if (ov > wv || (ov == wv && oe < we)) { wv = ov; we = oe; }
    }
    // the owning lane publishes the un-biased score and retires the expert
    float val = 0.f;
    if (we == be && be < E) {
#pragma unroll
      for (int j = 0; j < ROUTER_MAX_E_PER_LANE; ++j)
        if (j == bj) { val = s[j]; sel[j] = -CUDART_INF_F; }
    }
    val = __shfl_sync(0xffffffffu, val, we & 31);
    top_v[k] = val;
    top_e[k] = we;
  }
  if (score_fn == 2) {   // softmax over the selected logits
    float m2 = -CUDART_INF_F, sum = 0.f;
    for (int k = 0; k < topk; ++k) m2 = fmaxf(m2, top_v[k]);
    for (int k = 0; k < topk; ++k) { top_v[k] = __expf(top_v[k] - m2); sum += top_v[k]; }
    for (int k = 0; k < topk; ++k) top_v[k] /= sum;
  } else if (renormalize && topk > 1) {
    float sum = 1e-20f;
    for (int k = 0; k < topk; ++k) sum += top_v[k];
    for (int k = 0; k < topk; ++k) top_v[k] /= sum;
  }
  if (lane == 0) {
    for (int k = 0; k < topk; ++k) {
      probs[(size_t)warp * topk + k] = top_v[k] * scaling;
      ids[(size_t)warp * topk + k] = top_e[k];
      if (routing_map != nullptr) routing_map[(size_t)warp * E + top_e[k]] = 1;
      if (tokens_per_expert != nullptr) atomicAdd(&tokens_per_expert[top_e[k]], 1);
    }
  }
}

}  // namespace mb200

using namespace mb200;

extern "C" void mb200_moe_gather_rows(const void* in, void* out, const int64_t* src, const float* scale, int64_t n_out, int hidden, cudaStream_t s) {
  if (n_out == 0) return;
  moe_gather_rows_kernel<<<(unsigned)n_out, 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), src, scale, hidden / 8);
}
extern "C" void mb200_moe_combine_rows(const void* in, void* out, const int64_t* pos, const float* w, int64_t n_tokens, int topk, int hidden, cudaStream_t s) {
  if (n_tokens == 0) return;
  moe_combine_rows_kernel<<<(unsigned)n_tokens, 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), pos, w, topk, hidden / 8);
}
extern "C" int mb200_moe_topk_router(const float* logits, const float* expert_bias, int T, int E, int topk, int score_fn, int renormalize, float scaling, float* probs,
                                     int64_t* ids, uint8_t* routing_map, int* tokens_per_expert, cudaStream_t s) {
  if (E > 32 * ROUTER_MAX_E_PER_LANE || topk > ROUTER_MAX_K || topk < 1) return -10;
  if (T == 0) return 0;
  const int warps_per_block = 4;
  moe_topk_router_kernel<<<(T + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, s>>>(logits, expert_bias, T, E, topk, score_fn, renormalize, scaling, probs,
                                                                                                    ids, routing_map, tokens_per_expert);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// ---- expert-parallel dispatch / combine over NVLink peer memory -------------------------------------------------------------------
// (replaces DeepEP / HybridEP / NCCL-EP + the permute kernels around them, SURVEY X16-X18: ONE kernel moves every token row
//  straight into the slot of the destination rank's expert-grouped buffer, so the all-to-all and the permutation are the same pass)
namespace mb200 {

struct PeerPtrs {
  void* p[8];
};

// push: row src_row[i] of `in` → peer dst_rank[i], row dst_slot[i] of its receive buffer (16-byte vectors, one CTA per pair)
__global__ void __launch_bounds__(256) moe_push_rows_kernel(const uint4* __restrict__ in, const int64_t* __restrict__ src_row, const int32_t* __restrict__ dst_rank,
                                                            const int64_t* __restrict__ dst_slot, PeerPtrs peers, size_t dst_off_vecs, int row_vecs) {
  const int64_t i = blockIdx.x;
  const uint4* ip = in + src_row[i] * row_vecs;
  uint4* op = reinterpret_cast<uint4*>(peers.p[dst_rank[i]]) + dst_off_vecs + dst_slot[i] * row_vecs;
  for (int v = threadIdx.x; v < row_vecs; v += blockDim.x) op[v] = __ldg(ip + v);
}

// pull: out[t] = Σ_k w[t,k] · peer[rank[t,k]].buf[slot[t,k]]   (slot < 0 skipped; fp32 accumulation; bf16 rows or raw 16-byte copy for k = 1)
template <bool RAW>
__global__ void __launch_bounds__(256) moe_pull_rows_kernel(uint4* __restrict__ out, const int32_t* __restrict__ rank, const int64_t* __restrict__ slot,
                                                            const float* __restrict__ w, PeerPtrs peers, size_t src_off_vecs, int topk, int row_vecs) {
  const int64_t t = blockIdx.x;
  for (int v = threadIdx.x; v < row_vecs; v += blockDim.x) {
    if (RAW) {
      const int64_t s = slot[t];
      out[t * row_vecs + v] = s < 0 ? make_uint4(0, 0, 0, 0) : reinterpret_cast<const uint4*>(peers.p[rank[t]])[src_off_vecs + s * row_vecs + v];
      continue;
    }
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int j = 0; j < topk; ++j) {
      const int64_t s = slot[t * topk + j];
      if (s < 0) continue;
      const float wj = w != nullptr ? w[t * topk + j] : 1.f;
      const uint4 x = reinterpret_cast<const uint4*>(peers.p[rank[t * topk + j]])[src_off_vecs + s * row_vecs + v];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        acc[2 * k] += wj * f.x;
        acc[2 * k + 1] += wj * f.y;
      }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
    out[t * row_vecs + v] = o;
  }
}

}  // namespace mb200

extern "C" void mb200_moe_push_rows(const void* in, const int64_t* src_row, const int32_t* dst_rank, const int64_t* dst_slot, const int64_t* peer_ptrs, int world,
                                    size_t dst_off_bytes, int64_t n_pairs, int row_bytes, cudaStream_t s) {
  if (n_pairs == 0) return;
  mb200::PeerPtrs pp;
  for (int i = 0; i < 8; ++i) pp.p[i] = i < world ? reinterpret_cast<void*>(peer_ptrs[i]) : nullptr;
  mb200::moe_push_rows_kernel<<<(unsigned)n_pairs, 256, 0, s>>>(reinterpret_cast<const uint4*>(in), src_row, dst_rank, dst_slot, pp, dst_off_bytes / 16, row_bytes / 16);
}
extern "C" void mb200_moe_pull_rows(void* out, const int32_t* rank, const int64_t* slot, const float* w, const int64_t* peer_ptrs, int world, size_t src_off_bytes,
                                    int64_t n_out, int topk, int row_bytes, int raw, cudaStream_t s) {
  if (n_out == 0) return;
  mb200::PeerPtrs pp;
  for (int i = 0; i < 8; ++i) pp.p[i] = i < world ? reinterpret_cast<void*>(peer_ptrs[i]) : nullptr;
  if (raw)
    mb200::moe_pull_rows_kernel<true><<<(unsigned)n_out, 256, 0, s>>>(reinterpret_cast<uint4*>(out), rank, slot, w, pp, src_off_bytes / 16, 1, row_bytes / 16);
  else
    mb200::moe_pull_rows_kernel<false><<<(unsigned)n_out, 256, 0, s>>>(reinterpret_cast<uint4*>(out), rank, slot, w, pp, src_off_bytes / 16, topk, row_bytes / 16);
}
