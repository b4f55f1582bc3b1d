// Decode-time attention over a PAGED KV cache, and the fused KV append that feeds it (sm_100a; CUDA cores — one query token per request makes
// this a pure bandwidth problem: every K/V byte of every running request is read once per layer and step).
//
//   paged_kv_append_kernel     k_new/v_new [B, hk, d] -> the request's page at its current position (block table lookup on the device, one launch for K and V;
//                              reference: inference/contexts/fused_kv_append_kernel.py)
//   paged_decode_kernel        flash-decoding: grid = (kv splits, kv heads, requests).  A CTA walks its slice of one request's pages through the block
//                              table (no gather into a contiguous buffer), 32 tokens per warp step:
//                                QK:  lane = token.  Each lane reads its token's whole K row (256 B, sector-exact) and dots it with the group's query heads held
//                                     in shared memory (broadcast reads) — no cross-lane reduction per score.
//                                softmax: online, per query head, in log2 domain; one warp max / sum per 32 tokens.
//                                PV:  lane = 4 head-dim elements.  V rows are read coalesced; p_j is broadcast by shuffle.
//                              GQA: the `REP` query heads of a KV head share every K/V load.  Partial (o, m, l) per split, then
//   paged_decode_combine_kernel   merges the splits.
// Replaces: block-table gather + SDPA in the eager engine (round 1), reference flash-decode / Triton kernels.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb200 {

__global__ void paged_kv_append_kernel(const __nv_bfloat16* __restrict__ k_new, const __nv_bfloat16* __restrict__ v_new, __nv_bfloat16* __restrict__ k_pool,
                                       __nv_bfloat16* __restrict__ v_pool, const int32_t* __restrict__ block_table, const int32_t* __restrict__ positions, int table_width,
                                       int block_size, int hk, int d) {
  const int b = blockIdx.x;
  const int pos = positions[b];
  const int blk = block_table[(size_t)b * table_width + pos / block_size];
  const int off = pos % block_size;
  const size_t row = ((size_t)blk * block_size + off) * hk * d;        // [num_blocks, block_size, hk, d]
  const int n16 = hk * d / 8;                                          // 16-byte vectors per token
  const uint4* ks = reinterpret_cast<const uint4*>(k_new + (size_t)b * hk * d);
  const uint4* vs = reinterpret_cast<const uint4*>(v_new + (size_t)b * hk * d);
  uint4* kd = reinterpret_cast<uint4*>(k_pool + row);
  uint4* vd = reinterpret_cast<uint4*>(v_pool + row);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {
    kd[i] = ks[i];
    vd[i] = vs[i];
  }
}

constexpr int PD_WARPS = 4;

template <int D, int REP>
__global__ void __launch_bounds__(PD_WARPS * 32)
paged_decode_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k_pool, const __nv_bfloat16* __restrict__ v_pool,
                    const int32_t* __restrict__ block_table, const int32_t* __restrict__ lengths, float* __restrict__ o_part, float* __restrict__ ml_part, int table_width,
                    int block_size, int hq, int hk, float scale_log2, int tokens_per_split) {
  constexpr int EPL = D / 32;                       // head-dim elements per lane in the PV phase
  const int split = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int nsplit = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lengths[b];
  const int t_begin = split * tokens_per_split;
  const int t_end = min(len, t_begin + tokens_per_split);
  __shared__ float q_s[REP][D];
  __shared__ float red_o[PD_WARPS][REP][D];
  __shared__ float red_m[PD_WARPS][REP], red_l[PD_WARPS][REP];
  for (int i = threadIdx.x; i < REP * D; i += blockDim.x) {
    const int r = i / D, e = i % D;
    q_s[r][e] = __bfloat162float(q[((size_t)b * hq + hkv * REP + r) * D + e]) * scale_log2;     // scores come out in log2 units
  }
  __syncthreads();
  float m[REP], l[REP], o[REP][EPL];
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) o[r][e] = 0.f;
  }
  const size_t tok_stride = (size_t)hk * D;         // elements between consecutive tokens of a page
  const int32_t* table = block_table + (size_t)b * table_width;
  for (int t0 = t_begin + warp * 32; t0 < t_end; t0 += PD_WARPS * 32) {
    const int t = t0 + lane;
    const bool valid = t < t_end;
    // ---- QK: this lane's token --------------------------------------------------------------------------------------
    float s[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) s[r] = 0.f;
    size_t row = 0;
    if (valid) {
      row = ((size_t)table[t / block_size] * block_size + t % block_size) * tok_stride + (size_t)hkv * D;
      const uint4* kr = reinterpret_cast<const uint4*>(k_pool + row);
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        const uint4 kv = kr[c];
        const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&kv);
        float kf[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(kh[e]);
          kf[2 * e] = f.x;
          kf[2 * e + 1] = f.y;
        }
#pragma unroll
        for (int r = 0; r < REP; ++r) {
          const float4 qa = *reinterpret_cast<const float4*>(&q_s[r][c * 8]), qb = *reinterpret_cast<const float4*>(&q_s[r][c * 8 + 4]);
          s[r] = fmaf(kf[0], qa.x, fmaf(kf[1], qa.y, fmaf(kf[2], qa.z, fmaf(kf[3], qa.w, s[r]))));
          s[r] = fmaf(kf[4], qb.x, fmaf(kf[5], qb.y, fmaf(kf[6], qb.z, fmaf(kf[7], qb.w, s[r]))));
        }
      }
    }
    // ---- online softmax over the 32 tokens of this step -------------------------------------------------------------------
    float p[REP];
#pragma unroll
    for (int r = 0; r < REP; ++r) {
      float sm = valid ? s[r] : -INFINITY;
#pragma unroll
      for (int w = 16; w > 0; w >>= 1) sm = fmaxf(sm, __shfl_xor_sync(0xffffffffu, sm, w));
      const float m_new = fmaxf(m[r], sm);           // finite: the step has at least one valid token
      const float alpha = exp2f(m[r] - m_new);
      p[r] = valid ? exp2f(s[r] - m_new) : 0.f;
      float ps = p[r];
#pragma unroll
      for (int w = 16; w > 0; w >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, w);
      l[r] = l[r] * alpha + ps;
      m[r] = m_new;
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[r][e] *= alpha;
    }
    // ---- PV: lanes over the head dim; V rows coalesced ------------------------------------------------------------------------
    const int n_here = min(32, t_end - t0);
    for (int j = 0; j < n_here; ++j) {
      const size_t rj = __shfl_sync(0xffffffffu, row, j);
      const __nv_bfloat16* vr = v_pool + rj + lane * EPL;
      float vf[EPL];
      if (EPL == 4) {
        const uint2 vv = *reinterpret_cast<const uint2*>(vr);
        const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&vv);
        const float2 a = __bfloat1622float2(vh[0]), c = __bfloat1622float2(vh[1]);
        vf[0] = a.x; vf[1] = a.y; vf[2] = c.x; vf[3] = c.y;
      } else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) vf[e] = __bfloat162float(vr[e]);
      }
#pragma unroll
      for (int r = 0; r < REP; ++r) {
        const float pj = __shfl_sync(0xffffffffu, p[r], j);
#pragma unroll
        for (int e = 0; e < EPL; ++e) o[r][e] = fmaf(pj, vf[e], o[r][e]);
      }
    }
  }
  // ---- merge the warps of this CTA, write the split's partial ---------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < REP; ++r) {
    if (lane == 0) {
      red_m[warp][r] = m[r];
      red_l[warp][r] = l[r];
    }
#pragma unroll
    for (int e = 0; e < EPL; ++e) red_o[warp][r][lane * EPL + e] = o[r][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < REP * D; i += blockDim.x) {
    const int r = i / D, e = i % D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < PD_WARPS; ++w) M = fmaxf(M, red_m[w][r]);
    float acc = 0.f, L = 0.f;
#pragma unroll
    for (int w = 0; w < PD_WARPS; ++w) {
      const float a = red_m[w][r] == -INFINITY ? 0.f : exp2f(red_m[w][r] - M);
      acc = fmaf(a, red_o[w][r][e], acc);
      L = fmaf(a, red_l[w][r], L);
    }
    const size_t h = (size_t)b * hq + hkv * REP + r;
    o_part[(h * nsplit + split) * D + e] = acc;
    if (e == 0) {
      ml_part[(h * nsplit + split) * 2] = M;
      ml_part[(h * nsplit + split) * 2 + 1] = L;
    }
  }
}

template <int D>
__global__ void paged_decode_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part, __nv_bfloat16* __restrict__ out, int nsplit) {
  const size_t h = blockIdx.x;                     // (request, query head)
  const int e = threadIdx.x;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, ml_part[(h * nsplit + s) * 2]);
  float acc = 0.f, L = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = ml_part[(h * nsplit + s) * 2];
    const float a = ms == -INFINITY ? 0.f : exp2f(ms - M);
    acc = fmaf(a, o_part[(h * nsplit + s) * D + e], acc);
    L = fmaf(a, ml_part[(h * nsplit + s) * 2 + 1], L);
  }
  out[h * D + e] = __float2bfloat16_rn(L > 0.f ? acc / L : 0.f);
}

template <int D, int REP>
static void launch_decode(const void* q, const void* kp, const void* vp, const int32_t* table, const int32_t* lengths, float* o_part, float* ml_part, void* out, int B, int hq,
                          int hk, int table_width, int block_size, float scale, int nsplit, int tokens_per_split, cudaStream_t s) {
  dim3 grid(nsplit, hk, B);
  paged_decode_kernel<D, REP><<<grid, PD_WARPS * 32, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(kp),
                                                              reinterpret_cast<const __nv_bfloat16*>(vp), table, lengths, o_part, ml_part, table_width, block_size, hq, hk,
                                                              scale * 1.4426950408889634f, tokens_per_split);
  paged_decode_combine_kernel<D><<<B * hq, D, 0, s>>>(o_part, ml_part, reinterpret_cast<__nv_bfloat16*>(out), nsplit);
}

}  // namespace mb200

using namespace mb200;

extern "C" void mb200_paged_kv_append(const void* k_new, const void* v_new, void* k_pool, void* v_pool, const int32_t* block_table, const int32_t* positions, int B,
                                      int table_width, int block_size, int hk, int d, cudaStream_t s) {
  paged_kv_append_kernel<<<B, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(k_new), reinterpret_cast<const __nv_bfloat16*>(v_new),
                                           reinterpret_cast<__nv_bfloat16*>(k_pool), reinterpret_cast<__nv_bfloat16*>(v_pool), block_table, positions, table_width,
                                           block_size, hk, d);
}

// q, out: [B, hq, d] bf16; pools: [num_blocks, block_size, hk, d] bf16; block_table [B, table_width] int32; lengths [B] int32 (valid tokens incl. the new one);
// o_part: fp32 [B*hq*nsplit*d], ml_part: fp32 [B*hq*nsplit*2].  Returns 0, or -1 for an unsupported (d, hq/hk).
extern "C" int mb200_paged_decode(const void* q, const void* k_pool, const void* v_pool, const int32_t* block_table, const int32_t* lengths, float* o_part, float* ml_part,
                                  void* out, int B, int hq, int hk, int d, int table_width, int block_size, float scale, int nsplit, int tokens_per_split, cudaStream_t s) {
  const int rep = hq / hk;
  if (hq % hk != 0) return -1;
#define MB200_PD(DD, RR)                                                                                                                                       \
  if (d == DD && rep == RR) {                                                                                                                                 \
    launch_decode<DD, RR>(q, k_pool, v_pool, block_table, lengths, o_part, ml_part, out, B, hq, hk, table_width, block_size, scale, nsplit, tokens_per_split, s); \
    return cudaGetLastError() == cudaSuccess ? 0 : -4;                                                                                                        \
  }
  MB200_PD(128, 1) MB200_PD(128, 2) MB200_PD(128, 4) MB200_PD(128, 8) MB200_PD(64, 1) MB200_PD(64, 2) MB200_PD(64, 4) MB200_PD(64, 8)
#undef MB200_PD
  return -1;
}
