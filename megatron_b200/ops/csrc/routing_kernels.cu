// Router bookkeeping and MLA rotary kernels (indices <-> multi-hot converter, pad-routing-map, fused MoE aux loss, MLA YaRN RoPE apply).
// The reference implements these as Triton kernels (core/fusions/fused_indices_converter.py, fused_pad_routing_map.py, fused_mla_yarn_rope_apply.py) and
// Transformer-Engine router fusions; here they are plain CUDA: all of them are bandwidth-trivial and launch-latency bound, so each is ONE launch with
// warp-level primitives (ballot compaction, shuffle reductions) and no intermediate tensors.
#include "common.cuh"

namespace mb200 {

// ---- top-k indices [T, k] (-1 = dropped) + probs [T, k]  ->  multi-hot map [T, E] (uint8) + probs [T, E] -------------------------------------------------
// one warp per token: zero the row, then scatter the k entries (no pre-zeroed outputs, no atomics)
__global__ void __launch_bounds__(256) indices_to_multihot_kernel(const int64_t* __restrict__ idx, const float* __restrict__ probs, uint8_t* __restrict__ map,
                                                                  float* __restrict__ probs_out, long T, int k, int E) {
  const long t = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  for (int e = lane; e < E; e += 32) {
    map[t * E + e] = 0;
    probs_out[t * E + e] = 0.f;
  }
  __syncwarp();
  for (int j = lane; j < k; j += 32) {
    const int64_t e = idx[t * k + j];
    if (e >= 0 && e < E) {
      map[t * E + e] = 1;
      probs_out[t * E + e] = probs[t * k + j];
    }
  }
}

// gather: grad wrt probs [T, k] from grad wrt probs [T, E] (and the inverse scatter, selected by `scatter`)
__global__ void __launch_bounds__(256) multihot_probs_grad_kernel(const int64_t* __restrict__ idx, const float* __restrict__ g_in, float* __restrict__ g_out, long T, int k, int E,
                                                                  int scatter) {
  const long t = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  if (scatter) {                                     // g_in [T, k] -> g_out [T, E]
    for (int e = lane; e < E; e += 32) g_out[t * E + e] = 0.f;
    __syncwarp();
    for (int j = lane; j < k; j += 32) {
      const int64_t e = idx[t * k + j];
      if (e >= 0 && e < E) g_out[t * E + e] = g_in[t * k + j];
    }
  } else {                                           // g_in [T, E] -> g_out [T, k]
    for (int j = lane; j < k; j += 32) {
      const int64_t e = idx[t * k + j];
      g_out[t * k + j] = (e >= 0 && e < E) ? g_in[t * E + e] : 0.f;
    }
  }
}

// multi-hot map [T, E] + probs [T, E] -> indices [T, k] in expert order (padded with -1) + probs [T, k]: ordered stream compaction by warp ballot
__global__ void __launch_bounds__(256) multihot_to_indices_kernel(const uint8_t* __restrict__ map, const float* __restrict__ probs, int64_t* __restrict__ idx,
                                                                  float* __restrict__ probs_out, long T, int k, int E) {
  const long t = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  int n = 0;
  for (int e0 = 0; e0 < E; e0 += 32) {
    const int e = e0 + lane;
    const bool on = e < E && map[t * E + e] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (on) {
      const int slot = n + __popc(m & ((1u << lane) - 1u));
      if (slot < k) {
        idx[t * k + slot] = e;
        probs_out[t * k + slot] = probs[t * E + e];
      }
    }
    n += __popc(m);
  }
  for (int j = n + lane; j < k; j += 32) {
    idx[t * k + j] = -1;
    probs_out[t * k + j] = 0.f;
  }
}

// ---- pad every expert's token count up to a multiple: flip the first (-count mod m) zeros of its column, lowest token first --------------------------------
// one block per expert; the column is walked in token order, zeros ranked by a block-wide scan (ballot per warp + warp totals in smem)
__global__ void __launch_bounds__(256) pad_routing_map_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long T, int E, int multiple) {
  __shared__ int warp_tot[8];
  __shared__ int s_count;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_count = 0;
  __syncthreads();
  int ones = 0;
  for (long t = tid; t < T; t += 256) ones += in[t * E + e] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
  if (lane == 0) atomicAdd(&s_count, ones);
  __syncthreads();
  int need = (multiple - s_count % multiple) % multiple;
  for (long t0 = 0; t0 < T; t0 += 256) {
    const long t = t0 + tid;
    const bool valid = t < T;
    const uint8_t v = valid ? in[t * E + e] : 1;
    const bool zero = v == 0;
    const unsigned m = __ballot_sync(0xffffffffu, zero);
    if (lane == 0) warp_tot[wid] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      before += w < wid ? warp_tot[w] : 0;
      total += warp_tot[w];
    }
    const int rank = before + __popc(m & ((1u << lane) - 1u));
    if (valid) out[t * E + e] = (zero && rank < need) ? 1 : (v != 0);
    need = max(need - total, 0);
    __syncthreads();
  }
}

// ---- switch load-balancing loss: C * sum_e (sum_t probs[t, e]) * tokens_per_expert[e]; deterministic two-stage reduction ----------------------------------
__global__ void __launch_bounds__(256) aux_loss_partial_kernel(const float* __restrict__ probs, const float* __restrict__ tpe, float* __restrict__ partial, long T, int E,
                                                               long rows_per_block) {
  __shared__ float red[32];
  const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(T, r0 + rows_per_block);
  float acc = 0.f;
  // thread = (row lane, expert): consecutive threads read consecutive experts of a row
  const long n = (r1 - r0) * E;
  for (long i = threadIdx.x; i < n; i += 256) acc += probs[r0 * E + i] * tpe[i % E];
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(256) aux_loss_final_kernel(const float* __restrict__ partial, int n, float coeff, float* __restrict__ loss) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) *loss = tot * coeff;
}
__global__ void __launch_bounds__(256) aux_loss_bwd_kernel(const float* __restrict__ tpe, const float* __restrict__ gloss, float coeff, float* __restrict__ gprobs, long n, int E) {
  const float g = *gloss * coeff;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) gprobs[i] = g * tpe[i % E];
}

// ---- MLA rotary ------------------------------------------------------------------------------------------------------------------------------------------
// x [rows, H, nope + emb]: rotate the trailing emb channels of every head IN PLACE.  Input pairs are adjacent (x[2i], x[2i+1]) when `interleaved`, else
// (x[i], x[i + emb/2]); the output is always half-split ([left | right]) — what the attention kernel and the key built by the kv-split kernel use.
// `inverse` (the backward) maps a half-split gradient back to the input layout.  ang [positions, emb] holds ANGLES (fp32); position = pos[row] or row / batch.
// One warp per (row, head), lane = pair index (emb <= 64).
template <typename T>
__global__ void __launch_bounds__(256) mla_rope_inplace_kernel(const T* src, T* x, const float* __restrict__ ang, const int64_t* __restrict__ pos, long rows, int H, int nope,
                                                               int emb, int batch, float mscale, int interleaved, int inverse) {
  const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= rows * H) return;
  const long row = w / H;
  const long p = pos != nullptr ? pos[row] : row / batch;
  const int half = emb >> 1;
  T* base = x + w * (nope + emb) + nope;
  const T* sbase = src + w * (nope + emb) + nope;
  if (src != x)                                   // out of place: the untouched channels travel in the same pass
    for (int c = lane; c < nope; c += 32) x[w * (nope + emb) + c] = src[w * (nope + emb) + c];
  float a = 0.f, b = 0.f;
  const bool on = lane < half;
  const int ia = interleaved ? 2 * lane : lane, ib = interleaved ? 2 * lane + 1 : lane + half;
  float cl = 1.f, sl = 0.f, cr = 1.f, sr = 0.f;
  if (on) {
    sincosf(ang[p * emb + lane], &sl, &cl);
    sincosf(ang[p * emb + half + lane], &sr, &cr);
    cl *= mscale; sl *= mscale; cr *= mscale; sr *= mscale;
    if (!inverse) {
      a = to_f(sbase[ia]);
      b = to_f(sbase[ib]);
    } else {
      a = to_f(sbase[lane]);          // g_left, g_right
      b = to_f(sbase[lane + half]);
    }
  }
  __syncwarp();                        // every lane has read its pair before any lane overwrites another lane's input (layouts differ when interleaved)
  if (on) {
    if (!inverse) {
      base[lane] = from_f<T>(a * cl - b * sl);
      base[lane + half] = from_f<T>(b * cr + a * sr);
    } else {
      base[ia] = from_f<T>(a * cl + b * sr);
      base[ib] = from_f<T>(b * cr - a * sl);
    }
  }
}

// kv [rows, H, kd + vd], k_pe [rows, emb] (one rotary key shared by the heads; rotated here unless ang == nullptr)
//   -> key [rows, H, kd + emb] = [kv[..., :kd] | rope(k_pe)], value [rows, H, vd].  One block per row.
template <typename T>
__global__ void __launch_bounds__(256) mla_kv_split_fwd_kernel(const T* __restrict__ kv, const T* __restrict__ kpe, const float* __restrict__ ang, const int64_t* __restrict__ pos,
                                                               T* __restrict__ key, T* __restrict__ val, int H, int kd, int vd, int emb, int batch, float mscale, int interleaved) {
  __shared__ float pe[256];
  const long row = blockIdx.x;
  const int half = emb >> 1;
  if (ang == nullptr) interleaved = 0;            // k_pe is already rotated (and already half-split): plain copy
  if ((int)threadIdx.x < half) {
    const int i = threadIdx.x;
    const float a = to_f(kpe[row * emb + (interleaved ? 2 * i : i)]), b = to_f(kpe[row * emb + (interleaved ? 2 * i + 1 : i + half)]);
    if (ang != nullptr) {
      const long p = pos != nullptr ? pos[row] : row / batch;
      float cl, sl, cr, sr;
      sincosf(ang[p * emb + i], &sl, &cl);
      sincosf(ang[p * emb + half + i], &sr, &cr);
      pe[i] = (a * cl - b * sl) * mscale;
      pe[i + half] = (b * cr + a * sr) * mscale;
    } else {
      pe[i] = a;
      pe[i + half] = b;
    }
  }
  __syncthreads();
  const int kw = kd + emb, iw = kd + vd;
  for (int i = threadIdx.x; i < H * kw; i += 256) {
    const int h = i / kw, c = i % kw;
    key[(row * H + h) * kw + c] = c < kd ? kv[(row * H + h) * iw + c] : from_f<T>(pe[c - kd]);
  }
  for (int i = threadIdx.x; i < H * vd; i += 256) {
    const int h = i / vd, c = i % vd;
    val[(row * H + h) * vd + c] = kv[(row * H + h) * iw + kd + c];
  }
}

// backward: dkey [rows, H, kd + emb], dval [rows, H, vd] -> dkv [rows, H, kd + vd], dk_pe [rows, emb] (sum over heads, inverse rotation, input layout)
template <typename T>
__global__ void __launch_bounds__(256) mla_kv_split_bwd_kernel(const T* __restrict__ dkey, const T* __restrict__ dval, const float* __restrict__ ang, const int64_t* __restrict__ pos,
                                                               T* __restrict__ dkv, T* __restrict__ dkpe, int H, int kd, int vd, int emb, int batch, float mscale, int interleaved) {
  __shared__ float pe[256];
  const long row = blockIdx.x;
  const int half = emb >> 1, kw = kd + emb, iw = kd + vd;
  if (ang == nullptr) interleaved = 0;
  for (int i = threadIdx.x; i < H * iw; i += 256) {
    const int h = i / iw, c = i % iw;
    dkv[(row * H + h) * iw + c] = c < kd ? dkey[(row * H + h) * kw + c] : dval[(row * H + h) * vd + (c - kd)];
  }
  if ((int)threadIdx.x < emb) {
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc += to_f(dkey[(row * H + h) * kw + kd + threadIdx.x]);
    pe[threadIdx.x] = acc;
  }
  __syncthreads();
  if ((int)threadIdx.x < half) {
    const int i = threadIdx.x;
    const float gl = pe[i], gr = pe[i + half];
    float da = gl, db = gr;
    if (ang != nullptr) {
      const long p = pos != nullptr ? pos[row] : row / batch;
      float cl, sl, cr, sr;
      sincosf(ang[p * emb + i], &sl, &cl);
      sincosf(ang[p * emb + half + i], &sr, &cr);
      da = (gl * cl + gr * sr) * mscale;
      db = (gr * cr - gl * sl) * mscale;
    }
    dkpe[row * emb + (interleaved ? 2 * i : i)] = from_f<T>(da);
    dkpe[row * emb + (interleaved ? 2 * i + 1 : i + half)] = from_f<T>(db);
  }
}

}  // namespace mb200

using namespace mb200;

extern "C" void mb200_indices_to_multihot(const int64_t* idx, const float* probs, uint8_t* map, float* probs_out, long T, int k, int E, cudaStream_t s) {
  if (T > 0) indices_to_multihot_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(idx, probs, map, probs_out, T, k, E);
}
extern "C" void mb200_multihot_probs_grad(const int64_t* idx, const float* g_in, float* g_out, long T, int k, int E, int scatter, cudaStream_t s) {
  if (T > 0) multihot_probs_grad_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(idx, g_in, g_out, T, k, E, scatter);
}
extern "C" void mb200_multihot_to_indices(const uint8_t* map, const float* probs, int64_t* idx, float* probs_out, long T, int k, int E, cudaStream_t s) {
  if (T > 0) multihot_to_indices_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(map, probs, idx, probs_out, T, k, E);
}
extern "C" void mb200_pad_routing_map(const uint8_t* in, uint8_t* out, long T, int E, int multiple, cudaStream_t s) {
  if (T > 0 && E > 0) pad_routing_map_kernel<<<E, 256, 0, s>>>(in, out, T, E, multiple);
}
extern "C" void mb200_moe_aux_loss_fwd(const float* probs, const float* tpe, float* partial, int nblocks, float* loss, long T, int E, float coeff, cudaStream_t s) {
  const long rpb = (T + nblocks - 1) / nblocks;
  aux_loss_partial_kernel<<<nblocks, 256, 0, s>>>(probs, tpe, partial, T, E, rpb);
  aux_loss_final_kernel<<<1, 256, 0, s>>>(partial, nblocks, coeff, loss);
}
extern "C" void mb200_moe_aux_loss_bwd(const float* tpe, const float* gloss, float coeff, float* gprobs, long T, int E, cudaStream_t s) {
  const long n = T * E;
  if (n > 0) aux_loss_bwd_kernel<<<(unsigned)min((n + 255) / 256, (long)4096), 256, 0, s>>>(tpe, gloss, coeff, gprobs, n, E);
}
extern "C" int mb200_mla_rope_inplace(const void* src, void* x, const float* ang, const int64_t* pos, long rows, int H, int nope, int emb, int batch, float mscale, int interleaved, int inverse,
                                      int dtype, cudaStream_t s) {
  if (emb > 64 || (emb & 1)) return -1;
  const long warps = rows * H;
  if (warps == 0) return 0;
  const unsigned grid = (unsigned)((warps + 7) / 8);
  if (dtype == 1) mla_rope_inplace_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)x, ang, pos, rows, H, nope, emb, batch, mscale, interleaved, inverse);
  else if (dtype == 2) mla_rope_inplace_kernel<__half><<<grid, 256, 0, s>>>((const __half*)src, (__half*)x, ang, pos, rows, H, nope, emb, batch, mscale, interleaved, inverse);
  else mla_rope_inplace_kernel<float><<<grid, 256, 0, s>>>((const float*)src, (float*)x, ang, pos, rows, H, nope, emb, batch, mscale, interleaved, inverse);
  return 0;
}
extern "C" int mb200_mla_kv_split(const void* a, const void* b, const float* ang, const int64_t* pos, void* o0, void* o1, long rows, int H, int kd, int vd, int emb, int batch,
                                  float mscale, int interleaved, int backward, int dtype, cudaStream_t s) {
  if (emb > 256 || (emb & 1)) return -1;
  if (rows == 0) return 0;
#define MB200_KV_SPLIT(T)                                                                                                                                   \
  if (backward) mla_kv_split_bwd_kernel<T><<<(unsigned)rows, 256, 0, s>>>((const T*)a, (const T*)b, ang, pos, (T*)o0, (T*)o1, H, kd, vd, emb, batch, mscale, interleaved); \
  else mla_kv_split_fwd_kernel<T><<<(unsigned)rows, 256, 0, s>>>((const T*)a, (const T*)b, ang, pos, (T*)o0, (T*)o1, H, kd, vd, emb, batch, mscale, interleaved);
  if (dtype == 1) { MB200_KV_SPLIT(__nv_bfloat16) } else if (dtype == 2) { MB200_KV_SPLIT(__half) } else { MB200_KV_SPLIT(float) }
#undef MB200_KV_SPLIT
  return 0;
}
