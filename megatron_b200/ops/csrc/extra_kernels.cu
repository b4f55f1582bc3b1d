// Second batch of memory-bound sm_100a kernels:
//   * RoPE for packed (thd) tokens and in-place RoPE on the mixed QKV projection output (reference: TE fused RoPE thd / fused-QKV variants)
//   * depthwise causal conv1d (+SiLU) forward / backward for Mamba-2 and gated-delta-net (reference ssm/ops/common/causal_conv1d_triton.py)
//   * Mamba-2 SSD inter-chunk state passing (forward + reverse scan) and the single-token state update (reference ssm/ops/mamba2/ssd_state_passing.py,
//     selective_state_update.py) — the GEMM-shaped intra-chunk work stays on the tensor cores through batched GEMMs
//   * MXFP8 (OCP microscaling: 32-element blocks, E8M0 shared exponent, E4M3 payload) quantise / dequantise (reference quantization/mxfp8_quantize.py)
// All of them are bandwidth-bound: 16-byte accesses, one pass over the data, fp32 math in registers.
#include <cuda_fp4.h>
#include <cuda_fp8.h>

#include "common.cuh"

namespace mb200 {

static inline int grid_cap(long items, int threads, int per_sm = 16) {
  long g = (items + threads - 1) / threads;
  const long cap = 148L * per_sm;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------ RoPE variants
// t: [T, Hh, D] packed tokens, pos[T] = position of the token inside its sequence.
template <typename T>
__global__ void __launch_bounds__(256) rope_pos_kernel(const T* __restrict__ t, const float* __restrict__ freqs, const int* __restrict__ pos, T* __restrict__ out,
                                                         long tokens, int Hh, int D, int Drot, float mscale, int conj) {
  constexpr int VN = Vec<T>::N;
  const int half = Drot / 2, vec_half = half / VN, vec_pass = (D - Drot) / VN, per_head = vec_half + vec_pass;
  const long total = tokens * Hh * per_head;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long head = i / per_head;
    const int w = (int)(i - head * per_head);
    const long s = pos[head / Hh];
    const T* src = t + head * D;
    T* dst = out + head * D;
    if (w < vec_half) {
      const int j = w * VN;
      Vec<T> x1 = ld16_stream(src + j), x2 = ld16_stream(src + half + j), o1, o2;
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float sn, cs;
        sincosf(freqs[s * Drot + j + k], &sn, &cs);
        sn *= conj ? -mscale : mscale;
        cs *= mscale;
        const float a = to_f(x1.v[k]), b = to_f(x2.v[k]);
        o1.v[k] = from_f<T>(a * cs - b * sn);
        o2.v[k] = from_f<T>(b * cs + a * sn);
      }
      st16(dst + j, o1);
      st16(dst + half + j, o2);
    } else {
      const int j = Drot + (w - vec_half) * VN;
      st16(dst + j, ld16_stream(src + j));
    }
  }
}

// qkv: [S, B, NG, (QPG + 2) * D] — per query group QPG query heads, one key head, one value head.  Rotates the query and key heads in place
// (out may alias qkv); value heads are copied only when out != qkv.  The attention kernels consume the strided q / k / v views directly,
// so no split copies are needed after this.
template <typename T>
__global__ void __launch_bounds__(256) rope_qkv_kernel(const T* qkv, const float* __restrict__ freqs, T* out, int S, int B, int NG,
                                                         int QPG, int D, int Drot, float mscale, int conj) {
  constexpr int VN = Vec<T>::N;
  const int half = Drot / 2, vec_half = half / VN, vec_pass = (D - Drot) / VN, per_head = vec_half + vec_pass;
  const int slots = QPG + 2;
  const bool inplace = (out == qkv);
  const long total = (long)S * B * NG * slots * per_head;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long head = i / per_head;              // (s, b, g, slot)
    const int w = (int)(i - head * per_head);
    const int slot = (int)(head % slots);
    const long s = head / ((long)B * NG * slots);
    const T* src = qkv + head * D;
    T* dst = out + head * D;
    if (slot == slots - 1 || w >= vec_half) {    // value head or pass-through tail
      if (!inplace) {
        if (slot == slots - 1) {
          // a value head owns per_head work items but D / VN vectors: item w copies vector w, the first vec_half items also copy the second rotary half
          const int j = (w < vec_half) ? w * VN : Drot + (w - vec_half) * VN;
          st16(dst + j, ld16_stream(src + j));
          if (w < vec_half) st16(dst + half + j, ld16_stream(src + half + j));
        } else {
          const int j = Drot + (w - vec_half) * VN;
          st16(dst + j, ld16_stream(src + j));
        }
      }
      continue;
    }
    const int j = w * VN;
    Vec<T> x1 = ld16(src + j), x2 = ld16(src + half + j), o1, o2;
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      float sn, cs;
      sincosf(freqs[s * Drot + j + k], &sn, &cs);
      sn *= conj ? -mscale : mscale;
      cs *= mscale;
      const float a = to_f(x1.v[k]), b = to_f(x2.v[k]);
      o1.v[k] = from_f<T>(a * cs - b * sn);
      o2.v[k] = from_f<T>(b * cs + a * sn);
    }
    st16(dst + j, o1);
    st16(dst + half + j, o2);
  }
}

// ------------------------------------------------------------------------------------------------ causal conv1d
__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ float dsilu_f(float v) {
  const float sg = 1.f / (1.f + __expf(-v));
  return sg * (1.f + v * (1.f - sg));
}

// x, y: [b, d, l] (l contiguous); w: [d, K]; left: [b, d, K-1] or null.  One CTA per (b, d) row.
template <typename T, int K>
__global__ void __launch_bounds__(256) conv1d_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ bias, const T* __restrict__ left,
                                                           T* __restrict__ y, int d, int l, int act) {
  const long row = blockIdx.x;
  const int ch = (int)(row % d);
  float wr[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wr[j] = to_f(w[ch * K + j]);
  const float bs = bias != nullptr ? to_f(bias[ch]) : 0.f;
  const T* xr = x + row * l;
  const T* lf = left != nullptr ? left + row * (K - 1) : nullptr;
  T* yr = y + row * l;
  for (int t = threadIdx.x; t < l; t += blockDim.x) {
    float acc = bs;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int u = t - (K - 1) + j;
      const float xv = u >= 0 ? to_f(xr[u]) : (lf != nullptr ? to_f(lf[u + (K - 1)]) : 0.f);
      acc = fmaf(wr[j], xv, acc);
    }
    yr[t] = from_f<T>(act ? silu_f(acc) : acc);
  }
}

// gx: [b, d, l]; gw_acc: [d, K] fp32 (zero-initialised, atomically accumulated); gb_acc: [d] fp32 or null; gleft: [b, d, K-1] or null.
template <typename T, int K>
__global__ void __launch_bounds__(256) conv1d_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ bias,
                                                           const T* __restrict__ left, T* __restrict__ gx, T* __restrict__ gleft, float* __restrict__ gw_acc,
                                                           float* __restrict__ gb_acc, int d, int l, int act) {
  __shared__ float red[32];
  const long row = blockIdx.x;
  const int ch = (int)(row % d);
  float wr[K], dw[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    wr[j] = to_f(w[ch * K + j]);
    dw[j] = 0.f;
  }
  float db = 0.f;
  const float bs = bias != nullptr ? to_f(bias[ch]) : 0.f;
  const T* xr = x + row * l;
  const T* gr = gy + row * l;
  const T* lf = left != nullptr ? left + row * (K - 1) : nullptr;
  auto xat = [&](int u) -> float { return u >= 0 ? to_f(xr[u]) : (lf != nullptr ? to_f(lf[u + (K - 1)]) : 0.f); };
  // g_pre(t) = gy(t) * act'(pre(t)); recomputed where needed (K is 2..4, the window fits in registers / L1)
  auto gpre = [&](int t) -> float {
    if (t >= l) return 0.f;
    float g = to_f(gr[t]);
    if (act) {
      float acc = bs;
#pragma unroll
      for (int j = 0; j < K; ++j) acc = fmaf(wr[j], xat(t - (K - 1) + j), acc);
      g *= dsilu_f(acc);
    }
    return g;
  };
  // positions -(K-1) .. l-1: negative ones are the gradient of the carried-in state
  for (int t = (int)threadIdx.x - (K - 1); t < l; t += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int o = t + (K - 1) - j;          // output position that read x[t] through tap j
      if (o >= 0) acc = fmaf(wr[j], gpre(o), acc);
    }
    if (t >= 0) {
      gx[row * l + t] = from_f<T>(acc);
      const float g = gpre(t);
      db += g;
#pragma unroll
      for (int j = 0; j < K; ++j) dw[j] = fmaf(g, xat(t - (K - 1) + j), dw[j]);
    } else if (gleft != nullptr) {
      gleft[row * (K - 1) + t + (K - 1)] = from_f<T>(acc);
    }
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float s = block_sum(dw[j], red);
    if (threadIdx.x == 0) atomicAdd(gw_acc + ch * K + j, s);
  }
  if (gb_acc != nullptr) {
    const float s = block_sum(db, red);
    if (threadIdx.x == 0) atomicAdd(gb_acc + ch, s);
  }
}

// ------------------------------------------------------------------------------------------------ SSD state passing
// states: [b, c, h, E] (E = p*n) fp32 chunk contributions; decay: [b, h, c] log-decay of each chunk; init: [b, h, E] or null.
// prev[b, z, h, :] = state entering chunk z; final[b, h, :] = state after the last chunk.  One thread per (b, h, e); the chunk loop is sequential.
__global__ void __launch_bounds__(256) ssd_state_fwd_kernel(const float* __restrict__ states, const float* __restrict__ decay, const float* __restrict__ init,
                                                              float* __restrict__ prev, float* __restrict__ fin, int b, int c, int h, int E) {
  const long total = (long)b * h * E;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E);
    const long bh = i / E;
    const int hh = (int)(bh % h), bb = (int)(bh / h);
    float s = init != nullptr ? init[i] : 0.f;
    const float* dz = decay + ((long)bb * h + hh) * c;
    for (int z = 0; z < c; ++z) {
      const long off = (((long)bb * c + z) * h + hh) * E + e;
      prev[off] = s;
      s = fmaf(__expf(dz[z]), s, states[off]);
    }
    fin[i] = s;
  }
}

// Reverse scan.  g_prev: [b, c, h, E]; g_fin: [b, h, E] or null; prev: saved from forward.  Outputs g_states [b, c, h, E], g_init [b, h, E], g_decay [b, h, c].
// One CTA per (b, h): the decay gradient needs a reduction over E for every chunk.
__global__ void __launch_bounds__(256) ssd_state_bwd_kernel(const float* __restrict__ g_prev, const float* __restrict__ g_fin, const float* __restrict__ prev,
                                                              const float* __restrict__ decay, float* __restrict__ g_states, float* __restrict__ g_init,
                                                              float* __restrict__ g_decay, int b, int c, int h, int E) {
  constexpr int EPT = 32;                // elements per thread held in registers across the chunk walk (256 x 32 = one 64 x 128 state)
  extern __shared__ float sm[];          // [c] decay-gradient accumulators
  __shared__ float red[32];
  const int bh = blockIdx.x, hh = bh % h, bb = bh / h;
  const float* dz = decay + (long)bh * c;
  for (int z = threadIdx.x; z < c; z += blockDim.x) sm[z] = 0.f;
  __syncthreads();
  for (int e0 = 0; e0 < E; e0 += EPT * (int)blockDim.x) {
    float G[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int e = e0 + k * (int)blockDim.x + (int)threadIdx.x;
      G[k] = (e < E && g_fin != nullptr) ? g_fin[(long)bh * E + e] : 0.f;
    }
    for (int z = c - 1; z >= 0; --z) {
      const float ed = __expf(dz[z]);
      const long rowoff = (((long)bb * c + z) * h + hh) * E;
      float contrib = 0.f;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int e = e0 + k * (int)blockDim.x + (int)threadIdx.x;
        if (e < E) {
          g_states[rowoff + e] = G[k];
          contrib = fmaf(G[k], prev[rowoff + e], contrib);
          G[k] = fmaf(ed, G[k], g_prev[rowoff + e]);
        }
      }
      const float tot = block_sum(contrib * ed, red);
      if (threadIdx.x == 0) sm[z] += tot;
    }
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int e = e0 + k * (int)blockDim.x + (int)threadIdx.x;
      if (e < E) g_init[(long)bh * E + e] = G[k];
    }
  }
  __syncthreads();
  for (int z = threadIdx.x; z < c; z += blockDim.x) g_decay[(long)bh * c + z] = sm[z];
}

// Single-token update (decode): state [b, h, p, n] fp32 in place; x [b, h, p]; dt [b, h]; A [h]; B, C [b, g, n]; D [h] or null; y [b, h, p].
// One CTA per (b, h); a warp owns rows p = warp, warp + nw, ...; lanes stride over n.
template <typename T>
__global__ void __launch_bounds__(256) ssd_step_kernel(float* __restrict__ state, const T* __restrict__ x, const float* __restrict__ dt, const float* __restrict__ A,
                                                         const T* __restrict__ B, const T* __restrict__ C, const float* __restrict__ D, T* __restrict__ y, int h,
                                                         int g, int p, int n) {
  const int bh = blockIdx.x, hh = bh % h, bb = bh / h, grp = hh / (h / g);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float dtv = dt[bh], dA = __expf(A[hh] * dtv);
  const T* Br = B + ((long)bb * g + grp) * n;
  const T* Cr = C + ((long)bb * g + grp) * n;
  const float Dv = D != nullptr ? D[hh] : 0.f;
  for (int pp = warp; pp < p; pp += nw) {
    const float xv = to_f(x[(long)bh * p + pp]);
    float* sr = state + ((long)bh * p + pp) * n;
    float acc = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float s = fmaf(sr[j], dA, dtv * to_f(Br[j]) * xv);
      sr[j] = s;
      acc = fmaf(s, to_f(Cr[j]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) y[(long)bh * p + pp] = from_f<T>(acc + Dv * xv);
  }
}

// ------------------------------------------------------------------------------------------------ MXFP8
// x: [rows, K] bf16 (K % 32 == 0) → q: [rows, K] e4m3 bytes, sf: [rows, K/32] E8M0 bytes.  A thread owns 8 elements; the 4 threads of a 32-element block
// share the block maximum through two shuffles.  Scale = 2^(floor(log2(amax)) - 8) (8 = emax of E4M3), so the largest element lands in [256, 448].
__global__ void __launch_bounds__(256) mxfp8_quant_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf, long nvec, int K) {
  // warp-uniform trip count: the block maximum is exchanged with full-mask shuffles, so no lane may leave the loop early
  for (long base = blockIdx.x * (long)blockDim.x; base < nvec; base += (long)gridDim.x * blockDim.x) {
    const long i = base + threadIdx.x;
    const bool valid = i < nvec;
    float f[8], amax = 0.f;
    if (valid) {
      Vec<__nv_bfloat16> v = ld16_stream(x + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = __bfloat162float(v.v[k]);
        amax = fmaxf(amax, fabsf(f[k]));
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    int e = amax > 0.f ? (int)((__float_as_uint(amax) >> 23) & 0xff) - 127 - 8 : -127;     // floor(log2(amax)) - emax
    e = e < -127 ? -127 : (e > 127 ? 127 : e);
    const float inv = __uint_as_float((uint32_t)(127 - e) << 23);                           // 2^-e  (e = -127 → 2^254·… guarded by amax == 0 ⇒ all zeros)
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= (uint32_t)__nv_cvt_float_to_fp8(f[k] * inv, __NV_SATFINITE, __NV_E4M3) << (8 * k);
      hi |= (uint32_t)__nv_cvt_float_to_fp8(f[4 + k] * inv, __NV_SATFINITE, __NV_E4M3) << (8 * k);
    }
    if (valid) {
      *reinterpret_cast<uint2*>(q + i * 8) = make_uint2(lo, hi);
      if ((i & 3) == 0) sf[i >> 2] = (uint8_t)(e + 127);
    }
  }
}

__global__ void __launch_bounds__(256) mxfp8_dequant_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ sf, __nv_bfloat16* __restrict__ out, long nvec) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    const uint2 u = *reinterpret_cast<const uint2*>(q + i * 8);
    const float sc = __uint_as_float((uint32_t)sf[i >> 2] << 23);
    Vec<__nv_bfloat16> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint8_t b8 = (uint8_t)(((k < 4 ? u.x : u.y) >> (8 * (k & 3))) & 0xff);
      const __half_raw hr = __nv_cvt_fp8_to_halfraw(b8, __NV_E4M3);
      o.v[k] = __float2bfloat16_rn(__half2float(__half(hr)) * sc);
    }
    st16(out + i * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------ NVFP4
// x: [rows, K] bf16 (K % 16 == 0) → q: [rows, K/2] two E2M1 codes per byte (even k in the low nibble), sf: [rows, K/16] UE4M3 block scales.
// tscale (device scalar) = amax(x) / (6 * 448) is the tensor-level scale; block scale = amax_block / (6 * tscale) rounded to E4M3 (≥ 2^-9);
// payload = x / (block scale * tscale) rounded to nearest-even E2M1.  One thread per 16-element block.
__global__ void __launch_bounds__(256) nvfp4_quant_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ tscale, uint8_t* __restrict__ q,
                                                            uint8_t* __restrict__ sf, long nblocks) {
  const float inv_t = 1.f / __ldg(tscale);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nblocks; i += (long)gridDim.x * blockDim.x) {
    Vec<__nv_bfloat16> v0 = ld16_stream(x + i * 16), v1 = ld16_stream(x + i * 16 + 8);
    float f[16], amax = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = __bfloat162float(v0.v[k]) * inv_t;
      f[8 + k] = __bfloat162float(v1.v[k]) * inv_t;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) amax = fmaxf(amax, fabsf(f[k]));
    const float want = fmaxf(amax * (1.f / 6.f), 0.001953125f);                       // 2^-9: smallest scale the reference allows
    const uint8_t sb = (uint8_t)__nv_cvt_float_to_fp8(want, __NV_SATFINITE, __NV_E4M3);
    const float sc = __half2float(__half(__nv_cvt_fp8_to_halfraw(sb, __NV_E4M3)));
    const float inv = 1.f / sc;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= (uint32_t)(__nv_cvt_float2_to_fp4x2(make_float2(f[2 * k] * inv, f[2 * k + 1] * inv), __NV_E2M1, cudaRoundNearest) & 0xff) << (8 * k);
      hi |= (uint32_t)(__nv_cvt_float2_to_fp4x2(make_float2(f[8 + 2 * k] * inv, f[8 + 2 * k + 1] * inv), __NV_E2M1, cudaRoundNearest) & 0xff) << (8 * k);
    }
    *reinterpret_cast<uint2*>(q + i * 8) = make_uint2(lo, hi);
    sf[i] = sb;
  }
}

// ------------------------------------------------------------------------------------------------ scale + mask + softmax (unfused attention path)
// x, y: [rows, sk] with rows = b * h * sq; mask: uint8 [b, 1, sq, sk] (non-zero = masked with -10000 like the reference) or null; causal aligns the diagonal to the
// bottom-right (sk >= sq).  One CTA per row, three passes over a row that is L1/L2 resident (this is the arbitrary-mask fallback, not the flash path).
template <typename T>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, T* __restrict__ y, int h, int sq, int sk, float scale,
                                                            int causal) {
  __shared__ float red[32];
  const long row = blockIdx.x;
  const int qi = (int)(row % sq);
  const long bi = row / ((long)h * sq);
  const T* xr = x + row * sk;
  const uint8_t* mr = mask != nullptr ? mask + (bi * sq + qi) * (long)sk : nullptr;
  const int limit = causal ? qi + (sk - sq) : sk - 1;        // last visible key
  auto val = [&](int j) -> float {
    if (j > limit) return -INFINITY;
    float v = to_f(xr[j]) * scale;
    return (mr != nullptr && mr[j]) ? -10000.f : v;
  };
  float m = -INFINITY;
  for (int j = threadIdx.x; j < sk; j += blockDim.x) m = fmaxf(m, val(j));
  m = block_max(m, red);
  float sum = 0.f;
  for (int j = threadIdx.x; j < sk; j += blockDim.x) sum += __expf(val(j) - m);
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  T* yr = y + row * sk;
  for (int j = threadIdx.x; j < sk; j += blockDim.x) yr[j] = from_f<T>(__expf(val(j) - m) * inv);
}

// gx = scale * y * (gy - Σ gy·y)
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gx, int sk, float scale) {
  __shared__ float red[32];
  const long row = blockIdx.x;
  const T* gr = gy + row * sk;
  const T* yr = y + row * sk;
  float dot = 0.f;
  for (int j = threadIdx.x; j < sk; j += blockDim.x) dot += to_f(gr[j]) * to_f(yr[j]);
  dot = block_sum(dot, red);
  for (int j = threadIdx.x; j < sk; j += blockDim.x) gx[row * sk + j] = from_f<T>(scale * to_f(yr[j]) * (to_f(gr[j]) - dot));
}

// ------------------------------------------------------------------------------------------------ squared ReLU and quick-GeGLU (vectorised elementwise)
// mode 0: y = relu(x)^2 over n elements.   mode 1: x = [a | b] halves of width F per row, y = a·σ(1.702 a) · b
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long rows, int F, int mode) {
  constexpr int VN = Vec<T>::N;
  const long nvec = rows * F / VN;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    const long r = (i * VN) / F;
    const int c = (int)((i * VN) % F);
    Vec<T> o;
    if (mode == 0) {
      Vec<T> a = ld16_stream(x + i * VN);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float v = fmaxf(to_f(a.v[k]), 0.f);
        o.v[k] = from_f<T>(v * v);
      }
    } else {
      Vec<T> a = ld16_stream(x + r * 2 * F + c), b = ld16_stream(x + r * 2 * F + F + c);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float av = to_f(a.v[k]);
        o.v[k] = from_f<T>(av / (1.f + __expf(-1.702f * av)) * to_f(b.v[k]));
      }
    }
    st16(y + i * VN, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ g, const T* __restrict__ x, T* __restrict__ gx, long rows, int F, int mode) {
  constexpr int VN = Vec<T>::N;
  const long nvec = rows * F / VN;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    const long r = (i * VN) / F;
    const int c = (int)((i * VN) % F);
    Vec<T> gv = ld16_stream(g + i * VN);
    if (mode == 0) {
      Vec<T> a = ld16_stream(x + i * VN), o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = from_f<T>(2.f * fmaxf(to_f(a.v[k]), 0.f) * to_f(gv.v[k]));
      st16(gx + i * VN, o);
    } else {
      Vec<T> a = ld16_stream(x + r * 2 * F + c), b = ld16_stream(x + r * 2 * F + F + c), oa, ob;
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float av = to_f(a.v[k]), bv = to_f(b.v[k]), gg = to_f(gv.v[k]);
        const float sg = 1.f / (1.f + __expf(-1.702f * av));
        oa.v[k] = from_f<T>(gg * bv * sg * (1.f + 1.702f * av * (1.f - sg)));
        ob.v[k] = from_f<T>(gg * av * sg);
      }
      st16(gx + r * 2 * F + c, oa);
      st16(gx + r * 2 * F + F + c, ob);
    }
  }
}

}  // namespace mb200

using namespace mb200;

#define DISPATCH_T(dtype, ...)                                   \
  switch (dtype) {                                               \
    case kF32: { using T = float; __VA_ARGS__; break; }          \
    case kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    default: { using T = __half; __VA_ARGS__; break; }           \
  }

extern "C" void mb200_rope_pos(const void* t, const float* freqs, const int* pos, void* out, long tokens, int Hh, int D, int Drot, float mscale, int conj, int dtype,
                               cudaStream_t s) {
  DISPATCH_T(dtype, {
    const long items = tokens * Hh * ((Drot / 2 + (D - Drot)) / Vec<T>::N);
    rope_pos_kernel<T><<<grid_cap(items, 256), 256, 0, s>>>((const T*)t, freqs, pos, (T*)out, tokens, Hh, D, Drot, mscale, conj);
  });
}

extern "C" void mb200_rope_qkv(const void* qkv, const float* freqs, void* out, int S, int B, int NG, int QPG, int D, int Drot, float mscale, int conj, int dtype,
                               cudaStream_t s) {
  DISPATCH_T(dtype, {
    const long items = (long)S * B * NG * (QPG + 2) * ((Drot / 2 + (D - Drot)) / Vec<T>::N);
    rope_qkv_kernel<T><<<grid_cap(items, 256), 256, 0, s>>>((const T*)qkv, freqs, (T*)out, S, B, NG, QPG, D, Drot, mscale, conj);
  });
}

extern "C" int mb200_conv1d_fwd(const void* x, const void* w, const void* bias, const void* left, void* y, long rows, int d, int l, int K, int act, int dtype,
                                cudaStream_t s) {
  if (K < 2 || K > 4) return 1;
  DISPATCH_T(dtype, {
    if (K == 2) conv1d_fwd_kernel<T, 2><<<(unsigned)rows, 256, 0, s>>>((const T*)x, (const T*)w, (const T*)bias, (const T*)left, (T*)y, d, l, act);
    else if (K == 3) conv1d_fwd_kernel<T, 3><<<(unsigned)rows, 256, 0, s>>>((const T*)x, (const T*)w, (const T*)bias, (const T*)left, (T*)y, d, l, act);
    else conv1d_fwd_kernel<T, 4><<<(unsigned)rows, 256, 0, s>>>((const T*)x, (const T*)w, (const T*)bias, (const T*)left, (T*)y, d, l, act);
  });
  return 0;
}

extern "C" int mb200_conv1d_bwd(const void* gy, const void* x, const void* w, const void* bias, const void* left, void* gx, void* gleft, float* gw_acc, float* gb_acc,
                                long rows, int d, int l, int K, int act, int dtype, cudaStream_t s) {
  if (K < 2 || K > 4) return 1;
  DISPATCH_T(dtype, {
    if (K == 2) conv1d_bwd_kernel<T, 2><<<(unsigned)rows, 256, 0, s>>>((const T*)gy, (const T*)x, (const T*)w, (const T*)bias, (const T*)left, (T*)gx, (T*)gleft, gw_acc, gb_acc, d, l, act);
    else if (K == 3) conv1d_bwd_kernel<T, 3><<<(unsigned)rows, 256, 0, s>>>((const T*)gy, (const T*)x, (const T*)w, (const T*)bias, (const T*)left, (T*)gx, (T*)gleft, gw_acc, gb_acc, d, l, act);
    else conv1d_bwd_kernel<T, 4><<<(unsigned)rows, 256, 0, s>>>((const T*)gy, (const T*)x, (const T*)w, (const T*)bias, (const T*)left, (T*)gx, (T*)gleft, gw_acc, gb_acc, d, l, act);
  });
  return 0;
}

extern "C" void mb200_ssd_state_fwd(const float* states, const float* decay, const float* init, float* prev, float* fin, int b, int c, int h, int E, cudaStream_t s) {
  ssd_state_fwd_kernel<<<grid_cap((long)b * h * E, 256), 256, 0, s>>>(states, decay, init, prev, fin, b, c, h, E);
}

extern "C" void mb200_ssd_state_bwd(const float* g_prev, const float* g_fin, const float* prev, const float* decay, float* g_states, float* g_init, float* g_decay,
                                    int b, int c, int h, int E, cudaStream_t s) {
  ssd_state_bwd_kernel<<<b * h, 256, c * sizeof(float), s>>>(g_prev, g_fin, prev, decay, g_states, g_init, g_decay, b, c, h, E);
}

extern "C" void mb200_ssd_step(float* state, const void* x, const float* dt, const float* A, const void* B, const void* C, const float* D, void* y, int b, int h, int g,
                               int p, int n, int dtype, cudaStream_t s) {
  DISPATCH_T(dtype, (ssd_step_kernel<T><<<b * h, 256, 0, s>>>(state, (const T*)x, dt, A, (const T*)B, (const T*)C, D, (T*)y, h, g, p, n)));
}

extern "C" void mb200_mxfp8_quant(const void* x, void* q, void* sf, long rows, int K, cudaStream_t s) {
  const long nvec = rows * K / 8;
  mxfp8_quant_kernel<<<grid_cap(nvec, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (uint8_t*)q, (uint8_t*)sf, nvec, K);
}

extern "C" void mb200_mxfp8_dequant(const void* q, const void* sf, void* out, long rows, int K, cudaStream_t s) {
  const long nvec = rows * K / 8;
  mxfp8_dequant_kernel<<<grid_cap(nvec, 256), 256, 0, s>>>((const uint8_t*)q, (const uint8_t*)sf, (__nv_bfloat16*)out, nvec);
}

extern "C" void mb200_nvfp4_quant(const void* x, const float* tscale, void* q, void* sf, long rows, int K, cudaStream_t s) {
  const long nblocks = rows * K / 16;
  nvfp4_quant_kernel<<<grid_cap(nblocks, 256), 256, 0, s>>>((const __nv_bfloat16*)x, tscale, (uint8_t*)q, (uint8_t*)sf, nblocks);
}

extern "C" void mb200_softmax_fwd(const void* x, const void* mask, void* y, long rows, int h, int sq, int sk, float scale, int causal, int dtype, cudaStream_t s) {
  DISPATCH_T(dtype, (softmax_fwd_kernel<T><<<(unsigned)rows, 256, 0, s>>>((const T*)x, (const uint8_t*)mask, (T*)y, h, sq, sk, scale, causal)));
}
extern "C" void mb200_softmax_bwd(const void* gy, const void* y, void* gx, long rows, int sk, float scale, int dtype, cudaStream_t s) {
  DISPATCH_T(dtype, (softmax_bwd_kernel<T><<<(unsigned)rows, 256, 0, s>>>((const T*)gy, (const T*)y, (T*)gx, sk, scale)));
}
extern "C" void mb200_act_fwd(const void* x, void* y, long rows, int F, int mode, int dtype, cudaStream_t s) {
  DISPATCH_T(dtype, (act_fwd_kernel<T><<<grid_cap(rows * F / Vec<T>::N, 256), 256, 0, s>>>((const T*)x, (T*)y, rows, F, mode)));
}
extern "C" void mb200_act_bwd(const void* g, const void* x, void* gx, long rows, int F, int mode, int dtype, cudaStream_t s) {
  DISPATCH_T(dtype, (act_bwd_kernel<T><<<grid_cap(rows * F / Vec<T>::N, 256), 256, 0, s>>>((const T*)g, (const T*)x, (T*)gx, rows, F, mode)));
}
