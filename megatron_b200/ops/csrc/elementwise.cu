// SwiGLU (+bias, +per-token MoE prob) forward/backward and RoPE for sm_100a.
// Pure streaming kernels: 16-byte vector accesses, L1-bypassing loads, grid sized to
// a multiple of 148 SMs.  Replace the torch.compile'd "fusions/" bodies and TE's fused
// RoPE.
#include "common.cuh"

namespace mb200 {

__device__ __forceinline__ float sigmoidf_fast(float a) { return 1.f / (1.f + __expf(-a)); }

// y: [rows, 2F] (gate | up), out: [rows, F]
template <typename T>
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const T* __restrict__ y, const T* __restrict__ bias, const float* __restrict__ probs,
                                                           T* __restrict__ out, long rows, int F) {
  constexpr int VN = Vec<T>::N;
  const int vec_per_row = F / VN;
  const long total = rows * vec_per_row;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / vec_per_row;
    const int col = (int)(i - row * vec_per_row) * VN;
    const T* yr = y + row * 2 * F;
    Vec<T> a = ld16_stream(yr + col), b = ld16_stream(yr + F + col);
    float pa = probs ? probs[row] : 1.f;
    Vec<T> o;
    if (bias != nullptr) {
      Vec<T> ba = ld16(bias + col), bb = ld16(bias + F + col);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float av = to_f(a.v[k]) + to_f(ba.v[k]), bv = to_f(b.v[k]) + to_f(bb.v[k]);
        o.v[k] = from_f<T>(av * sigmoidf_fast(av) * bv * pa);
      }
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float av = to_f(a.v[k]), bv = to_f(b.v[k]);
        o.v[k] = from_f<T>(av * sigmoidf_fast(av) * bv * pa);
      }
    }
    st16(out + row * F + col, o);
  }
}

// one CTA per row when probs are present (needs the row-sum for dprobs); grid-stride otherwise
template <typename T, bool HAS_PROBS>
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const T* __restrict__ g, const T* __restrict__ y, const T* __restrict__ bias,
                                                           const float* __restrict__ probs, T* __restrict__ dy, float* __restrict__ dprobs, long rows,
                                                           int F) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[32];
  const int vec_per_row = F / VN;
  if (HAS_PROBS) {
    for (long row = blockIdx.x; row < rows; row += gridDim.x) {
      const float pa = probs[row];
      float acc = 0.f;
      for (int vi = threadIdx.x; vi < vec_per_row; vi += blockDim.x) {
        const int col = vi * VN;
        Vec<T> a = ld16_stream(y + row * 2 * F + col), b = ld16_stream(y + row * 2 * F + F + col), gg = ld16_stream(g + row * F + col);
        Vec<T> da, db;
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          float av = to_f(a.v[k]), bv = to_f(b.v[k]);
          if (bias != nullptr) { av += to_f(bias[col + k]); bv += to_f(bias[F + col + k]); }
          const float s = sigmoidf_fast(av), act = av * s, gv = to_f(gg.v[k]);
          acc += gv * act * bv;
          const float gp = gv * pa;
          da.v[k] = from_f<T>(gp * bv * s * (1.f + av * (1.f - s)));
          db.v[k] = from_f<T>(gp * act);
        }
        st16(dy + row * 2 * F + col, da);
        st16(dy + row * 2 * F + F + col, db);
      }
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) dprobs[row] = acc;
    }
  } else {
    const long total = rows * vec_per_row;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const long row = i / vec_per_row;
      const int col = (int)(i - row * vec_per_row) * VN;
      Vec<T> a = ld16_stream(y + row * 2 * F + col), b = ld16_stream(y + row * 2 * F + F + col), gg = ld16_stream(g + row * F + col);
      Vec<T> da, db;
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float av = to_f(a.v[k]), bv = to_f(b.v[k]);
        if (bias != nullptr) { av += to_f(bias[col + k]); bv += to_f(bias[F + col + k]); }
        const float s = sigmoidf_fast(av), gv = to_f(gg.v[k]);
        da.v[k] = from_f<T>(gv * bv * s * (1.f + av * (1.f - s)));
        db.v[k] = from_f<T>(gv * av * s);
      }
      st16(dy + row * 2 * F + col, da);
      st16(dy + row * 2 * F + F + col, db);
    }
  }
}

// t: [S, B, Hh, D]; freqs: [S, Drot] angles (rotate-half layout: first Drot/2 == second Drot/2).
// out[.., j]        = t[j]*cos(f_j) - t[j+Drot/2]*sin(f_j)          (j <  Drot/2)
// out[.., j+Drot/2] = t[j+Drot/2]*cos(f_j) + t[j]*sin(f_j)
// conj=1 negates sin (backward pass).  Channels >= Drot pass through.
template <typename T>
__global__ void __launch_bounds__(256) rope_kernel(const T* __restrict__ t, const float* __restrict__ freqs, T* __restrict__ out, int S, int B, int Hh,
                                                     int D, int Drot, float mscale, int conj) {
  constexpr int VN = Vec<T>::N;
  const int half = Drot / 2;
  const int vec_half = half / VN;            // vectors in one rotary half
  const int vec_pass = (D - Drot) / VN;      // pass-through vectors
  const int per_head = vec_half + vec_pass;  // work items per (s,b,h)
  const long total = (long)S * B * Hh * per_head;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long head = i / per_head;
    const int w = (int)(i - head * per_head);
    const long s = head / ((long)B * Hh);
    const T* src = t + head * D;
    T* dst = out + head * D;
    if (w < vec_half) {
      const int j = w * VN;
      Vec<T> x1 = ld16_stream(src + j), x2 = ld16_stream(src + half + j);
      Vec<T> o1, o2;
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float sn, cs;
        sincosf(freqs[s * Drot + j + k], &sn, &cs);
        sn *= mscale; cs *= mscale;
        if (conj) sn = -sn;
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

static inline int grid_for(long work_items, int threads) {
  long g = (work_items + threads - 1) / threads;
  const long cap = 148L * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mb200

using namespace mb200;

#define DISPATCH(dtype, ...)                                   \
  switch (dtype) {                                             \
    case kF32: { using T = float; __VA_ARGS__; break; }        \
    case kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    default: { using T = __half; __VA_ARGS__; break; }         \
  }

extern "C" void mb200_swiglu_fwd(const void* y, const void* bias, const float* probs, void* out, long rows, int F, int dtype, cudaStream_t s) {
  DISPATCH(dtype, {
    const long items = rows * (F / Vec<T>::N);
    swiglu_fwd_kernel<T><<<grid_for(items, 256), 256, 0, s>>>((const T*)y, (const T*)bias, probs, (T*)out, rows, F);
  });
}
extern "C" void mb200_swiglu_bwd(const void* g, const void* y, const void* bias, const float* probs, void* dy, float* dprobs, long rows, int F, int dtype,
                                 cudaStream_t s) {
  DISPATCH(dtype, {
    if (probs != nullptr) {
      const int grid = (int)(rows < 148L * 16 ? rows : 148L * 16);
      swiglu_bwd_kernel<T, true><<<grid, 256, 0, s>>>((const T*)g, (const T*)y, (const T*)bias, probs, (T*)dy, dprobs, rows, F);
    } else {
      const long items = rows * (F / Vec<T>::N);
      swiglu_bwd_kernel<T, false><<<grid_for(items, 256), 256, 0, s>>>((const T*)g, (const T*)y, (const T*)bias, probs, (T*)dy, dprobs, rows, F);
    }
  });
}
extern "C" void mb200_rope(const void* t, const float* freqs, void* out, int S, int B, int Hh, int D, int Drot, float mscale, int conj, int dtype,
                           cudaStream_t s) {
  DISPATCH(dtype, {
    const long items = (long)S * B * Hh * ((Drot / 2 + (D - Drot)) / Vec<T>::N);
    rope_kernel<T><<<grid_for(items, 256), 256, 0, s>>>((const T*)t, freqs, (T*)out, S, B, Hh, D, Drot, mscale, conj);
  });
}
