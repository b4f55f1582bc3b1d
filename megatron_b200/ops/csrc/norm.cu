// RMSNorm / LayerNorm forward + backward for sm_100a.
//
// Memory-bound: one pass over x in fwd (row cached in registers), one pass over (gy, x)
// in bwd.  Rows are distributed over a persistent grid; each thread owns NV 16-byte
// column vectors so weight-gradient partials accumulate in registers across all rows a
// CTA visits and are written once per CTA ([nblocks, H] fp32), then tree-reduced by a
// second tiny kernel.  Replaces Apex FusedLayerNorm / TE RMSNorm.
#include "common.cuh"

namespace mb200 {

template <typename T, int NV, bool RMS>
__global__ void __launch_bounds__(256) norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b, T* __restrict__ y,
                                                         float* __restrict__ mu_out, float* __restrict__ rstd_out, int rows, int H, float eps,
                                                         int zero_centered, const T* __restrict__ res, T* __restrict__ res_out) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  // weights are row-invariant: keep in registers
  float wr[NV][VN], br[NV][VN];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * nt + tid) * VN;
    if (col < H) {
      Vec<T> wv = ld16(w + col);
#pragma unroll
      for (int i = 0; i < VN; ++i) wr[v][i] = to_f(wv.v[i]) + (zero_centered ? 1.f : 0.f);
      if (!RMS && b != nullptr) {
        Vec<T> bv = ld16(b + col);
#pragma unroll
        for (int i = 0; i < VN; ++i) br[v][i] = to_f(bv.v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < VN; ++i) br[v][i] = 0.f;
      }
    }
  }
  const float invH = 1.f / (float)H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const T* xr = x + (size_t)row * H;
    float xv[NV][VN];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * nt + tid) * VN;
      if (col < H) {
        Vec<T> t = ld16_stream(xr + col);
        if (res != nullptr) {
          // fused residual add: h = x + residual is rounded to T once, written out (the next residual) and normalised
          Vec<T> rv = ld16_stream(res + (size_t)row * H + col);
#pragma unroll
          for (int i = 0; i < VN; ++i) t.v[i] = from_f<T>(to_f(t.v[i]) + to_f(rv.v[i]));
          st16(res_out + (size_t)row * H + col, t);
        }
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          xv[v][i] = to_f(t.v[i]);
          s1 += xv[v][i];
          s2 += xv[v][i] * xv[v][i];
        }
      }
    }
    float mean = 0.f, rstd;
    if (RMS) {
      s2 = block_sum(s2, red);
      rstd = rsqrtf(s2 * invH + eps);
    } else {
      s1 = block_sum(s1, red);
      mean = s1 * invH;
      float sv = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int col = (v * nt + tid) * VN;
        if (col < H) {
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            const float d = xv[v][i] - mean;
            sv += d * d;
          }
        }
      }
      sv = block_sum(sv, red);
      rstd = rsqrtf(sv * invH + eps);
    }
    if (tid == 0) {
      rstd_out[row] = rstd;
      if (!RMS) mu_out[row] = mean;
    }
    T* yr = y + (size_t)row * H;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * nt + tid) * VN;
      if (col < H) {
        Vec<T> o;
#pragma unroll
        for (int i = 0; i < VN; ++i) o.v[i] = from_f<T>((xv[v][i] - mean) * rstd * wr[v][i] + br[v][i]);
        st16(yr + col, o);
      }
    }
  }
}

// partial layout: [nblocks][2][H]  (slot 0 = dgamma, slot 1 = dbeta)
template <typename T, int NV, bool RMS>
__global__ void __launch_bounds__(256) norm_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ w,
                                                         const float* __restrict__ mu, const float* __restrict__ rstd, T* __restrict__ gx,
                                                         float* __restrict__ partial, int rows, int H, int zero_centered,
                                                         const T* __restrict__ gres) {
  constexpr int VN = Vec<T>::N;
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  float wr[NV][VN], dw[NV][VN], db[NV][VN];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * nt + tid) * VN;
#pragma unroll
    for (int i = 0; i < VN; ++i) dw[v][i] = db[v][i] = 0.f;
    if (col < H) {
      Vec<T> wv = ld16(w + col);
#pragma unroll
      for (int i = 0; i < VN; ++i) wr[v][i] = to_f(wv.v[i]) + (zero_centered ? 1.f : 0.f);
    }
  }
  const float invH = 1.f / (float)H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float r = rstd[row];
    const float m = RMS ? 0.f : mu[row];
    float xh[NV][VN], gg[NV][VN];
    float c1 = 0.f, c2 = 0.f;  // mean(g*w*xhat), mean(g*w)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * nt + tid) * VN;
      if (col < H) {
        Vec<T> xt = ld16_stream(x + (size_t)row * H + col);
        Vec<T> gt = ld16_stream(gy + (size_t)row * H + col);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float g = to_f(gt.v[i]);
          xh[v][i] = (to_f(xt.v[i]) - m) * r;
          dw[v][i] += g * xh[v][i];
          db[v][i] += g;
          gg[v][i] = g * wr[v][i];
          c1 += gg[v][i] * xh[v][i];
          c2 += gg[v][i];
        }
      }
    }
    c1 = block_sum(c1, red) * invH;
    if (!RMS) c2 = block_sum(c2, red) * invH; else c2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * nt + tid) * VN;
      if (col < H) {
        Vec<T> o;
        if (gres != nullptr) {   // gradient arriving through the residual branch is added in the same pass
          Vec<T> gr = ld16_stream(gres + (size_t)row * H + col);
#pragma unroll
          for (int i = 0; i < VN; ++i) o.v[i] = from_f<T>(r * (gg[v][i] - c2 - xh[v][i] * c1) + to_f(gr.v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < VN; ++i) o.v[i] = from_f<T>(r * (gg[v][i] - c2 - xh[v][i] * c1));
        }
        st16(gx + (size_t)row * H + col, o);
      }
    }
  }
  float* pw = partial + (size_t)blockIdx.x * 2 * H;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * nt + tid) * VN;
    if (col < H) {
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        pw[col + i] = dw[v][i];
        if (!RMS) pw[H + col + i] = db[v][i];
      }
    }
  }
}

template <typename T>
__global__ void norm_bwd_reduce_kernel(const float* __restrict__ partial, T* __restrict__ gw, T* __restrict__ gb, int nblocks, int H) {
  // one warp per 32 columns; lanes stride over partial rows, then shuffle-reduce
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rlane = threadIdx.x >> 5, nr = blockDim.x >> 5;
  __shared__ float sm[2][32][33];
  float a = 0.f, b = 0.f;
  if (col < H) {
    for (int r = rlane; r < nblocks; r += nr) {
      a += partial[(size_t)r * 2 * H + col];
      if (gb != nullptr) b += partial[(size_t)r * 2 * H + H + col];
    }
  }
  sm[0][rlane][threadIdx.x & 31] = a;
  sm[1][rlane][threadIdx.x & 31] = b;
  __syncthreads();
  if (rlane == 0 && col < H) {
    float sa = 0.f, sb = 0.f;
    for (int r = 0; r < nr; ++r) {
      sa += sm[0][r][threadIdx.x & 31];
      sb += sm[1][r][threadIdx.x & 31];
    }
    gw[col] = from_f<T>(sa);
    if (gb != nullptr) gb[col] = from_f<T>(sb);
  }
}

template <typename T, bool RMS>
void launch_fwd(const void* x, const void* w, const void* b, void* y, float* mu, float* rstd, int rows, int H, float eps, int zc, cudaStream_t s,
                const void* res = nullptr, void* res_out = nullptr) {
  constexpr int VN = Vec<T>::N;
  const int vecs = H / VN;
  int threads = vecs >= 256 ? 256 : ((vecs + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  const int nv = (vecs + threads - 1) / threads;
  const int grid = rows < 148 * 8 ? rows : 148 * 8;
#define L(NV)                                                                                                                          \
  norm_fwd_kernel<T, NV, RMS><<<grid, threads, 0, s>>>((const T*)x, (const T*)w, (const T*)b, (T*)y, mu, rstd, rows, H, eps, zc, (const T*)res, (T*)res_out)
  if (nv <= 1) L(1); else if (nv <= 2) L(2); else if (nv <= 4) L(4); else if (nv <= 8) L(8); else L(16);
#undef L
}

template <typename T, bool RMS>
void launch_bwd(const void* gy, const void* x, const void* w, const float* mu, const float* rstd, void* gx, float* partial, void* gw, void* gb, int rows,
                int H, int zc, int nblocks, cudaStream_t s, const void* gres = nullptr) {
  constexpr int VN = Vec<T>::N;
  const int vecs = H / VN;
  int threads = vecs >= 256 ? 256 : ((vecs + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  const int nv = (vecs + threads - 1) / threads;
#define L(NV)                                                                                                                          \
  norm_bwd_kernel<T, NV, RMS><<<nblocks, threads, 0, s>>>((const T*)gy, (const T*)x, (const T*)w, mu, rstd, (T*)gx, partial, rows, H, zc, (const T*)gres)
  if (nv <= 1) L(1); else if (nv <= 2) L(2); else if (nv <= 4) L(4); else if (nv <= 8) L(8); else L(16);
#undef L
  norm_bwd_reduce_kernel<T><<<(H + 31) / 32, 256, 0, s>>>(partial, (T*)gw, RMS ? nullptr : (T*)gb, nblocks, H);
}

}  // namespace mb200

using namespace mb200;

#define DISPATCH(dtype, ...)                                   \
  switch (dtype) {                                             \
    case kF32: { using T = float; __VA_ARGS__; break; }        \
    case kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    default: { using T = __half; __VA_ARGS__; break; }         \
  }

extern "C" void mb200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int H, float eps, int zc, int dtype, cudaStream_t s) {
  DISPATCH(dtype, (launch_fwd<T, true>(x, w, nullptr, y, nullptr, rstd, rows, H, eps, zc, s)));
}
extern "C" void mb200_rmsnorm_bwd(const void* gy, const void* x, const void* w, const float* rstd, void* gx, float* partial, void* gw, int rows, int H, int zc,
                                  int dtype, int nblocks, cudaStream_t s) {
  DISPATCH(dtype, (launch_bwd<T, true>(gy, x, w, nullptr, rstd, gx, partial, gw, nullptr, rows, H, zc, nblocks, s)));
}
extern "C" void mb200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mu, float* rstd, int rows, int H, float eps, int zc, int dtype,
                                    cudaStream_t s) {
  DISPATCH(dtype, (launch_fwd<T, false>(x, w, b, y, mu, rstd, rows, H, eps, zc, s)));
}
extern "C" void mb200_layernorm_bwd(const void* gy, const void* x, const void* w, const float* mu, const float* rstd, void* gx, float* partial, void* gw,
                                    void* gb, int rows, int H, int zc, int dtype, int nblocks, cudaStream_t s) {
  DISPATCH(dtype, (launch_bwd<T, false>(gy, x, w, mu, rstd, gx, partial, gw, gb, rows, H, zc, nblocks, s)));
}

// Fused residual-add + RMSNorm: h = x + residual (written to res_out), y = rmsnorm(h) * w.  Backward takes the gradient of y and the gradient that
// reaches h through the residual branch and returns their sum in one pass (reference: TE fused-residual norm / inference fused RS+residual+norm).
extern "C" void mb200_add_rmsnorm_fwd(const void* x, const void* res, const void* w, void* y, void* res_out, float* rstd, int rows, int H, float eps, int zc,
                                      int dtype, cudaStream_t s) {
  DISPATCH(dtype, (launch_fwd<T, true>(x, w, nullptr, y, nullptr, rstd, rows, H, eps, zc, s, res, res_out)));
}
extern "C" void mb200_add_rmsnorm_bwd(const void* gy, const void* gres, const void* h, const void* w, const float* rstd, void* gx, float* partial, void* gw, int rows,
                                      int H, int zc, int dtype, int nblocks, cudaStream_t s) {
  DISPATCH(dtype, (launch_bwd<T, true>(gy, h, w, nullptr, rstd, gx, partial, gw, nullptr, rows, H, zc, nblocks, s, gres)));
}
