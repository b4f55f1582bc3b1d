// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA → smem ring → tcgen05.mma → TMEM →
// epilogue.  One CTA per SM, 128x256 output tile, BLOCK_K = 64 (one 128-byte swizzle row),
// 4-stage smem pipeline, 2 accumulator buffers in TMEM (2 x 256 columns) so the epilogue of
// tile i overlaps the main loop of tile i+1.
//
//   warp 0 : TMA producer (one elected lane)
//   warp 1 : MMA issuer   (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue (tcgen05.ld → convert → global store, optional fp32 accumulate)
//
// Layouts (all row-major tensors, bf16):
//   0 NT : C[M,N] = A[M,K]  · B[N,K]^T     forward          (A, B K-major)
//   1 NN : C[M,N] = A[M,K]  · B[K,N]       dgrad            (B MN-major)
//   2 TN : C[M,N] = A[K,M]^T · B[K,N]      wgrad            (A, B MN-major), C may be fp32 with
//                                                            beta = 1 (main_grad accumulation)
// Replaces cuBLAS-through-torch.matmul and Apex `wgrad_gemm_accum_fp32`.
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include <stdio.h>
#include <unordered_map>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace mb200 {
using namespace ptx;

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  int M, N, K;
  int ldc;
  int accumulate;  // C += A*B
  int group_m;     // rasterisation group (in 128-row blocks)
};

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int& m_blk, int& n_blk) {
  // grouped rasterisation: walk `group_m` row-blocks down before moving to the next column
  const int per_group = group_m * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int gsz = min(group_m, tiles_m - first_m);
  const int r = tile - g * per_group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

template <bool A_MN, bool B_MN, bool C_F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ Cptr, GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)  // box {64 m, 64 k} per chunk
              tma_load_2d(sa + c * (BK * 128), &tmap_a, &full_bar[stage], m_blk * BM + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);  // box {64 k, 256 n}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (BK * 128), &tmap_b, &full_bar[stage], n_blk * BN + c * 64, kb * BK);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major : rows are 128 B, 8-row atoms 1024 B apart, +32 B per UMMA_K
            // MN-major: 64-element chunks BK*128 B apart (LBO), 8-k-row atoms 1024 B apart (SBO), +2048 B per UMMA_K
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, BK * 128, 1024) : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, BK * 128, 1024) : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete → epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue =======================================
    const int ew = warp - 4;  // == warp % 4 → TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + ew * 32 + lane;
      const uint32_t t_base = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_base + c * 32, r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // all of this warp's TMEM reads are done → hand the accumulator back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int col0 = n_blk * BN + c * 32;
        if (!row_ok || col0 >= p.N) continue;
        if (C_F32) {
          float* crow = reinterpret_cast<float*>(Cptr) + (size_t)row * p.ldc + col0;
          if (col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (p.accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(crow + j);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              *reinterpret_cast<float4*>(crow + j) = v;
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) crow[j] = __uint_as_float(r[j]) + (p.accumulate ? crow[j] : 0.f);
          }
        } else {
          __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(Cptr) + (size_t)row * p.ldc + col0;
          if (col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(r[j + q]);
              if (p.accumulate) {
                const uint4 o = *reinterpret_cast<const uint4*>(crow + j);
                const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(&o);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] += __bfloat162float(ob[q]);
              }
              uint4 o;
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int q = 0; q < 4; ++q) ob[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
              *reinterpret_cast<uint4*>(crow + j) = o;
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) {
              float f = __uint_as_float(r[j]);
              if (p.accumulate) f += __bfloat162float(crow[j]);
              crow[j] = __float2bfloat16_rn(f);
            }
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// row-major bf16 matrix [rows, cols]; box = {box_cols (inner, <= 64), box_rows}; 128B swizzle
static bool make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <bool A_MN, bool B_MN, bool C_F32>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, void* C, const GemmParams& p, cudaStream_t s) {
  auto kern = gemm_bf16_kernel<A_MN, B_MN, C_F32>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, C, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace mb200

using namespace mb200;

// returns 0 on success, <0 on failure (caller raises)
extern "C" int mb200_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int layout, int accumulate, int c_dtype, cudaStream_t s) {
  CUtensorMap ta, tb;
  bool ok = true;
  const bool a_mn = layout == 2, b_mn = layout != 0;
  // A: K-major → stored [M, K];  MN-major → stored [K, M]
  ok &= a_mn ? make_tmap_bf16(&ta, A, K, M, 64, BK) : make_tmap_bf16(&ta, A, M, K, BK, BM);
  ok &= b_mn ? make_tmap_bf16(&tb, B, K, N, 64, BK) : make_tmap_bf16(&tb, B, N, K, BK, BN);
  if (!ok) return -1;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = N; p.accumulate = accumulate; p.group_m = 8;
  const bool f32 = c_dtype == kF32;
  if (layout == 0) return f32 ? launch<false, false, true>(ta, tb, C, p, s) : launch<false, false, false>(ta, tb, C, p, s);
  if (layout == 1) return f32 ? launch<false, true, true>(ta, tb, C, p, s) : launch<false, true, false>(ta, tb, C, p, s);
  if (layout == 2) return f32 ? launch<true, true, true>(ta, tb, C, p, s) : launch<true, true, false>(ta, tb, C, p, s);
  return -2;
}
