// FP8 (E4M3 / E5M2) GEMM for sm_100a:  C[M,N] (bf16) = alpha * A[M,K] · B[N,K]ᵀ  with per-tensor scales folded into alpha.
// tcgen05.mma kind::f8f6f4 (UMMA_K = 32 one-byte elements): twice the bf16 tensor throughput and half the operand bytes.
// Both operands are K-major (the "NT" form); dgrad / wgrad reach this form through transposed quantised copies, the same
// data flow TransformerEngine uses for its per-tensor-scaled recipes.  Same persistent producer / issuer /
// 8-warp-epilogue pipeline as gemm_sm100.cu (1-CTA, 128 x 256 tile, 128-byte K blocks).
#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int F8_BK = 128;          // one-byte elements per K block = 128 B rows (SWIZZLE_128B)
constexpr int F8_UMMA_K = 32;

// a_fmt / b_fmt: 0 = E4M3, 1 = E5M2 (cute::UMMA::MXF8F6F4Format)
__host__ __device__ constexpr uint32_t make_idesc_f8(uint32_t M, uint32_t N, uint32_t a_fmt, uint32_t b_fmt) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Fp8Params {
  GemmParams g;
  float alpha;
  const float* alpha_dev;   // optional device scalar multiplied into alpha (dequantisation scale computed on the GPU, no host sync)
  uint32_t idesc;
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, __nv_bfloat16* __restrict__ C, const Fp8Params p) {
  constexpr int A_BYTES = BM * F8_BK, B_BYTES = BN * F8_BK;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmParams& g = p.g;
  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (g.K + F8_BK - 1) / F8_BK;

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
      mbar_init(&tmem_empty[i], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, g.group_m, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(smem_a + stage * A_BYTES, &tmap_a, &full_bar[stage], kb * F8_BK, m_blk * BM);
          tma_load_2d(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * F8_BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_BYTES), b_addr = smem_u32(smem_b + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < F8_BK / F8_UMMA_K; ++k)
            umma_f8(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 16, 1024), make_smem_desc_sw128(b_addr + k * 32, 16, 1024), p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
    constexpr int CH = BN / 32 / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float alpha = p.alpha * (p.alpha_dev != nullptr ? __ldg(p.alpha_dev) : 1.f);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, g.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      const int row = m_blk * BM + ew * 32 + lane;
#pragma unroll 1
      for (int c = half * CH; c < (half + 1) * CH; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_base + c * 32, r);
        tmem_ld_wait();
        if (c == (half + 1) * CH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int col0 = n_blk * BN + c * 32;
        if (row >= g.M || col0 >= g.N) continue;
        __nv_bfloat16* crow = C + (size_t)row * g.ldc + col0;
        if (col0 + 32 <= g.N && g.ldc % 16 == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            uint32_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[j + 2 * q]) * alpha, __uint_as_float(r[j + 2 * q + 1]) * alpha);
              v[q] = *reinterpret_cast<uint32_t*>(&h);
            }
            st_global_v8(crow + j, v);
          }
        } else {
          for (int j = 0; j < 32 && col0 + j < g.N; ++j) crow[j] = __float2bfloat16_rn(__uint_as_float(r[j]) * alpha);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace mb200

using namespace mb200;

// one-byte tensor map: rows x cols bytes, box {128 bytes, box_rows}
bool mb200_make_tmap_u8(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);

// A [M,K], B [N,K] one-byte fp8 (a_fmt/b_fmt: 0 = e4m3, 1 = e5m2), C [M,N] bf16 = alpha * A·Bᵀ
extern "C" int mb200_gemm_fp8_nt(const void* A, const void* B, void* C, int M, int N, int K, int a_fmt, int b_fmt, float alpha, const float* alpha_dev, cudaStream_t s) {
  if (K % 16 != 0 || N % 8 != 0) return -11;
  constexpr int BN = 256;
  constexpr int STAGE_BYTES = BM * F8_BK + BN * F8_BK;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ta, tb;
  if (!mb200_make_tmap_u8(&ta, A, M, K, BM) || !mb200_make_tmap_u8(&tb, B, N, K, BN)) return -1;
  auto kern = gemm_fp8_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  Fp8Params p;
  p.g.M = M; p.g.N = N; p.g.K = K; p.g.ldc = N; p.g.accumulate = 0; p.g.group_m = 8;
  p.alpha = alpha;
  p.alpha_dev = alpha_dev;
  p.idesc = make_idesc_f8(BM, BN, (uint32_t)a_fmt, (uint32_t)b_fmt);
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, reinterpret_cast<__nv_bfloat16*>(C), p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
