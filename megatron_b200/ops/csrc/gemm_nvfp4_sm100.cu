// NVFP4 block-scaled GEMM for sm_100a:  C[M,N] (bf16) = alpha · Σ (A_q · sfa) · (B_q · sfb)ᵀ
// A_q [M,K], B_q [N,K]: E2M1 (4-bit) codes packed two per byte (low nibble = even k), K-major.  One UE4M3 scale per 16 consecutive K elements of a row
// plus one fp32 tensor scale per operand folded into alpha (NVFP4).  tcgen05.mma kind::mxf4nvf4.block_scale.block16: K = 64 elements per MMA
// (32 bytes of packed operand), FOUR scale bytes per row per MMA — i.e. one whole 32-bit TMEM column per row, so every MMA has its own scale atom:
//   atom (512 B) = 128 rows x 4 scales of one K=64 step, byte (r % 32) * 16 + (r / 32) * 4 + j;  atoms stored [row_block][K / 64] in global memory.
// Per 256-element k-block (128 bytes per row, SWIZZLE_128B): A brings 4 atoms, B (256 rows) 2 x 4; they are bulk-copied next to the operand tiles and moved to TMEM
// with tcgen05.cp before the four MMAs of the block.  TMEM: 256 accumulator columns + 16 (SFA) + 32 (SFB).  Same pipeline as gemm_mxfp8_sm100.cu (128 x 256, one accumulator).
#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int MX_BK = 128;          // BYTES of packed operand per row and k-block = 256 fp4 elements = four K=64 MMAs
constexpr int MX_SF_BYTES = 512;    // one scale atom: 128 rows x 4 k-groups

// cute::UMMA::InstrDescriptorBlockScaled for kind::mxf4nvf4: a/b format E2M1 (MXF4Format = 1), K-major, scale format UE4M3 (bit 23 = 0), sf ids 0, K = 64 (bit 31 = 0)
__host__ __device__ constexpr uint32_t make_idesc_nvf4(uint32_t M, uint32_t N) {
  return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_nvf4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.block16 [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// smem descriptor of a scale atom: no swizzle, 32 rows of 16 bytes, 8-row core matrices 128 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc_sf(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;   // SBO
  d |= (uint64_t)1 << 46;                        // version 1, layout_type 0 = SWIZZLE_NONE
  return d;
}
__device__ __forceinline__ void utccp_32x128b_warpx4(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

struct Nvf4Params {
  GemmParams g;
  const uint8_t* sfa;   // [ceil(M/128)][K/64][512]
  const uint8_t* sfb;   // [ceil(N/128)][K/64][512]
  int n_atoms;          // ceil(N/128)
  float alpha;          // product of the two tensor scales
  const float* alpha_dev;
};

template <int BN, int ACC_BUFS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_nvfp4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, __nv_bfloat16* __restrict__ C, const Nvf4Params p) {
  constexpr int NB_ATOMS = BN / 128;                  // 128-row blocks of the B tile
  constexpr int KSTEPS = 4;                           // K=64 MMAs per k-block, one scale atom each
  constexpr int SF_STAGE = (1 + NB_ATOMS) * KSTEPS * MX_SF_BYTES;
  constexpr int A_BYTES = BM * MX_BK, B_BYTES = BN * MX_BK;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES + SF_STAGE;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 512;                 // accumulators + 4 (SFA) + 4 per 128 B rows (SFB), rounded up to a power of two
  constexpr uint32_t SFA_COL = ACC_BUFS * BN, SFB_COL = ACC_BUFS * BN + 4 * KSTEPS;
  static_assert(ACC_BUFS * BN + 4 * KSTEPS * (1 + NB_ATOMS) <= 512, "tensor memory budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_sf = smem + STAGES * (A_BYTES + B_BYTES);          // per stage: SFA atom | SFB atom
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
  const int k_blocks = (g.K + MX_BK - 1) / MX_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_BUFS; ++i) {
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
          tma_load_2d(smem_a + stage * A_BYTES, &tmap_a, &full_bar[stage], kb * MX_BK, m_blk * BM);
          tma_load_2d(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * MX_BK, n_blk * BN);
          bulk_load_1d(smem_sf + stage * SF_STAGE, p.sfa + ((size_t)m_blk * k_blocks + kb) * KSTEPS * MX_SF_BYTES, KSTEPS * MX_SF_BYTES, &full_bar[stage]);
#pragma unroll
          for (int a = 0; a < NB_ATOMS; ++a) {
            // rows beyond N have no atom: clamp to the last one (their products land in columns that are never stored)
            int nb = n_blk * NB_ATOMS + a;
            nb = nb < p.n_atoms ? nb : p.n_atoms - 1;
            bulk_load_1d(smem_sf + stage * SF_STAGE + (1 + a) * KSTEPS * MX_SF_BYTES, p.sfb + ((size_t)nb * k_blocks + kb) * KSTEPS * MX_SF_BYTES, KSTEPS * MX_SF_BYTES,
                         &full_bar[stage]);
          }
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
          const uint32_t sf_addr = smem_u32(smem_sf + stage * SF_STAGE);
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            utccp_32x128b_warpx4(tmem_base + SFA_COL + 4 * k, make_smem_desc_sf(sf_addr + k * MX_SF_BYTES));
#pragma unroll
            for (int a = 0; a < NB_ATOMS; ++a)    // the MMA of step k reads its B scales from 4 * NB_ATOMS consecutive columns: row block a at + 4 a
              utccp_32x128b_warpx4(tmem_base + SFB_COL + 4 * NB_ATOMS * k + 4 * a, make_smem_desc_sf(sf_addr + ((1 + a) * KSTEPS + k) * MX_SF_BYTES));
          }
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
            umma_nvf4(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 16, 1024), make_smem_desc_sw128(b_addr + k * 32, 16, 1024), make_idesc_nvf4(BM, BN),
                      tmem_base + SFA_COL + 4 * k, tmem_base + SFB_COL + 4 * NB_ATOMS * k, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
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
      if (++acc == ACC_BUFS) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace mb200

using namespace mb200;

bool mb200_make_tmap_u8(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);

template <int BN, int ACC_BUFS>
static int launch_nvf4(const void* A, const void* B, const void* sfa, const void* sfb, void* C, int M, int N, int K, float alpha, const float* alpha_dev, cudaStream_t s) {
  constexpr int STAGE_BYTES = BM * MX_BK + BN * MX_BK + (1 + BN / 128) * 4 * MX_SF_BYTES;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ta, tb;
  // packed fp4: K / 2 bytes per row; a byte-wise SWIZZLE_128B copy produces exactly the dense 4-bit K-major layout the MMA reads
  if (!mb200_make_tmap_u8(&ta, A, M, K / 2, BM) || !mb200_make_tmap_u8(&tb, B, N, K / 2, BN)) return -1;
  auto kern = gemm_nvfp4_kernel<BN, ACC_BUFS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  Nvf4Params p;
  p.g.M = M; p.g.N = N; p.g.K = K / 2; p.g.ldc = N; p.g.accumulate = 0; p.g.group_m = 8;      // K in BYTES: the kernel walks 128-byte k-blocks
  p.sfa = reinterpret_cast<const uint8_t*>(sfa);
  p.sfb = reinterpret_cast<const uint8_t*>(sfb);
  p.n_atoms = (N + 127) / 128;
  p.alpha = alpha;
  p.alpha_dev = alpha_dev;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, reinterpret_cast<__nv_bfloat16*>(C), p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// A [M,K/2], B [N,K/2] packed e2m1 bytes; sfa / sfb: UE4M3 scale atoms [ceil(rows/128)][K/64][512]; C [M,N] bf16 = alpha * A·Bᵀ.  K % 256 == 0.
extern "C" int mb200_gemm_nvfp4_nt(const void* A, const void* B, const void* sfa, const void* sfb, void* C, int M, int N, int K, float alpha, const float* alpha_dev,
                                   cudaStream_t s) {
  if (K % 256 != 0 || N % 8 != 0) return -11;
  return launch_nvf4<256, 1>(A, B, sfa, sfb, C, M, N, K, alpha, alpha_dev, s);
}
