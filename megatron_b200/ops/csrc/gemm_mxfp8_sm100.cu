// MXFP8 block-scaled GEMM for sm_100a:  C[M,N] (bf16) = Σ_kb (A_q · 2^(sfa-127)) · (B_q · 2^(sfb-127))ᵀ
// A_q [M,K], B_q [N,K]: E4M3 bytes, K-major.  One E8M0 scale per 32 consecutive K elements of a row (OCP microscaling).
// tcgen05.mma kind::mxf8f6f4.block_scale applies the scales INSIDE the tensor core: per MMA (K = 32) one scale byte per A row and per B row is read
// from tensor memory, so the accumulator needs no per-block rescale pass (which is what makes 1x32 scaling cost nothing over per-tensor fp8).
//
// Scale-factor staging (the layout the hardware expects, cute Sm1xxBlockScaledBasicChunk): for a 128-row x 128-K block the 128 x 4 scale bytes form a
// 512-byte atom, byte offset = (row % 32) * 16 + (row / 32) * 4 + k.  Atoms are stored [row_block][k_block] in global memory (see mxfp8_swizzle_scales on the
// Python side), fetched with ONE 1-D bulk copy per operand and stage, and moved smem → TMEM with tcgen05.cp.32x128b.warpx4 (32 rows x 16 B, broadcast to the four
// lane quarters): TMEM column c of lane l then holds the four k-scales of row (l % 32) + 32 c, and the MMA of k-step k selects byte k through the
// a_sf_id / b_sf_id fields of the instruction descriptor.  tcgen05.cp and tcgen05.mma execute in issue order, so one SF buffer in TMEM suffices.
//
// Pipeline: same persistent producer / issuer / 8-warp epilogue structure as gemm_fp8_sm100.cu.  Two tile shapes: 128 x 256 with ONE accumulator (256 + 12 of the 512
// TMEM columns; the epilogue is not overlapped, but operand traffic per flop is 2/3 of the small tile's — measured 1.66 PF at 128 x 128 was L2-bound) and 128 x 128 with
// two accumulators for narrow N.
#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int MX_BK = 128;          // K elements (bytes) per block = four scale groups
constexpr int MX_SF_BYTES = 512;    // one scale atom: 128 rows x 4 k-groups

// cute::UMMA::InstrDescriptorBlockScaled: a/b format E4M3 (0), K-major, scale format E8M0 (bit 23), no c_format field (fp32 accumulate implied)
__host__ __device__ constexpr uint32_t make_idesc_mxf8(uint32_t M, uint32_t N, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (a_sf_id << 29);
}
__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
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

struct MxParams {
  GemmParams g;
  const uint8_t* sfa;   // [ceil(M/128)][k_blocks][512]
  const uint8_t* sfb;   // [ceil(N/128)][k_blocks][512]
  int n_atoms;          // ceil(N/128)
};

template <int BN, int ACC_BUFS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, __nv_bfloat16* __restrict__ C, const MxParams p) {
  constexpr int NB_ATOMS = BN / 128;                  // scale atoms of the B tile
  constexpr int SF_STAGE = (1 + NB_ATOMS) * MX_SF_BYTES;
  constexpr int A_BYTES = BM * MX_BK, B_BYTES = BN * MX_BK;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES + SF_STAGE;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 512;                 // accumulators + 4 (SFA) + 4 per 128 B rows (SFB), rounded up to a power of two
  constexpr uint32_t SFA_COL = ACC_BUFS * BN, SFB_COL = ACC_BUFS * BN + 4;
  static_assert(ACC_BUFS * BN + 4 + 4 * NB_ATOMS <= 512, "tensor memory budget");
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
          bulk_load_1d(smem_sf + stage * SF_STAGE, p.sfa + ((size_t)m_blk * k_blocks + kb) * MX_SF_BYTES, MX_SF_BYTES, &full_bar[stage]);
#pragma unroll
          for (int a = 0; a < NB_ATOMS; ++a) {
            // rows beyond N have no atom: clamp to the last one (their products land in columns that are never stored)
            int nb = n_blk * NB_ATOMS + a;
            nb = nb < p.n_atoms ? nb : p.n_atoms - 1;
            bulk_load_1d(smem_sf + stage * SF_STAGE + (1 + a) * MX_SF_BYTES, p.sfb + ((size_t)nb * k_blocks + kb) * MX_SF_BYTES, MX_SF_BYTES, &full_bar[stage]);
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
          utccp_32x128b_warpx4(tmem_base + SFA_COL, make_smem_desc_sf(sf_addr));
#pragma unroll
          for (int a = 0; a < NB_ATOMS; ++a) utccp_32x128b_warpx4(tmem_base + SFB_COL + 4 * a, make_smem_desc_sf(sf_addr + (1 + a) * MX_SF_BYTES));
#pragma unroll
          for (int k = 0; k < MX_BK / 32; ++k)
            umma_mxf8(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 16, 1024), make_smem_desc_sw128(b_addr + k * 32, 16, 1024), make_idesc_mxf8(BM, BN, k, k),
                      tmem_base + SFA_COL, tmem_base + SFB_COL, (kb > 0 || k > 0) ? 1u : 0u);
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
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[j + 2 * q]), __uint_as_float(r[j + 2 * q + 1]));
              v[q] = *reinterpret_cast<uint32_t*>(&h);
            }
            st_global_v8(crow + j, v);
          }
        } else {
          for (int j = 0; j < 32 && col0 + j < g.N; ++j) crow[j] = __float2bfloat16_rn(__uint_as_float(r[j]));
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
static int launch_mx(const void* A, const void* B, const void* sfa, const void* sfb, void* C, int M, int N, int K, cudaStream_t s) {
  constexpr int STAGE_BYTES = BM * MX_BK + BN * MX_BK + (1 + BN / 128) * MX_SF_BYTES;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ta, tb;
  if (!mb200_make_tmap_u8(&ta, A, M, K, BM) || !mb200_make_tmap_u8(&tb, B, N, K, BN)) return -1;
  auto kern = gemm_mxfp8_kernel<BN, ACC_BUFS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  MxParams p;
  p.g.M = M; p.g.N = N; p.g.K = K; p.g.ldc = N; p.g.accumulate = 0; p.g.group_m = 8;
  p.sfa = reinterpret_cast<const uint8_t*>(sfa);
  p.sfb = reinterpret_cast<const uint8_t*>(sfb);
  p.n_atoms = (N + 127) / 128;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, reinterpret_cast<__nv_bfloat16*>(C), p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// A [M,K], B [N,K] e4m3 bytes; sfa / sfb: swizzled scale atoms [ceil(rows/128)][K/128][512]; C [M,N] bf16.  K % 128 == 0.
// tile: 0 = auto (128 x 256 when N > 128), 128 or 256 forces the tile width.
extern "C" int mb200_gemm_mxfp8_nt(const void* A, const void* B, const void* sfa, const void* sfb, void* C, int M, int N, int K, int tile, cudaStream_t s) {
  if (K % MX_BK != 0 || N % 8 != 0) return -11;
  if (tile == 0) tile = N > 128 ? 256 : 128;
  if (tile == 256) return launch_mx<256, 1>(A, B, sfa, sfb, C, M, N, K, s);
  return launch_mx<128, 2>(A, B, sfa, sfb, C, M, N, K, s);
}
