// Grouped (per-expert) bf16 GEMMs for MoE on sm_100a: ONE persistent tcgen05 kernel walks every (expert, tile) pair.
//
//   mode 0  fwd    out[g_e] = x[g_e] · w[e]ᵀ          x [T, K]   w [E, N, K]   out [T, N]       (rows grouped by expert)
//   mode 1  dgrad  gx[g_e]  = gy[g_e] · w[e]          gy [T, N]  w [E, N, K]   gx [T, K]
//   mode 2  wgrad  gw[e]   (+)= gy[g_e]ᵀ · x[g_e]     gy [T, N]  x [T, K]      gw [E, N, K]  (fp32 or bf16, optional accumulate)
//
// Expert segments have arbitrary (unaligned) lengths.  Each expert gets its OWN TMA descriptor over exactly its rows, so
// partial tiles / partial reduction blocks are zero-filled by the TMA unit and never see a neighbouring expert's tokens;
// the descriptors are encoded on the host per call and read by the kernel from global memory.  Output rows past the end of a
// segment are masked by the epilogue.  Same producer / issuer / 8-warp-epilogue pipeline as ``gemm_sm100.cu`` (1-CTA tiles).
// Replaces TE GroupedLinear / cuBLAS grouped GEMM and the per-expert loop of SequentialMLP.
#include <vector>

#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int MAX_EXPERTS = 128;

struct GroupedParams {
  int E;
  int rows_out;                    // output rows per tile-owner: modes 0/1 → unused (per expert), mode 2 → N_w
  int n_out;                       // output columns (N of the per-expert GEMM)
  int k_red;                       // reduction length for modes 0/1 (fixed); mode 2 uses the per-expert token count
  int b_rows_per_expert;           // rows of one expert inside the shared B descriptor (modes 0/1)
  int ldc;
  int accumulate;
  long c_expert_stride;            // mode 2: elements between gw[e] and gw[e+1]
  int tile_prefix[MAX_EXPERTS + 1];
  int offsets[MAX_EXPERTS + 1];
};

__device__ __forceinline__ int find_expert(const GroupedParams& p, int tile) {
  int e = 0;
  while (e + 1 < p.E && p.tile_prefix[e + 1] <= tile) ++e;
  return e;
}

// MODE 0: A K-major, B K-major.  MODE 1: A K-major, B MN-major.  MODE 2: A, B MN-major (reduction over tokens).
template <int MODE, bool C_F32, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
grouped_gemm_kernel(const CUtensorMap* __restrict__ maps, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ Cptr, const __grid_constant__ GroupedParams p) {
  constexpr bool A_MN = MODE == 2, B_MN = MODE != 0;
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tile_prefix[p.E];
  const int tiles_n = (p.n_out + BN - 1) / BN;

  if (warp == 0 && lane == 0 && MODE != 2) prefetch_tmap(&tmap_b);
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

  // tile → (expert, m_blk, n_blk, k_blocks); m fastest inside an expert so concurrently running CTAs share the weight tile
  auto decode = [&](int tile, int& e, int& m_blk, int& n_blk, int& k_blocks) {
    e = find_expert(p, tile);
    const int local = tile - p.tile_prefix[e];
    const int tokens = p.offsets[e + 1] - p.offsets[e];
    if (MODE == 2) {
      const int tiles_m = (p.rows_out + BM - 1) / BM;
      m_blk = local % tiles_m;
      n_blk = local / tiles_m;
      k_blocks = (tokens + BK - 1) / BK;
    } else {
      const int tiles_m = (tokens + BM - 1) / BM;
      m_blk = local % tiles_m;
      n_blk = local / tiles_m;
      k_blocks = (p.k_red + BK - 1) / BK;
    }
    (void)tiles_n;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int e, m_blk, n_blk, k_blocks;
        decode(tile, e, m_blk, n_blk, k_blocks);
        const CUtensorMap* ma = &maps[e];
        const CUtensorMap* mb = MODE == 2 ? &maps[p.E + e] : &tmap_b;
        const int b_row0 = MODE == 2 ? 0 : e * p.b_rows_per_expert;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, ma, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), ma, &full_bar[stage], m_blk * BM + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, mb, &full_bar[stage], kb * BK, b_row0 + n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), mb, &full_bar[stage], n_blk * BN + c * 64, b_row0 + kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int e, m_blk, n_blk, k_blocks;
        decode(tile, e, m_blk, n_blk, k_blocks);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, BK * 128, 1024) : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, BK * 128, 1024) : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int e, m_blk, n_blk, k_blocks;
      decode(tile, e, m_blk, n_blk, k_blocks);
      GemmParams g;
      g.N = p.n_out; g.K = 0; g.ldc = p.ldc; g.accumulate = p.accumulate; g.group_m = 1;
      void* cbase = Cptr;
      int row;
      if (MODE == 2) {
        g.M = p.rows_out;
        row = m_blk * BM + ew * 32 + lane;
        cbase = C_F32 ? static_cast<void*>(reinterpret_cast<float*>(Cptr) + (size_t)e * p.c_expert_stride)
                      : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(Cptr) + (size_t)e * p.c_expert_stride);
      } else {
        g.M = p.offsets[e + 1];                                   // absolute row limit of this expert's segment
        row = p.offsets[e] + m_blk * BM + ew * 32 + lane;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<C_F32>(cbase, g, tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16), row, n_blk * BN, half * CH, (half + 1) * CH, lane, &tmem_empty[acc], false);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int MODE, bool C_F32, int BN>
static int launch_grouped(const CUtensorMap* maps_dev, const CUtensorMap& tb, void* C, const GroupedParams& p, cudaStream_t s) {
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BN * BK * 2;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  auto kern = grouped_gemm_kernel<MODE, C_F32, BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  const int tiles = p.tile_prefix[p.E];
  if (tiles == 0) return 0;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(maps_dev, tb, C, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace mb200

using namespace mb200;

// offsets: host int32 [E+1] (token prefix).  maps_dev: device scratch of at least 2*E*sizeof(CUtensorMap) bytes (64-byte aligned).
// a / b / c as in the header comment for each mode; c_dtype: 0 fp32, 1 bf16.
extern "C" int mb200_grouped_gemm_bf16(const void* a, const void* b, void* c, const int* offsets, int E, int dim_n, int dim_k, int mode, int accumulate, int c_dtype,
                                       void* maps_dev, cudaStream_t s) {
  if (E < 1 || E > MAX_EXPERTS) return -10;
  if (dim_n % 8 != 0 || dim_k % 8 != 0) return -11;
  GroupedParams p;
  p.E = E; p.accumulate = accumulate; p.c_expert_stride = (long)dim_n * dim_k;
  for (int e = 0; e <= E; ++e) p.offsets[e] = offsets[e];
  const int T = offsets[E];
  std::vector<CUtensorMap> maps(2 * E);
  CUtensorMap tb;
  bool ok = true;
  const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(a);
  const __nv_bfloat16* B = reinterpret_cast<const __nv_bfloat16*>(b);
  int out_cols;
  if (mode == 0) {            // x [T, K] · w[e] [N, K]ᵀ
    p.n_out = dim_n; p.k_red = dim_k; p.b_rows_per_expert = dim_n; p.ldc = dim_n; p.rows_out = 0;
    out_cols = dim_n;
  } else if (mode == 1) {     // gy [T, N] · w[e] [N, K]
    p.n_out = dim_k; p.k_red = dim_n; p.b_rows_per_expert = dim_n; p.ldc = dim_k; p.rows_out = 0;
    out_cols = dim_k;
  } else {                    // gy[g_e]ᵀ [N, M_e] · x[g_e] [M_e, K]
    p.n_out = dim_k; p.k_red = 0; p.b_rows_per_expert = 0; p.ldc = dim_k; p.rows_out = dim_n;
    out_cols = dim_k;
  }
  const bool small_n = out_cols <= 128;
  const int BNsel = small_n ? 128 : 256;
  const int tiles_n = (out_cols + BNsel - 1) / BNsel;
  p.tile_prefix[0] = 0;
  for (int e = 0; e < E; ++e) {
    const int tokens = offsets[e + 1] - offsets[e];
    int tiles;
    if (mode == 2) tiles = tokens > 0 ? ((dim_n + BM - 1) / BM) * tiles_n : 0;
    else tiles = ((tokens + BM - 1) / BM) * tiles_n;
    p.tile_prefix[e + 1] = p.tile_prefix[e] + tiles;
    const uint64_t rows = tokens > 0 ? (uint64_t)tokens : 1;   // a descriptor needs a non-empty extent; empty experts own no tiles
    const size_t row0 = (size_t)offsets[e];
    if (mode == 0) ok &= make_tmap_bf16(&maps[e], A + row0 * dim_k, rows, dim_k, BK, BM);
    else if (mode == 1) ok &= make_tmap_bf16(&maps[e], A + row0 * dim_n, rows, dim_n, BK, BM);
    else {
      ok &= make_tmap_bf16(&maps[e], A + row0 * dim_n, rows, dim_n, 64, BK);        // gy segment, MN-major A
      ok &= make_tmap_bf16(&maps[E + e], B + row0 * dim_k, rows, dim_k, 64, BK);    // x segment, MN-major B
    }
  }
  (void)T;
  if (mode == 0) ok &= make_tmap_bf16(&tb, b, (uint64_t)E * dim_n, dim_k, BK, BNsel);
  else if (mode == 1) ok &= make_tmap_bf16(&tb, b, (uint64_t)E * dim_n, dim_k, 64, BK);
  else tb = maps[0];
  if (!ok) return -1;
  if (cudaMemcpyAsync(maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s) != cudaSuccess) return -2;
  const CUtensorMap* md = reinterpret_cast<const CUtensorMap*>(maps_dev);
  const bool f32 = c_dtype == 0;
  if (mode == 0) return small_n ? launch_grouped<0, false, 128>(md, tb, c, p, s) : launch_grouped<0, false, 256>(md, tb, c, p, s);
  if (mode == 1) return small_n ? launch_grouped<1, false, 128>(md, tb, c, p, s) : launch_grouped<1, false, 256>(md, tb, c, p, s);
  if (f32) return small_n ? launch_grouped<2, true, 128>(md, tb, c, p, s) : launch_grouped<2, true, 256>(md, tb, c, p, s);
  return small_n ? launch_grouped<2, false, 128>(md, tb, c, p, s) : launch_grouped<2, false, 256>(md, tb, c, p, s);
}
