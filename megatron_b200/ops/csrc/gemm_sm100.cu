// Persistent, warp-specialised bf16 GEMMs for sm_100a: TMA → smem ring → tcgen05.mma → TMEM → epilogue.
//
// Two kernels share the producer / issuer / epilogue structure:
//   gemm_1cta : one CTA per SM, 128 x BN tile (BN = 256 or 128), tcgen05.mma.cta_group::1
//   gemm_2cta : CTA *pairs* (cluster 2x1x1) own a 256 x BN tile; each CTA stages its 128 rows of A and
//               HALF of B, the leader issues tcgen05.mma.cta_group::2 (M = 256).  ncu on the 1-CTA kernel
//               (profiles/r1_gemm_ncu.md) shows the tensor pipe 73 % active with L2→SM at 54 %: per-SM
//               shared-memory bandwidth (96 B/clk UMMA reads + 96 B/clk TMA writes vs 128 B/clk) is the
//               limiter; halving the B traffic per SM removes it.
//
//   warp 0 : TMA producer (one elected lane)          warp 2 : TMEM allocator
//   warp 1 : MMA issuer   (one elected lane)          warps 4-11 : epilogue (tcgen05.ld → cvt → 256-bit st.global)
//
// Layouts (row-major bf16 tensors):
//   0 NT : C[M,N] = A[M,K]  · B[N,K]^T     forward          (A, B K-major)
//   1 NN : C[M,N] = A[M,K]  · B[K,N]       dgrad            (B MN-major)
//   2 TN : C[M,N] = A[K,M]^T · B[K,N]      wgrad            (A, B MN-major); C may be fp32/bf16 with beta = 1
// Replaces cuBLAS-through-torch.matmul and Apex `wgrad_gemm_accum_fp32`.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdio.h>

#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

// =========================================================================================================
// 1-CTA kernel: 128 x BN tile
// =========================================================================================================
template <bool A_MN, bool B_MN, bool C_F32, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_1cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ Cptr, GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator buffers (power of two for BN = 128/256)
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
        tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tmap_a, &full_bar[stage], m_blk * BM + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tmap_b, &full_bar[stage], n_blk * BN + c * 64, kb * BK);
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
    constexpr int CH = BN / 32 / 2;  // 32-column chunks per column half
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<C_F32>(Cptr, p, tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16), m_blk * BM + ew * 32 + lane, n_blk * BN, half * CH, (half + 1) * CH, lane,
                           &tmem_empty[acc], false);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// =========================================================================================================
// 2-CTA kernel: a CTA pair owns a 256 x BN tile; CTA r stages A rows [r*128, +128) and B rows [r*BN/2, +BN/2)
// =========================================================================================================
template <bool A_MN, bool B_MN, bool C_F32, int BN, bool TMA_EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_c, void* __restrict__ Cptr,
                 GemmParams p) {
  constexpr int BNH = BN / 2;                       // B rows staged by each CTA
  constexpr int B_STAGE_BYTES = BNH * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // per CTA
  constexpr int EPI_BYTES = TMA_EPI ? EPI_SMEM_BYTES : 0;     // staging tiles of the TMA-store epilogue come out of the stage budget
  constexpr int STAGES = (SMEM_BUDGET - EPI_BYTES) / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int PM = 2 * BM;                        // pair tile rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem_base;
  uint8_t* smem = smem_base + EPI_BYTES;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_m = (p.M + PM - 1) / PM, tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);   // leader: expect_tx arrive + the peer's remote arrive
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);   // multicast tcgen05.commit
      mbar_init(&tmem_empty[i], 2 * EPI_WARPS);  // epilogue warps of BOTH CTAs (the leader's copy is the one waited on)
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================= TMA producer (both CTAs; completion is signalled on the LEADER's full barrier) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
        const int m0 = m_blk * PM + (int)cta_rank * BM;
        const int n0 = n_blk * BN + (int)cta_rank * BNH;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES); else mbar_arrive_remote(&full_bar[stage], 0);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(sa + c * (BK * 128), &tmap_a, &full_bar[stage], m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BNH / 64; ++c) tma_load_2d_2sm(sb + c * (BK * 128), &tmap_b, &full_bar[stage], n0 + c * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: leader CTA only, M = 256 across the pair ===============================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PM, BN, A_MN, B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
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
            umma_f16_2sm(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);  // frees the slot in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: every CTA drains its own 128 TMEM lanes =====================================
    const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
    constexpr int CH = BN / 32 / 2;
    int acc = 0;
    uint32_t acc_phase = 0, epi_parity = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (TMA_EPI)
        epilogue_tile_tma<C_F32>(&tmap_c, smem_epi + (warp - 4) * EPI_BYTES_PER_WARP, epi_parity, p, tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16),
                                 m_blk * PM + (int)cta_rank * BM + ew * 32, n_blk * BN, half * CH, (half + 1) * CH, lane, &tmem_empty[acc], !leader);
      else
        epilogue_tile<C_F32>(Cptr, p, tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16), m_blk * PM + (int)cta_rank * BM + ew * 32 + lane, n_blk * BN, half * CH,
                             (half + 1) * CH, lane, &tmem_empty[acc], !leader);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMA_EPI && lane == 0) tma_store_wait_read<0>();   // the staging tiles must outlive the stores that read them
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is a DRIVER call: it fails with CUDA_ERROR_INVALID_CONTEXT (201) on a thread that has not yet bound the primary context —
// e.g. PyTorch's autograd thread when the caching allocator served every allocation from its cache and no runtime call has run there yet.
static void ensure_context_on_this_thread() {
  static thread_local bool ready = false;
  if (!ready) {
    cudaFree(nullptr);
    ready = true;
  }
}

static EncodeTiledFn get_encode_fn() {
  ensure_context_on_this_thread();
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
bool make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  return make_tmap_bf16_strided(out, ptr, rows, cols, cols * 2, box_cols, box_rows);
}

// same, with an explicit row pitch in bytes (multiple of 16): views into wider buffers (e.g. q/k/v slices of a fused QKV output)
bool make_tmap_bf16_strided(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("MB200_DEBUG_TMAP"))
    fprintf(stderr, "[mb200] cuTensorMapEncodeTiled failed (%d): ptr %p rows %llu cols %llu pitch %llu box %u x %u\n", (int)r, ptr, (unsigned long long)rows,
            (unsigned long long)cols, (unsigned long long)row_pitch_bytes, box_cols, box_rows);
  return r == CUDA_SUCCESS;
}

// row-major fp32 matrix [rows, cols]; box = {box_cols (inner, <= 32), box_rows}; 128B swizzle (TMA-store / reduce-add epilogue of fp32 outputs)
bool make_tmap_f32_strided(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bf16 [d2][d1][d0] with d0 contiguous; pitches in bytes (multiples of 16); box = {box0 (<= 64), box1, 1}; 128B swizzle.
// Out-of-bounds parts of a box are clipped on stores and zero-filled on loads PER outer index (tiles never bleed into the next matrix).
bool make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace mb200

// one-byte elements (fp8): rows x cols bytes, box {128 bytes, box_rows}, 128B swizzle
bool mb200_make_tmap_u8(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  mb200::EncodeTiledFn enc = mb200::get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace mb200 {

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <bool A_MN, bool B_MN, bool C_F32, int BN, bool TWO_CTA, bool TMA_EPI>
static int launch(const void* A, const void* B, void* C, GemmParams p, cudaStream_t s) {
  static_assert(TWO_CTA || !TMA_EPI, "the TMA-store epilogue is implemented for the 2-CTA kernels");
  constexpr int BNH = TWO_CTA ? BN / 2 : BN;  // B rows per CTA
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BNH * BK * 2;
  constexpr int EPI_BYTES = TMA_EPI ? EPI_SMEM_BYTES : 0;
  constexpr int STAGES = (SMEM_BUDGET - EPI_BYTES) / STAGE_BYTES;
  constexpr int SMEM_BYTES = EPI_BYTES + STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ta, tb, tc;
  bool ok = true;
  ok &= A_MN ? make_tmap_bf16(&ta, A, p.K, p.M, 64, BK) : make_tmap_bf16(&ta, A, p.M, p.K, BK, BM);
  ok &= B_MN ? make_tmap_bf16(&tb, B, p.K, p.N, 64, BK) : make_tmap_bf16(&tb, B, p.N, p.K, BK, BNH);
  if (TMA_EPI)
    ok &= C_F32 ? make_tmap_f32_strided(&tc, C, p.M, p.N, (uint64_t)p.ldc * 4, 32, 32) : make_tmap_bf16_strided(&tc, C, p.M, p.N, (uint64_t)p.ldc * 2, 64, 32);
  else
    tc = ta;
  if (!ok) return -1;
  static bool configured = false;
  if (TWO_CTA) {
    auto kern = gemm_2cta_kernel<A_MN, B_MN, C_F32, BN, TMA_EPI>;
    if (!configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
      configured = true;
    }
    const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN);
    const int clusters = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
    kern<<<clusters * 2, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, tc, C, p);
  } else {
    auto kern = gemm_1cta_kernel<A_MN, B_MN, C_F32, BN>;
    if (!configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
      configured = true;
    }
    const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, C, p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

template <int BN, bool TWO_CTA, bool TMA_EPI = false>
static int dispatch_layout(const void* A, const void* B, void* C, const GemmParams& p, int layout, bool f32, cudaStream_t s) {
  if (layout == 0) return f32 ? launch<false, false, true, BN, TWO_CTA, TMA_EPI>(A, B, C, p, s) : launch<false, false, false, BN, TWO_CTA, TMA_EPI>(A, B, C, p, s);
  if (layout == 1) return f32 ? launch<false, true, true, BN, TWO_CTA, TMA_EPI>(A, B, C, p, s) : launch<false, true, false, BN, TWO_CTA, TMA_EPI>(A, B, C, p, s);
  if (layout == 2) return f32 ? launch<true, true, true, BN, TWO_CTA, TMA_EPI>(A, B, C, p, s) : launch<true, true, false, BN, TWO_CTA, TMA_EPI>(A, B, C, p, s);
  return -2;
}

}  // namespace mb200

using namespace mb200;

// variant: 0 = heuristic, 1 = 1cta/256, 2 = 1cta/128, 3 = 2cta/256, 4 = 2cta/128, 5 / 6 = 3 / 4 with the TMA-store epilogue.  returns 0 on success.
extern "C" int mb200_gemm_bf16_v(const void* A, const void* B, void* C, int M, int N, int K, int layout, int accumulate, int c_dtype, int variant,
                                 cudaStream_t s) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = N; p.accumulate = accumulate; p.group_m = 8;
  const bool f32 = c_dtype == kF32;
  if (variant == 0) {
    const long t2 = (long)((M + 255) / 256) * ((N + 255) / 256);
    const long t1 = (long)((M + 127) / 128) * ((N + 255) / 256);
    if (t2 >= 2 * 74) variant = 3;           // >= 2 waves of 256x256 pair tiles
    else if (t1 >= 148) variant = 1;
    else variant = 2;
  }
  switch (variant) {
    case 1: return dispatch_layout<256, false>(A, B, C, p, layout, f32, s);
    case 2: return dispatch_layout<128, false>(A, B, C, p, layout, f32, s);
    case 3: return dispatch_layout<256, true>(A, B, C, p, layout, f32, s);
    case 4: return dispatch_layout<128, true>(A, B, C, p, layout, f32, s);
    case 5: return dispatch_layout<256, true, true>(A, B, C, p, layout, f32, s);   // 2-CTA 256x256 + TMA-store / reduce-add epilogue
    case 6: return dispatch_layout<128, true, true>(A, B, C, p, layout, f32, s);   // 2-CTA 256x128 + TMA-store / reduce-add epilogue
    default: return -5;
  }
}
extern "C" int mb200_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int layout, int accumulate, int c_dtype, cudaStream_t s) {
  return mb200_gemm_bf16_v(A, B, C, M, N, K, layout, accumulate, c_dtype, 0, s);
}
