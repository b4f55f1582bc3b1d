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
#include <mutex>
#include <stdio.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace mb200 {
using namespace ptx;

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 384;  // warps 0-3: producer / issuer / allocator / spare; warps 4-11: epilogue
constexpr int EPI_WARPS = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB per CTA
constexpr int SMEM_BUDGET = 200 * 1024;

struct GemmParams {
  int M, N, K;
  int ldc;
  int accumulate;  // C += A*B
  int group_m;     // rasterisation group (in tile rows)
};

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int& m_blk, int& n_blk) {
  const int per_group = group_m * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int gsz = min(group_m, tiles_m - first_m);
  const int r = tile - g * per_group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

// ---- cluster helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(local_bar)), "r"(cta)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_holder) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// ---- 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256) -----------------------------------------------
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

// ---- shared epilogue: this warp's 32 TMEM lanes (rows) x columns [c_begin*32, c_end*32) -----------------------
// 8 epilogue warps: warp%4 selects the TMEM lane quarter (hardware restriction), warp/4 the column half, so two
// warps drain each quarter concurrently.  Every thread owns one output row and writes whole 32-byte sectors.
template <bool C_F32>
__device__ __forceinline__ void epilogue_tile(void* __restrict__ Cptr, const GemmParams& p, uint32_t t_base, int row, int col_base, int c_begin, int c_end, int lane,
                                              uint64_t* done_bar, bool done_remote) {
  const bool row_ok = row < p.M;
  const bool vec32 = (p.ldc % (C_F32 ? 8 : 16)) == 0;  // 32-byte aligned rows → 256-bit stores
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_base + c * 32, r);
    tmem_ld_wait();
    if (c == c_end - 1) {
      // all of this warp's TMEM reads are done → hand the accumulator back before the global stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (done_remote) mbar_arrive_remote(done_bar, 0); else mbar_arrive(done_bar);
      }
    }
    const int col0 = col_base + c * 32;
    if (!row_ok || col0 >= p.N) continue;
    if (C_F32) {
      float* crow = reinterpret_cast<float*>(Cptr) + (size_t)row * p.ldc + col0;
      if (col0 + 32 <= p.N && vec32) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t v[8];
          if (p.accumulate) {
            ld_global_v8(crow + j, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(r[j + q]));
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = r[j + q];
          }
          st_global_v8(crow + j, v);
        }
      } else {
        for (int j = 0; j < 32 && col0 + j < p.N; ++j) crow[j] = __uint_as_float(r[j]) + (p.accumulate ? crow[j] : 0.f);
      }
    } else {
      __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(Cptr) + (size_t)row * p.ldc + col0;
      if (col0 + 32 <= p.N && vec32) {
#pragma unroll
        for (int j = 0; j < 32; j += 16) {
          float f[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(r[j + q]);
          if (p.accumulate) {
            uint32_t o[8];
            ld_global_v8(crow + j, o);
            const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(o);
#pragma unroll
            for (int q = 0; q < 16; ++q) f[q] += __bfloat162float(ob[q]);
          }
          uint32_t v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
            v[q] = *reinterpret_cast<uint32_t*>(&h);
          }
          st_global_v8(crow + j, v);
        }
      } else if (col0 + 32 <= p.N) {
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
}

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
template <bool A_MN, bool B_MN, bool C_F32, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ Cptr, GemmParams p) {
  constexpr int BNH = BN / 2;                       // B rows staged by each CTA
  constexpr int B_STAGE_BYTES = BNH * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // per CTA
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int PM = 2 * BM;                        // pair tile rows
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
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, p.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<C_F32>(Cptr, p, tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16), m_blk * PM + (int)cta_rank * BM + ew * 32 + lane, n_blk * BN, half * CH,
                           (half + 1) * CH, lane, &tmem_empty[acc], !leader);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

// ---- host side ---------------------------------------------------------------------------------------------
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
bool make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
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

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <bool A_MN, bool B_MN, bool C_F32, int BN, bool TWO_CTA>
static int launch(const void* A, const void* B, void* C, GemmParams p, cudaStream_t s) {
  constexpr int BNH = TWO_CTA ? BN / 2 : BN;  // B rows per CTA
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BNH * BK * 2;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ta, tb;
  bool ok = true;
  ok &= A_MN ? make_tmap_bf16(&ta, A, p.K, p.M, 64, BK) : make_tmap_bf16(&ta, A, p.M, p.K, BK, BM);
  ok &= B_MN ? make_tmap_bf16(&tb, B, p.K, p.N, 64, BK) : make_tmap_bf16(&tb, B, p.N, p.K, BK, BNH);
  if (!ok) return -1;
  static bool configured = false;
  if (TWO_CTA) {
    auto kern = gemm_2cta_kernel<A_MN, B_MN, C_F32, BN>;
    if (!configured) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
      configured = true;
    }
    const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN);
    const int clusters = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
    kern<<<clusters * 2, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, C, p);
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

template <int BN, bool TWO_CTA>
static int dispatch_layout(const void* A, const void* B, void* C, const GemmParams& p, int layout, bool f32, cudaStream_t s) {
  if (layout == 0) return f32 ? launch<false, false, true, BN, TWO_CTA>(A, B, C, p, s) : launch<false, false, false, BN, TWO_CTA>(A, B, C, p, s);
  if (layout == 1) return f32 ? launch<false, true, true, BN, TWO_CTA>(A, B, C, p, s) : launch<false, true, false, BN, TWO_CTA>(A, B, C, p, s);
  if (layout == 2) return f32 ? launch<true, true, true, BN, TWO_CTA>(A, B, C, p, s) : launch<true, true, false, BN, TWO_CTA>(A, B, C, p, s);
  return -2;
}

}  // namespace mb200

using namespace mb200;

// variant: 0 = heuristic, 1 = 1cta/256, 2 = 1cta/128, 3 = 2cta/256, 4 = 2cta/128.  returns 0 on success.
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
    default: return -5;
  }
}
extern "C" int mb200_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int layout, int accumulate, int c_dtype, cudaStream_t s) {
  return mb200_gemm_bf16_v(A, B, C, M, N, K, layout, accumulate, c_dtype, 0, s);
}
