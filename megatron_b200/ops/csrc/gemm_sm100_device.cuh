// Device-side building blocks shared by the plain (gemm_sm100.cu) and collective-fused (fused_tp_gemm.cu)
// tcgen05 GEMM kernels: tile constants, cluster / 2-SM PTX wrappers, 256-bit global accesses, epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

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


// ---- TMA-store epilogue ------------------------------------------------------------------------------------------------------------
// The direct epilogue above writes one 32-byte sector per thread and row (32 scattered sectors per warp instruction); on short-K problems
// (TP=8 proj: K = 512) the tile's stores then cost as much as its MMAs.  Here a warp converts its 32 rows x 128 bytes (64 bf16 / 32 fp32 columns),
// writes them into its own 128B-swizzled 4 KiB staging tile (conflict-free 16-byte st.shared), and ONE elected lane hands the tile to the TMA unit:
// full 128-byte lines, asynchronous, clipped at the matrix edge by hardware.  `accumulate` (wgrad into fp32/bf16 main_grad) uses the TMA
// reduce-add instead of a read-modify-write through registers.  Two staging tiles per warp alternate; `wait_group.read 1` keeps one store in flight.
constexpr int EPI_TILE_BYTES = 32 * 128;
constexpr int EPI_BYTES_PER_WARP = 2 * EPI_TILE_BYTES;
constexpr int EPI_SMEM_BYTES = EPI_WARPS * EPI_BYTES_PER_WARP;   // 64 KiB

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tmap, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// this warp's 32 TMEM lanes (rows row0 .. row0+31) x columns [c_begin*32, c_end*32) of the tile at column col_base.  `epi` = this warp's two staging tiles
// (1024-byte aligned), `parity` alternates between them across calls.
template <bool C_F32>
__device__ __forceinline__ void epilogue_tile_tma(const CUtensorMap* tmap_c, uint8_t* epi, uint32_t& parity, const GemmParams& p, uint32_t t_base, int row0, int col_base,
                                                  int c_begin, int c_end, int lane, uint64_t* done_bar, bool done_remote) {
  constexpr int CPS = C_F32 ? 1 : 2;      // 32-column TMEM chunks per 128-byte store row
  const int sw = lane & 7;
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += CPS) {
    uint32_t r[CPS][32];
#pragma unroll
    for (int u = 0; u < CPS; ++u) tmem_ld_32x32b_x32(t_base + (c + u) * 32, r[u]);
    tmem_ld_wait();
    if (c + CPS >= c_end) {
      // all of this warp's TMEM reads are done: hand the accumulator back before the stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (done_remote) mbar_arrive_remote(done_bar, 0); else mbar_arrive(done_bar);
      }
    }
    const int col0 = col_base + c * 32;
    if (row0 >= p.M || col0 >= p.N) continue;
    uint8_t* tile = epi + (parity & 1u) * EPI_TILE_BYTES;
    parity ^= 1u;
    // the store that last used this staging tile (two stores ago) must have finished reading it
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
    uint8_t* my_row = tile + lane * 128;
    if (C_F32) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = make_uint4(r[0][4 * j], r[0][4 * j + 1], r[0][4 * j + 2], r[0][4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = 8 * j + 2 * q;         // column within the 64-column row
          __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[e >> 5][e & 31]), __uint_as_float(r[(e + 1) >> 5][(e + 1) & 31]));
          v[q] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (p.accumulate) tma_reduce_add_2d(tmap_c, tile, col0, row0); else tma_store_2d(tmap_c, tile, col0, row0);
      tma_store_commit();
    }
  }
}

// row-major fp32 matrix [rows, cols]; box = {box_cols (<= 32), box_rows}; 128B swizzle (defined in gemm_sm100.cu)
bool make_tmap_f32_strided(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows);
// row-major bf16 matrix [rows, cols]; box = {box_cols (inner, <= 64), box_rows}; 128B swizzle (defined in gemm_sm100.cu)
bool make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);
bool make_tmap_bf16_strided(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows);
bool make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box0, uint32_t box1);
int num_sms();

}  // namespace mb200
