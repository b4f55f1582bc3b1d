// Flash attention backward for sm_100a (tcgen05 / TMEM / TMA) — two tensor-core kernels and no atomics.
//
// Round-1 version: one key-block-stationary kernel that also produced dQ partials and added them to an fp32 buffer with
// red.global.add.  Measured 415 TFLOP/s: 7,100 cycles per (128 key x 64 query) step against a tensor floor of 1,280 — the
// 4.3 GB of L2 reductions (32 KB per step) ran at ~1.3 TB/s and bound the kernel.  This version removes that traffic:
//
//   fa_delta_kernel      delta[b,h,q] = sum_d dO*O                                    (bandwidth-bound prologue)
//   fa_bwd_dkv_kernel    one CTA = a 128-key block of one KV head (or of one query head in split-heads mode), walking the
//                        64-query blocks of its GQA group.  TRANSPOSED orientation (key = MMA M = TMEM lane):
//                          ST  = K_j Q_i^T          UMMA 128x64x16 (K-major A, B)                        -> TMEM [kv, q]
//                          dPT = V_j dO_i^T         same                                                  -> TMEM [kv, q]
//                          PT  = exp2(ST*c - lse_q) ;  dST = PT o (dPT - D_q) * scale     (thread = one key row)
//                          dV += PT  dO_i           UMMA 128x128x16, A = PT  FROM TMEM, B = dO_i (MN-major view of the same smem tile)
//                          dK += dST Q_i            UMMA 128x128x16, A = dST FROM TMEM, B = Q_i
//                        dK / dV stay in tensor memory for the whole walk; the bf16 dST tile is ALSO written to a swizzled smem tile and
//                        leaves the SM with ONE TMA store per step into a global scratch dSt[b*hq][sk][sq] (2.1 GB for the causal Llama-3
//                        8B shape, streamed at ~2 TB/s under the MMAs).
//   fa_bwd_dq_kernel     persistent tcgen05 GEMM over (head, 128-query block) tiles, heaviest first:  dQ = sum_kb dS[q-block, kb] K[kb]
//                        with A = the dSt tile read back by TMA as an MN-major operand, B = K_kb (MN-major), accumulators double-buffered
//                        in TMEM, bf16 dQ written straight from the epilogue (no fp32 dQ buffer, no zero-fill, no conversion pass).
//
// Tensor work is unchanged (5 MMAs per pair, no recomputation); the price is 2 x 2.1 GB of HBM traffic that overlaps with MMAs.
// Split-heads mode (grid.y = query heads; fp32 dK/dV partials + fa_bwd_reduce_kernel) keeps all 148 SMs busy when a TP rank owns
// only a few heads (TP=8: 4 query heads / 1 KV head -> 64 CTAs otherwise).
// Warps 0-7: compute (thread = TMEM lane x column half), warp 8: TMA producer, warp 9: MMA issuer + TMEM allocator.
#include <cstdlib>
#include <type_traits>

#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int FB_KV = 128;      // keys per CTA
constexpr int FB_Q = 64;        // queries per step
constexpr int FB_THREADS = 320;   // 8 compute warps (two per TMEM lane quarter: each takes 32 of the 64 query columns) + producer + issuer
constexpr int FB_D = 128;       // head dim (this version)
constexpr int FB_STAGES = 3;    // Q_i / dO_i ring: the load of step st+2 must be in flight while step st's dV/dK MMAs still hold their stage (2 stages put a TMA round trip on every step)

struct FaBwdParams {
  int sq, sk, b, hq, hk;
  int causal;
  int split_heads;       // 1: blockIdx.y = query head, dK/dV leave as fp32 partials [b, hq, sk, d]
  int dbg;               // bottleneck experiments (MB200_FA_BWD_DBG): 1 skip the score math, 2 skip the dS TMA store, 4 skip dV/dK MMAs, 8 skip score MMAs — results are WRONG
  float scale, scale_log2;
  long q_sb, q_sh, k_sb, k_sh, v_sb, v_sh, do_sb, do_sh;   // element strides of batch / head inside one sequence row
  const float* vec;      // [b, hq, ceil(sq/64), 128]: per 64-query block lse*log2e | delta*scale (fa_delta_kernel)
  void* dk;              // [sk, b, hk, d] bf16
  void* dv;
  float* dk_part;        // split-heads: [b, hq, sk, d] fp32
  float* dv_part;
  const int* col_hi;     // optional [sk]: one past the LAST query that sees each key (monotone non-decreasing) — sliding windows / packed sequences; the lower
                         // edge of a key's visibility is the causal diagonal
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- per-query vectors -------------------------------------------------------------------------------------------------------------
// vec[((b*hq + h) * nq64 + i) * 128 + {0..63: lse * log2(e) | 64..127: rowsum(dO o O) * scale}] for the 64 queries of block i (zero beyond sq): the 512 bytes
// a dK/dV step needs, contiguous, so that the TMA producer fetches them with ONE bulk copy next to Q_i / dO_i (no thread of the compute warps waits on a
// global load).  One warp per (q, b, h) row of 128 bf16: lane reads 8 bytes of each tensor.
__global__ void fa_delta_kernel(const __nv_bfloat16* __restrict__ go, const __nv_bfloat16* __restrict__ o, const float* __restrict__ lse, float* __restrict__ vec, int sq,
                                int b, int hq, float scale, long go_ss, long go_sb, long go_sh, long o_ss, long o_sb, long o_sh) {
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nq64 = (sq + FB_Q - 1) / FB_Q;
  const long total = (long)nq64 * FB_Q * b * hq;
  if (row >= total) return;
  const int h = (int)(row % hq), bi = (int)((row / hq) % b), q = (int)(row / ((long)hq * b));
  float s = 0.f, l = 0.f;
  if (q < sq) {
    const uint2 a = *reinterpret_cast<const uint2*>(go + q * go_ss + bi * go_sb + h * go_sh + lane * 4);
    const uint2 c = *reinterpret_cast<const uint2*>(o + q * o_ss + bi * o_sb + h * o_sh + lane * 4);
    const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 x = __bfloat1622float2(ah[i]), y = __bfloat1622float2(ch[i]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) l = lse[((long)bi * hq + h) * sq + q] * 1.4426950408889634f;
  }
  if (lane == 0) {
    float* dst = vec + (((long)bi * hq + h) * nq64 + q / FB_Q) * (2 * FB_Q) + (q % FB_Q);
    dst[0] = l;
    dst[FB_Q] = s * scale;
  }
}

// ---- dK / dV (+ dSt spill) ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_THREADS, 1)
fa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                  const __grid_constant__ CUtensorMap tmap_do, const __grid_constant__ CUtensorMap tmap_ds, const FaBwdParams p) {
  constexpr int D = FB_D;
  constexpr int KV_CHUNK = FB_KV * 128;            // [128 rows x 128 B]
  constexpr int KV_BYTES = 2 * KV_CHUNK;           // K_j or V_j: 32 KiB
  constexpr int Q_CHUNK = FB_Q * 128;              // [64 rows x 128 B]
  constexpr int Q_BYTES = 2 * Q_CHUNK;             // Q_i or dO_i: 16 KiB
  constexpr int DS_BYTES = FB_KV * 128;            // dST bf16 [128 kv x 64 q]: 16 KiB
  constexpr int STAGES = FB_STAGES;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t DK_COL = 0, DV_COL = 128, BUF_COL = 256, BUF_STRIDE = 128, DPT_OFF = 64;   // score buffer b: ST at BUF_COL + 128 b, dPT 64 columns further

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_k = smem;
  uint8_t* smem_v = smem_k + KV_BYTES;
  uint8_t* smem_q = smem_v + KV_BYTES;                       // STAGES x Q_BYTES
  uint8_t* smem_do = smem_q + STAGES * Q_BYTES;              // STAGES x Q_BYTES
  uint8_t* smem_ds = smem_do + STAGES * Q_BYTES;             // 2 x DS_BYTES (a TMA store may still be reading the previous step's tile)
  float* smem_vec = reinterpret_cast<float*>(smem_ds + 2 * DS_BYTES);   // STAGES x (lse2[64] | delta_s[64]): arrives with Q_i / dO_i on q_full
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_vec + STAGES * 2 * FB_Q);
  uint64_t* kv_full = bars;              // 1
  uint64_t* q_full = bars + 1;           // STAGES
  uint64_t* q_empty = q_full + STAGES;   // STAGES
  uint64_t* s_ready = q_empty + STAGES;  // [2] per score buffer
  uint64_t* p_ready = s_ready + 2;       // 1 (8 arrivals = compute warps, one phase per step)
  uint64_t* acc_done = p_ready + 1;      // 1: all dV / dK MMAs have completed
  uint64_t* ds_free = acc_done + 1;      // [2]: the TMA store that was reading dS tile b has finished with it
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(ds_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = blockIdx.x, bi = blockIdx.z;
  const int group = p.hq / p.hk;
  const int hkv = p.split_heads ? (int)blockIdx.y / group : (int)blockIdx.y;
  const int h_first = p.split_heads ? (int)blockIdx.y : hkv * group;
  const int n_heads = p.split_heads ? 1 : group;
  const int k0 = jb * FB_KV;
  const int off = p.sk - p.sq;
  const int nq = (p.sq + FB_Q - 1) / FB_Q;
  int i0 = 0;
  if (p.causal) {
    const int first_q = k0 - off;          // first query that can see key k0
    i0 = first_q <= 0 ? 0 : (first_q / FB_Q) & ~1;   // even: the dQ kernel reads 128-query tiles, both 64-query halves of a visited tile must be written
  }
  int i_end = nq;
  int hi_min = 0x7fffffff;
  if (p.col_hi != nullptr) {
    hi_min = p.col_hi[k0];
    const int hi_max = p.col_hi[min(k0 + FB_KV - 1, p.sk - 1)];
    i_end = min(nq, ((hi_max + FB_Q - 1) / FB_Q + 1) & ~1);          // even again: whole 128-query tiles of the dS scratch are written
  }
  const int steps_per_head = i_end > i0 ? i_end - i0 : 0;
  const int total_steps = steps_per_head * n_heads;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_do);
    prefetch_tmap(&tmap_ds);
    mbar_init(kv_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&s_ready[i], 1);
    mbar_init(p_ready, 8);      // ONE arrival per compute warp: 256 per-thread arrivals serialise on the smem atomic unit (measured: >1,000 cycles per step)
    mbar_init(acc_done, 1);
    mbar_init(&ds_free[0], 1);
    mbar_init(&ds_free[1], 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 8) {
    // ============================== TMA producer ==========================================================================
    if (lane == 0 && total_steps > 0) {
      const int kcol = (int)(bi * p.k_sb + hkv * p.k_sh), vcol = (int)(bi * p.v_sb + hkv * p.v_sh);
      mbar_expect_tx(kv_full, 2 * KV_BYTES);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tma_load_2d(smem_k + c * KV_CHUNK, &tmap_k, kv_full, kcol + c * 64, k0);
        tma_load_2d(smem_v + c * KV_CHUNK, &tmap_v, kv_full, vcol + c * 64, k0);
      }
      for (int st = 0; st < total_steps; ++st) {
        const int h = h_first + st / steps_per_head, i = i0 + st % steps_per_head;
        const int s = st % STAGES;
        mbar_wait(&q_empty[s], ((uint32_t)(st / STAGES) & 1u) ^ 1u);
        if (p.dbg & 16) { mbar_arrive(&q_full[s]); continue; }
        mbar_expect_tx(&q_full[s], 2 * Q_BYTES + 2 * FB_Q * 4);
        bulk_load(smem_vec + s * 2 * FB_Q, p.vec + (((size_t)bi * p.hq + h) * nq + i) * (2 * FB_Q), 2 * FB_Q * 4, &q_full[s]);
        const int qcol = (int)(bi * p.q_sb + h * p.q_sh), docol = (int)(bi * p.do_sb + h * p.do_sh);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_2d(smem_q + s * Q_BYTES + c * Q_CHUNK, &tmap_q, &q_full[s], qcol + c * 64, i * FB_Q);
          tma_load_2d(smem_do + s * Q_BYTES + c * Q_CHUNK, &tmap_do, &q_full[s], docol + c * 64, i * FB_Q);
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer ================================================================================
    if (lane == 0 && total_steps > 0) {
      constexpr uint32_t idesc_st = make_idesc_bf16(FB_KV, FB_Q, false, false);   // ST, dPT : K-major A and B
      constexpr uint32_t idesc_dkv = make_idesc_bf16(FB_KV, D, false, true);      // dV, dK : A from TMEM, B MN-major
      // The issuing thread is on the critical path of every step (24 MMAs): descriptors are a precomputed low word + an immediate, never rebuilt.
      constexpr uint32_t HI = smem_desc_hi_sw128(1024);
      const uint32_t k_lo = smem_desc_lo(smem_u32(smem_k), 16), v_lo = smem_desc_lo(smem_u32(smem_v), 16);
      const uint32_t q_lo0 = smem_desc_lo(smem_u32(smem_q), 16), do_lo0 = smem_desc_lo(smem_u32(smem_do), 16);                 // K-major views (B of ST / dPT)
      const uint32_t qmn_lo0 = smem_desc_lo(smem_u32(smem_q), Q_CHUNK), domn_lo0 = smem_desc_lo(smem_u32(smem_do), Q_CHUNK);     // MN-major views (B of dK / dV)
      auto issue_scores = [&](int st) {       // ST and dPT of step st into score buffer st & 1
        const int s = st % STAGES;
        const uint32_t q_lo = q_lo0 + s * (Q_BYTES >> 4), do_lo = do_lo0 + s * (Q_BYTES >> 4);
        const uint32_t st_col = tmem_base + BUF_COL + (st & 1) * BUF_STRIDE;
        mbar_wait(&q_full[s], (uint32_t)(st / STAGES) & 1u);
        tc_fence_after();
        if (!(p.dbg & 8))
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t acc = (c > 0 || kk > 0) ? 1u : 0u;
            umma_f16(st_col, smem_desc_at(k_lo, HI, c * KV_CHUNK + kk * 32), smem_desc_at(q_lo, HI, c * Q_CHUNK + kk * 32), idesc_st, acc);
            umma_f16(st_col + DPT_OFF, smem_desc_at(v_lo, HI, c * KV_CHUNK + kk * 32), smem_desc_at(do_lo, HI, c * Q_CHUNK + kk * 32), idesc_st, acc);
          }
        umma_commit(&s_ready[st & 1]);
      };
      mbar_wait(kv_full, 0);
      issue_scores(0);
      for (int st = 0; st < total_steps; ++st) {
        const int s = st % STAGES, bsel = st & 1;
        const uint32_t qmn_lo = qmn_lo0 + s * (Q_BYTES >> 4), domn_lo = domn_lo0 + s * (Q_BYTES >> 4);
        const uint32_t st_col = tmem_base + BUF_COL + bsel * BUF_STRIDE;
        // One step ahead.  Score buffer (st+1)&1 was last used by step st-1: its ST/dPT were read by the compute warps before p_ready(st-1)
        // (waited on below in the previous iteration) and its PT/dST are operands of the dV/dK MMAs issued in that iteration — the tensor
        // pipe executes in issue order, so the score MMAs of step st+1 overwrite them only after those MMAs have consumed them.
        if (st + 1 < total_steps) issue_scores(st + 1);
        mbar_wait(p_ready, (uint32_t)st & 1u);
        tc_fence_after();
        if (!(p.dbg & 4))
#pragma unroll
        for (int kk = 0; kk < FB_Q / 16; ++kk) {
          const uint32_t acc = (st > 0 || kk > 0) ? 1u : 0u;
          // B = dO_i / Q_i read as [K = q rows, N = d]: MN-major, two 64-column chunks Q_CHUNK apart
          // PT / dST (bf16 pairs) of query half hf live in columns [32 hf, 32 hf + 16) of the ST / dPT region: each compute warp overwrote only columns it had read itself
          const uint32_t a_col = (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8);
          umma_f16_ts(tmem_base + DV_COL, st_col + a_col, smem_desc_at(domn_lo, HI, kk * 2048), idesc_dkv, acc);
          umma_f16_ts(tmem_base + DK_COL, st_col + DPT_OFF + a_col, smem_desc_at(qmn_lo, HI, kk * 2048), idesc_dkv, acc);
        }
        umma_commit(&q_empty[s]);
      }
      umma_commit(acc_done);
    }
  } else {
    // ============================== compute warps: thread = key row of ST / dPT ===================================================
    const int half = warp >> 2;                          // which 32 of the 64 query columns this warp handles
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int kv_idx = k0 + row;
    const int hi = p.col_hi != nullptr ? p.col_hi[min(kv_idx, p.sk - 1)] : 0x7fffffff;
    const int sw = row & 7;
    const int tid = threadIdx.x;                         // 0..255 among the compute warps
    // 8 columns of a half row: PT = exp2(ST c - lse), dST = PT (dPT scale - delta scale); MASKED only on tiles that touch a boundary or the causal diagonal
    auto score_math8 = [&](const uint32_t* sv, const uint32_t* dpv, uint32_t* pw, uint32_t* dw, uint32_t vec_addr, int qbase, auto masked) {
      constexpr bool MASKED = decltype(masked)::value;
#pragma unroll
      for (int c4 = 0; c4 < 8; c4 += 4) {
        const float4 l4 = lds_f4(vec_addr + c4 * 4), d4 = lds_f4(vec_addr + FB_Q * 4 + c4 * 4);
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
        float pr[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pv = ex2(fmaf(__uint_as_float(sv[c4 + e]), p.scale_log2, -lv[e]));
          if (MASKED) {
            const int q = qbase + c4 + e;
            const bool ok = kv_idx < p.sk && q < p.sq && q < hi && (!p.causal || kv_idx <= q + off);
            pv = ok ? pv : 0.f;
          }
          pr[e] = pv;
          ds[e] = pv * fmaf(__uint_as_float(dpv[c4 + e]), p.scale, -dl[e]);
        }
        __nv_bfloat162 pb0 = __floats2bfloat162_rn(pr[0], pr[1]), pb1 = __floats2bfloat162_rn(pr[2], pr[3]);
        __nv_bfloat162 db0 = __floats2bfloat162_rn(ds[0], ds[1]), db1 = __floats2bfloat162_rn(ds[2], ds[3]);
        pw[c4 / 2] = *reinterpret_cast<uint32_t*>(&pb0);
        pw[c4 / 2 + 1] = *reinterpret_cast<uint32_t*>(&pb1);
        dw[c4 / 2] = *reinterpret_cast<uint32_t*>(&db0);
        dw[c4 / 2 + 1] = *reinterpret_cast<uint32_t*>(&db1);
      }
    };
    // The 32 columns of this warp in FOUR chunks of 8, software-pipelined: the tensor-memory loads of chunk c+1 are in flight while chunk c is being computed
    // (tcgen05.wait::ld covers every load issued so far, so each wait is placed after the NEXT chunk's loads have been issued).  Tensor-memory reads are the
    // scarce resource of this kernel (64 KB of fp32 scores per step); a load-everything-then-compute order serialised them with the MUFU/FMA work.
    auto score_pipeline = [&](uint32_t st_addr, uint32_t (&pw)[16], uint32_t (&dw)[16], uint32_t vec_addr, int q0, auto masked) {
      uint32_t sv[2][8], dpv[2][8];
      tmem_ld_32x32b_x8(st_addr + half * 32, sv[0]);
      tmem_ld_32x32b_x8(st_addr + DPT_OFF + half * 32, dpv[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c < 3) {
          tmem_ld_32x32b_x8(st_addr + half * 32 + (c + 1) * 8, sv[(c + 1) & 1]);
          tmem_ld_32x32b_x8(st_addr + DPT_OFF + half * 32 + (c + 1) * 8, dpv[(c + 1) & 1]);
        }
        score_math8(sv[c & 1], dpv[c & 1], pw + c * 4, dw + c * 4, vec_addr + c * 32, q0 + half * 32 + c * 8, masked);
      }
    };
    int h = h_first, i = i0;                             // (head, query block) of the current step, advanced incrementally
    for (int st = 0; st < total_steps; ++st) {
      const int q0 = i * FB_Q;
      const int s = st % STAGES;
      const uint32_t st_addr = tmem_base + lane_base + BUF_COL + (st & 1) * BUF_STRIDE;
      uint8_t* ds_row = smem_ds + (st & 1) * DS_BYTES + row * 128;
      mbar_wait(&q_full[s], (uint32_t)(st / STAGES) & 1u);       // acquire on the barrier the bulk copy of this step's vectors completed on (long since satisfied)
      mbar_wait(&s_ready[st & 1], (uint32_t)(st >> 1) & 1u);
      tc_fence_after();
      uint32_t pw[16], dw[16];
      {
        const bool interior = (k0 + FB_KV <= p.sk) && (q0 + FB_Q <= p.sq) && (q0 + FB_Q <= hi_min) && (!p.causal || (k0 + FB_KV - 1 <= q0 + off));
        const uint32_t vec_addr = smem_u32(smem_vec + s * 2 * FB_Q) + half * 32 * 4;
        if (p.dbg & 1) {
#pragma unroll
          for (int e = 0; e < 16; ++e) pw[e] = dw[e] = (uint32_t)(st + e);
        } else if (interior) score_pipeline(st_addr, pw, dw, vec_addr, q0, std::false_type{});
        else score_pipeline(st_addr, pw, dw, vec_addr, q0, std::true_type{});
      }
      // PT and dST (bf16 pairs) back over the first 16 of the 32 columns this warp has just read (no other warp touches them); dST also into the swizzled smem
      // tile that leaves by TMA
      if (!(p.dbg & 64)) {
        tmem_st_32x32b_x16(st_addr + half * 32, pw);
        tmem_st_32x32b_x16(st_addr + DPT_OFF + half * 32, dw);
      }
      // the TMA store of step st-2 must have finished READING this dS tile before it is overwritten (flag set by thread 0 one step ago)
      if (st >= 2) mbar_wait(&ds_free[st & 1], (uint32_t)((st >> 1) - 1) & 1u);
      if (!(p.dbg & 512))
#pragma unroll
      for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(ds_row + (((half * 4 + u) ^ sw) << 4)) = make_uint4(dw[u * 4], dw[u * 4 + 1], dw[u * 4 + 2], dw[u * 4 + 3]);
      if (!(p.dbg & 128)) tmem_st_wait();
      if (!(p.dbg & 256)) fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      if (tid == 0) {
        // all 256 writers have fenced and arrived: the tile is complete and visible to the async proxy
        mbar_wait(p_ready, (uint32_t)st & 1u);
        if (!(p.dbg & 2)) tma_store_3d(&tmap_ds, smem_ds + (st & 1) * DS_BYTES, q0, k0, bi * p.hq + h);
        tma_store_commit();
        if (st >= 1) {
          tma_store_wait_read<1>();                    // store(st-1) has read its tile: release it for step st+1
          mbar_arrive(&ds_free[(st + 1) & 1]);
        }
      }
      if (++i == i_end) { i = i0; ++h; }
    }
    // ---- epilogue: dK, dV [kv = row, d] ------------------------------------------------------------------------------------------
    if (total_steps > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
      const bool valid = kv_idx < p.sk;
      if (!p.split_heads) {
        __nv_bfloat16* dk_row = reinterpret_cast<__nv_bfloat16*>(p.dk) + ((size_t)kv_idx * p.b * p.hk + (size_t)bi * p.hk + hkv) * D;
        __nv_bfloat16* dv_row = reinterpret_cast<__nv_bfloat16*>(p.dv) + ((size_t)kv_idx * p.b * p.hk + (size_t)bi * p.hk + hkv) * D;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
#pragma unroll 1
          for (int ch = half * (D / 64); ch < (half + 1) * (D / 64); ++ch) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + (which ? DV_COL : DK_COL) + ch * 32, o);
            tmem_ld_wait();
            if (valid) {
              __nv_bfloat16* dst = (which ? dv_row : dk_row) + ch * 32;
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                uint32_t v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(o[g * 16 + 2 * e]), __uint_as_float(o[g * 16 + 2 * e + 1]));
                  v[e] = *reinterpret_cast<uint32_t*>(&hb);
                }
                st_global_v8(dst + g * 16, v);
              }
            }
          }
        }
      } else {
        float* dk_row = p.dk_part + (((size_t)bi * p.hq + h_first) * p.sk + kv_idx) * D;
        float* dv_row = p.dv_part + (((size_t)bi * p.hq + h_first) * p.sk + kv_idx) * D;
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
#pragma unroll 1
          for (int ch = half * (D / 64); ch < (half + 1) * (D / 64); ++ch) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + (which ? DV_COL : DK_COL) + ch * 32, o);
            tmem_ld_wait();
            if (valid) {
              float* dst = (which ? dv_row : dk_row) + ch * 32;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = o[g * 8 + e];
                st_global_v8(dst + g * 8, v);
              }
            }
          }
        }
      }
    } else if (kv_idx < p.sk) {
      // no query sees this key block (cannot happen with sk >= sq causal, kept for safety): gradients are zero
      if (!p.split_heads) {
        uint4 z = make_uint4(0, 0, 0, 0);
        __nv_bfloat16* dk_row = reinterpret_cast<__nv_bfloat16*>(p.dk) + ((size_t)kv_idx * p.b * p.hk + (size_t)bi * p.hk + hkv) * D;
        __nv_bfloat16* dv_row = reinterpret_cast<__nv_bfloat16*>(p.dv) + ((size_t)kv_idx * p.b * p.hk + (size_t)bi * p.hk + hkv) * D;
        for (int c = half * (D / 16); c < (half + 1) * (D / 16); ++c) {
          reinterpret_cast<uint4*>(dk_row)[c] = z;
          reinterpret_cast<uint4*>(dv_row)[c] = z;
        }
      } else {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float* dk_row = p.dk_part + (((size_t)bi * p.hq + h_first) * p.sk + kv_idx) * D;
        float* dv_row = p.dv_part + (((size_t)bi * p.hq + h_first) * p.sk + kv_idx) * D;
        for (int c = half * (D / 8); c < (half + 1) * (D / 8); ++c) {
          reinterpret_cast<float4*>(dk_row)[c] = z;
          reinterpret_cast<float4*>(dv_row)[c] = z;
        }
      }
    }
    if (tid == 0) tma_store_wait<0>();     // the scratch writes must be complete before the kernel ends (the dQ kernel reads them)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// split-heads mode: dK/dV = sum over the query heads of each GQA group of the fp32 partials -> bf16 [sk, b, hk, d]
__global__ void fa_bwd_reduce_kernel(const float* __restrict__ dk_part, const float* __restrict__ dv_part, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                                     int sk, int b, int hq, int hk) {
  const int group = hq / hk;
  const long n4 = (long)sk * b * hk * (FB_D / 4);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (FB_D / 4));
    long r = i / (FB_D / 4);
    const int hkv = (int)(r % hk);
    r /= hk;
    const int bi = (int)(r % b);
    const int kv = (int)(r / b);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), v = a;
    for (int g = 0; g < group; ++g) {
      const size_t src = ((((size_t)bi * hq + hkv * group + g) * sk + kv) * FB_D) / 4 + c;
      const float4 x = reinterpret_cast<const float4*>(dk_part)[src], y = reinterpret_cast<const float4*>(dv_part)[src];
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
    }
    __nv_bfloat162 k0 = __floats2bfloat162_rn(a.x, a.y), k1 = __floats2bfloat162_rn(a.z, a.w), v0 = __floats2bfloat162_rn(v.x, v.y), v1 = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(dk)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&k0), *reinterpret_cast<uint32_t*>(&k1));
    reinterpret_cast<uint2*>(dv)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&v0), *reinterpret_cast<uint32_t*>(&v1));
  }
}

// ---- dQ = dS K ------------------------------------------------------------------------------------------------------------------------
struct FaDqParams {
  int sq, sk, b, hq, hk, causal;
  long k_sb, k_sh;
  void* dq;              // [sq, b, hq, d] bf16 contiguous
  const int* row_lo;     // optional [sq]: first visible key of every query (band masks): key blocks before it were never written to the dS scratch
};
constexpr int DQ_BM = 128, DQ_BK = 128, DQ_STAGES = 3;
constexpr int DQ_A_BYTES = DQ_BK * DQ_BM * 2;     // dSt tile: [128 keys][128 q] bf16 as two 64-q chunks of [128 rows x 128 B]
constexpr int DQ_B_BYTES = DQ_BK * FB_D * 2;      // K tile:   [128 keys][128 d]
constexpr int DQ_STAGE_BYTES = DQ_A_BYTES + DQ_B_BYTES;
static_assert(DQ_A_BYTES == DQ_B_BYTES, "the dQ issuer derives the B descriptor from the A descriptor");

__global__ void __launch_bounds__(NUM_THREADS, 1)
fa_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_ds, const __grid_constant__ CUtensorMap tmap_k, const FaDqParams p) {
  constexpr uint32_t TMEM_COLS = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_STAGES * DQ_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + DQ_STAGES;
  uint64_t* tmem_full = bars + 2 * DQ_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = (p.sq + DQ_BM - 1) / DQ_BM, nkb = (p.sk + DQ_BK - 1) / DQ_BK;
  const int heads = p.b * p.hq;
  const int num_tiles = nqb * heads;
  const int off = p.sk - p.sq;
  const int group = p.hq / p.hk;
  // tile index -> (query block, head): latest (= heaviest under a causal mask) query blocks first
  auto tile_of = [&](int t, int& qb, int& bh, int& kb_lo, int& n_kb) {
    qb = nqb - 1 - t / heads;
    bh = t % heads;
    kb_lo = p.row_lo != nullptr ? max(p.row_lo[qb * DQ_BM], 0) / DQ_BK : 0;
    if (p.causal) {
      const int last = min(qb * DQ_BM + DQ_BM - 1, p.sq - 1) + off;
      n_kb = last < 0 ? 0 : min(nkb, last / DQ_BK + 1);
    } else {
      n_kb = nkb;
    }
    kb_lo = min(kb_lo, n_kb);
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_ds);
    prefetch_tmap(&tmap_k);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < DQ_STAGES; ++i) {
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
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int qb, bh, kb_lo, n_kb;
        tile_of(t, qb, bh, kb_lo, n_kb);
        const int bi = bh / p.hq, h = bh % p.hq;
        const int kcol = (int)(bi * p.k_sb + (h / group) * p.k_sh);
        for (int kb = kb_lo; kb < n_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], DQ_STAGE_BYTES);
          uint8_t* sa = smem + stage * DQ_STAGE_BYTES;
          uint8_t* sb = sa + DQ_A_BYTES;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tma_load_3d(sa + c * (DQ_A_BYTES / 2), &tmap_ds, &full_bar[stage], qb * DQ_BM + c * 64, kb * DQ_BK, bh);
            tma_load_2d(sb + c * (DQ_B_BYTES / 2), &tmap_k, &full_bar[stage], kcol + c * 64, kb * DQ_BK);
          }
          if (++stage == DQ_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(DQ_BM, FB_D, true, true);   // A = dSt tile (M = q contiguous), B = K tile (N = d contiguous)
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int qb, bh, kb_lo, n_kb;
        tile_of(t, qb, bh, kb_lo, n_kb);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * FB_D;
        for (int kb = kb_lo; kb < n_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(smem_u32(smem + stage * DQ_STAGE_BYTES), DQ_A_BYTES / 2), b_lo = a_lo + (DQ_A_BYTES >> 4);   // same LBO for both (DQ_A_BYTES == DQ_B_BYTES)
#pragma unroll
          for (int kk = 0; kk < DQ_BK / UMMA_K; ++kk)
            umma_f16(d_tmem, smem_desc_at(a_lo, smem_desc_hi_sw128(1024), kk * 2048), smem_desc_at(b_lo, smem_desc_hi_sw128(1024), kk * 2048), idesc, (kb > kb_lo || kk > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == DQ_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);      // with n_kb == kb_lo this completes at once and the epilogue writes zeros
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int qb, bh, kb_lo, n_kb;
      tile_of(t, qb, bh, kb_lo, n_kb);
      const int bi = bh / p.hq, h = bh % p.hq;
      const int q = qb * DQ_BM + ew * 32 + lane;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.dq) + (((size_t)q * p.b + bi) * p.hq + h) * FB_D + half * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * FB_D + ((uint32_t)(ew * 32) << 16) + half * 64 + c * 32, r);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (q < p.sq) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float x0 = n_kb > kb_lo ? __uint_as_float(r[g * 16 + 2 * e]) : 0.f, x1 = n_kb > kb_lo ? __uint_as_float(r[g * 16 + 2 * e + 1]) : 0.f;
              __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);
              v[e] = *reinterpret_cast<uint32_t*>(&hb);
            }
            st_global_v8(dst + c * 32 + g * 16, v);
          }
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

// bytes of scratch the caller must provide: dSt [b*hq, sk, sq_al] bf16 (+ fp32 dK/dV partials [b, hq, sk, d] x 2 in split-heads mode)
extern "C" size_t mb200_flash_attn_bwd_scratch_bytes(int sq, int sk, int b, int hq, int hk, int split_heads) {
  const size_t sq_al = ((size_t)sq + 7) / 8 * 8;
  size_t n = (size_t)b * hq * sk * sq_al * 2;
  n = (n + 255) / 256 * 256;
  if (split_heads) n += 2 * (size_t)b * hq * sk * FB_D * 4;
  return n;
}

// split-heads (one CTA per key block and QUERY head, fp32 dK/dV partials reduced afterwards) measured faster than one CTA per KV head at every head
// count of the Llama-3 8B shape (32/8: 1.76 vs 1.86 ms; 4/1: 0.32 vs 0.69 ms): 4x the CTAs balance the causal tail and keep all SMs busy under TP
extern "C" int mb200_flash_attn_bwd_split_heads(int sk, int b, int hq, int hk) {
  (void)sk; (void)b;
  return hq > hk ? 1 : 0;
}

// q, do, o: [sq, b, hq, 128]; k, v: [sk, b, hk, 128] (element strides given, d contiguous); lse: [b, hq, sq] fp32;
// vec: [b, hq, ceil(sq/64), 128] fp32 scratch (filled here); dq: [sq, b, hq, 128] bf16 contiguous; dk, dv: [sk, b, hk, 128] bf16 contiguous.
extern "C" int mb200_flash_attn_bwd(const void* q, const void* k, const void* v, const void* dout, const void* out, const float* lse, float* vec, void* dq, void* dk,
                                    void* dv, void* scratch, int split_heads, int sq, int sk, int b, int hq, int hk, int d, long q_ss, long q_sb, long q_sh, long k_ss,
                                    long k_sb, long k_sh, long v_ss, long v_sb, long v_sh, long do_ss, long do_sb, long do_sh, long o_ss, long o_sb, long o_sh, float scale,
                                    int causal, const int* row_lo, const int* col_hi, cudaStream_t s) {
  if (d != FB_D || hq % hk != 0) return -10;
  if ((q_ss | q_sb | q_sh | k_ss | k_sb | k_sh | v_ss | v_sb | v_sh | do_ss | do_sb | do_sh | o_ss | o_sb | o_sh) % 8 != 0) return -11;
  constexpr int SMEM_DKV = 2 * (2 * FB_KV * 128) + FB_STAGES * 2 * (2 * FB_Q * 128) + 2 * FB_KV * 128 + FB_STAGES * 2 * FB_Q * 4 + 1024 + 256;
  constexpr int SMEM_DQ = DQ_STAGES * DQ_STAGE_BYTES + 1024 + 256;
  const size_t sq_al = ((size_t)sq + 7) / 8 * 8;
  CUtensorMap tq, tk, tv, tdo, tds, tk2;
  if (!make_tmap_bf16_strided(&tq, q, sq, q_ss, q_ss * 2, 64, FB_Q)) return -21;
  if (!make_tmap_bf16_strided(&tk, k, sk, k_ss, k_ss * 2, 64, FB_KV)) return -22;
  if (!make_tmap_bf16_strided(&tv, v, sk, v_ss, v_ss * 2, 64, FB_KV)) return -23;
  if (!make_tmap_bf16_strided(&tdo, dout, sq, do_ss, do_ss * 2, 64, FB_Q)) return -24;
  if (!make_tmap_bf16_3d(&tds, scratch, (uint64_t)sq, (uint64_t)sk, (uint64_t)b * hq, sq_al * 2, (uint64_t)sk * sq_al * 2, 64, FB_KV)) return -25;
  if (!make_tmap_bf16_strided(&tk2, k, sk, k_ss, k_ss * 2, 64, DQ_BK)) return -26;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(fa_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV) != cudaSuccess) return -3;
    if (cudaFuncSetAttribute(fa_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ) != cudaSuccess) return -3;
    configured = true;
  }
  {
    const long rows = (long)((sq + FB_Q - 1) / FB_Q) * FB_Q * b * hq;
    const int wpb = 8;
    fa_delta_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(out), lse, vec,
                                                                          sq, b, hq, scale, do_ss, do_sb, do_sh, o_ss, o_sb, o_sh);
  }
  FaBwdParams p;
  p.sq = sq; p.sk = sk; p.b = b; p.hq = hq; p.hk = hk; p.causal = causal; p.split_heads = split_heads;
  {
    static const int dbg = getenv("MB200_FA_BWD_DBG") ? atoi(getenv("MB200_FA_BWD_DBG")) : 0;
    p.dbg = dbg;
  }
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.q_sb = q_sb; p.q_sh = q_sh; p.k_sb = k_sb; p.k_sh = k_sh; p.v_sb = v_sb; p.v_sh = v_sh; p.do_sb = do_sb; p.do_sh = do_sh;
  p.vec = vec; p.dk = dk; p.dv = dv; p.col_hi = col_hi;
  size_t ds_bytes = (size_t)b * hq * sk * sq_al * 2;
  ds_bytes = (ds_bytes + 255) / 256 * 256;
  p.dk_part = split_heads ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + ds_bytes) : nullptr;
  p.dv_part = split_heads ? p.dk_part + (size_t)b * hq * sk * FB_D : nullptr;
  dim3 grid((sk + FB_KV - 1) / FB_KV, split_heads ? hq : hk, b);
  fa_bwd_dkv_kernel<<<grid, FB_THREADS, SMEM_DKV, s>>>(tq, tk, tv, tdo, tds, p);
  if (split_heads) fa_bwd_reduce_kernel<<<num_sms() * 4, 256, 0, s>>>(p.dk_part, p.dv_part, reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv), sk, b, hq, hk);
  FaDqParams dqp;
  dqp.sq = sq; dqp.sk = sk; dqp.b = b; dqp.hq = hq; dqp.hk = hk; dqp.causal = causal; dqp.k_sb = k_sb; dqp.k_sh = k_sh; dqp.dq = dq; dqp.row_lo = row_lo;
  const int tiles = ((sq + DQ_BM - 1) / DQ_BM) * b * hq;
  fa_bwd_dq_kernel<<<tiles < num_sms() ? tiles : num_sms(), NUM_THREADS, SMEM_DQ, s>>>(tds, tk2, dqp);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
