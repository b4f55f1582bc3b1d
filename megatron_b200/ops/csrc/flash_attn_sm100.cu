// Flash attention forward for sm_100a: tcgen05 MMAs with S and O accumulators in tensor memory, TMA-fed K/V ring,
// two 128-row query tiles per CTA whose softmax warpgroups ping-pong against ONE MMA-issuing thread.
//
//   warps 0-3  : softmax warpgroup of query tile 0 (thread = one query row = one TMEM lane)
//   warps 4-7  : softmax warpgroup of query tile 1
//   warp  8    : TMA producer (Q once, then K_j / V_j through a 3-stage ring each)
//   warps 9,10 : MMA issuers, one elected thread per query tile (independent issue streams: no head-of-line blocking between
//                the two tiles' P-ready hand-offs); warp 9 also owns the TMEM allocation
//
// Per KV block j (64 keys) and tile t:     S_t = Q_t K_jᵀ   (UMMA 128x64x16, K-major A and B from smem, D in TMEM)
//                                          softmax WG t: TMEM → regs, online softmax in log2 domain, P_t(bf16) → swizzled smem
//                                          O_t += P_t V_j   (UMMA 128xDx16, B = V tile MN-major straight from the [kv, d] layout)
// S_t and P_t are double-buffered and the issuer runs one block ahead:  … PV(t,j) ; QK(t,j+2) ; commit(s_full[t][j%2]) …, so the
// softmax of block j+1 never waits for a tensor-core round trip, and "S(t,j+2) ready" implies "PV(t,j) complete" (its P buffer may be
// overwritten).  O is rescaled lazily — only when the running max grows by more than 2^8 (the stale reference max is used otherwise;
// exact after the final 1/l) — and a rescale at block j first waits for PV(t,j-1) on one of FOUR rotating pv_done barriers (a parity wait is
// only unambiguous within one phase of lag; with four barriers a rarely-taken wait can never alias), then happens before p_ready(j).
//
// Layouts: q [sq, b, hq, d], k/v [sk, b, hk, d] with arbitrary (16-byte aligned) s/b/h strides and contiguous d — the k/v
// views produced by splitting a fused QKV projection are consumed in place.  out [sq, b, hq, d] contiguous,
// lse [b, hq, sq] fp32 (natural log).  Causal masking is bottom-right aligned (kv ≤ q + sk − sq).
// Replaces: TE DotProductAttention / cuDNN fused attention and the unfused baddbmm+softmax+bmm path (X6).
#include <cstdlib>

#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int FA_BM = 128;       // query rows per tile
constexpr int FA_BN = 64;        // keys per block
constexpr int FA_STAGES = 3;     // K and V ring depth
constexpr int FA_THREADS = 352;   // 8 softmax warps + TMA producer + one MMA-issuing warp per query tile
constexpr float FA_RESCALE_THRESHOLD = 8.0f;  // log2 units
// ncu (profiles/r1_flash_attn_ncu.md): with one issuer per tile the softmax warps are ISSUE-bound (50 % issue slots, MUFU only 34 % busy), so the
// 8-instruction software exp2 costs more than the MUFU slot it saves; kept for the regime where MUFU saturates (e.g. smaller head dims).
constexpr bool FA_POLY_EXP2 = false;

struct FaParams {
  int sq, sk, b, hq, hk;
  int causal;
  int pair_mode;                 // 0: a CTA owns two neighbouring query tiles, 1: mirrored tiles (x, T-1-x) — see the kernel
  float scale_log2;              // softmax_scale * log2(e)
  long q_sb, q_sh, k_sb, k_sh, v_sb, v_sh;  // element strides of batch / head inside one sequence row
  void* out;
  long o_pitch;                  // elements between consecutive query rows of out
  float* lse;
  const int* row_lo;             // optional [sq]: first visible key of every query row (monotone non-decreasing) — sliding windows and packed sequences
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA/ALU pipes (degree-3 polynomial after magic-number range reduction; rel. error ~6e-4, below bf16 resolution of P).
// The MUFU unit does 16 ex2/clk/SM while a 128x64 S tile needs 8192 of them — as long as the two MMAs of that tile.  Evaluating a
// quarter of the exponentials here takes that work off the critical unit (same idea as FlashAttention-4's software exp2).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float r = x + 12582912.f;                 // 1.5 * 2^23: integer part of x lands in the low mantissa bits (round to nearest)
  const float n = r - 12582912.f;
  const float f = x - n;                          // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// P_TMEM: the bf16 probability tile is written back over the first 32 columns of its S buffer with tcgen05.st and the PV MMA reads operand A from
// tensor memory (no shared-memory round trip, no proxy fence, half the smem operand traffic of PV); otherwise P goes through a swizzled smem tile.
template <int D, bool P_TMEM>
__global__ void __launch_bounds__(FA_THREADS, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v, const FaParams p) {
  constexpr int DCH = D / 64;                          // 64-column (128-byte) chunks of the head dim
  constexpr int Q_CHUNK_BYTES = FA_BM * 128;           // [128 rows x 128 B]
  constexpr int Q_TILE_BYTES = DCH * Q_CHUNK_BYTES;
  constexpr int KV_CHUNK_BYTES = FA_BN * 128;          // [64 rows x 128 B]
  constexpr int KV_STAGE_BYTES = DCH * KV_CHUNK_BYTES;
  constexpr int P_TILE_BYTES = FA_BM * 128;            // [128 rows x 64 bf16]
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t S_COL = 0, O_COL = 256;           // S[t][b] at (2t+b)*64, O_t at 256 + t*D

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                        // 2 tiles
  uint8_t* smem_k = smem_q + 2 * Q_TILE_BYTES;                   // FA_STAGES
  uint8_t* smem_v = smem_k + FA_STAGES * KV_STAGE_BYTES;         // FA_STAGES
  uint8_t* smem_p = smem_v + FA_STAGES * KV_STAGE_BYTES;         // 2 tiles x 2 buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 4 * P_TILE_BYTES);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // FA_STAGES
  uint64_t* k_empty = k_full + FA_STAGES;
  uint64_t* v_full = k_empty + FA_STAGES;
  uint64_t* v_empty = v_full + FA_STAGES;
  uint64_t* s_full = v_empty + FA_STAGES;        // [tile][buffer] = 4
  uint64_t* p_ready = s_full + 4;                // [tile][buffer] = 4
  uint64_t* pv_done = p_ready + 4;               // [tile][j % 4] = 8: PV(t, j) commits to barrier j % 4 (phase j / 4)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;   // heaviest (latest) causal tiles are scheduled first
  // Which two 128-row query tiles this CTA owns.  pair_mode 0: neighbours (2 qt, 2 qt + 1) — both tiles need (almost) the same K/V blocks, the ping-pong
  // runs for the whole walk; with several waves of CTAs, heaviest first.  pair_mode 1 ("mirrored", for grids of at most ~one wave, e.g. 4 local heads
  // under TP=8): tiles (x, T-1-x) — every CTA then owns T+1 causal blocks instead of 2..2T, so the single wave is balanced (the heavy tile finishes solo).
  const int n_tiles = (p.sq + FA_BM - 1) / FA_BM;
  int tile_row0[2];
  if (p.pair_mode == 0) {
    tile_row0[0] = qt * 2 * FA_BM;
    tile_row0[1] = tile_row0[0] + FA_BM;
  } else {
    tile_row0[0] = (int)blockIdx.x * FA_BM;
    tile_row0[1] = (n_tiles - 1 - (int)blockIdx.x) * FA_BM;
    if (tile_row0[1] == tile_row0[0]) tile_row0[1] = p.sq;      // odd tile count: the middle CTA owns one tile
  }
  const int h = blockIdx.y, bi = blockIdx.z;
  const int hkv = h / (p.hq / p.hk);
  const int off = p.sk - p.sq;                           // bottom-right causal alignment
  const int nkv = (p.sk + FA_BN - 1) / FA_BN;
  int n_t[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int first = tile_row0[t];
    if (first >= p.sq) {
      n_t[t] = 0;
    } else if (p.causal) {
      const int last = min(first + FA_BM - 1, p.sq - 1) + off;
      n_t[t] = last < 0 ? 0 : min(nkv, last / FA_BN + 1);
    } else {
      n_t[t] = nkv;
    }
  }
  // Band masks (sliding window / packed sequences): keys before row_lo[q] are hidden.  row_lo is monotone, so the CTA's walk starts at the block holding the
  // first visible key of its first row; every index below is RELATIVE to that block (jlo) except the K/V coordinates and the mask.
  int jlo = 0;
  if (p.row_lo != nullptr && tile_row0[0] < p.sq) jlo = min(max(p.row_lo[tile_row0[0]], 0) / FA_BN, max(n_t[0] - 1, 0));   // tile 0 is the earlier one (pair_mode 0 with a band)
  n_t[0] = max(n_t[0] - jlo, 0);
  n_t[1] = max(n_t[1] - jlo, 0);
  const int n = max(n_t[0], n_t[1]);

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);   // one arrival per tile's issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], FA_BM / 32);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&pv_done[i], 1);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 8) {
    // ================================= TMA producer =====================================================
    if (lane == 0 && n > 0) {
      const int qcol = (int)(bi * p.q_sb + h * p.q_sh);
      const int kcol = (int)(bi * p.k_sb + hkv * p.k_sh);
      const int vcol = (int)(bi * p.v_sb + hkv * p.v_sh);
      mbar_expect_tx(q_full, 2 * Q_TILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_2d(smem_q + t * Q_TILE_BYTES + c * Q_CHUNK_BYTES, &tmap_q, q_full, qcol + c * 64, min(tile_row0[t], p.sq - 1));
      for (int j = 0; j < n; ++j) {
        const int s = j % FA_STAGES;
        const uint32_t ph = (uint32_t)(j / FA_STAGES) & 1u;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], KV_STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_2d(smem_k + s * KV_STAGE_BYTES + c * KV_CHUNK_BYTES, &tmap_k, &k_full[s], kcol + c * 64, (j + jlo) * FA_BN);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], KV_STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) tma_load_2d(smem_v + s * KV_STAGE_BYTES + c * KV_CHUNK_BYTES, &tmap_v, &v_full[s], vcol + c * 64, (j + jlo) * FA_BN);
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ================================= MMA issuer ===============================================================
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(FA_BM, FA_BN, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(FA_BM, D, false, true);
      // descriptors = a precomputed low word + an immediate (the issuing threads are on the critical path of every KV block)
      constexpr uint32_t HI = smem_desc_hi_sw128(1024);
      const uint32_t q_lo0 = smem_desc_lo(smem_u32(smem_q), 16), k_lo0 = smem_desc_lo(smem_u32(smem_k), 16), p_lo0 = smem_desc_lo(smem_u32(smem_p), 16);
      const uint32_t v_lo0 = smem_desc_lo(smem_u32(smem_v), KV_CHUNK_BYTES);
      auto issue_qk = [&](int t, int s, int buf) {
        const uint32_t q_lo = q_lo0 + t * (Q_TILE_BYTES >> 4), k_lo = k_lo0 + s * (KV_STAGE_BYTES >> 4);
#pragma unroll
        for (int c = 0; c < DCH; ++c)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base + S_COL + (2 * t + buf) * FA_BN, smem_desc_at(q_lo, HI, c * Q_CHUNK_BYTES + kk * 32), smem_desc_at(k_lo, HI, c * KV_CHUNK_BYTES + kk * 32),
                     idesc_qk, (c > 0 || kk > 0) ? 1u : 0u);
      };
      auto issue_pv = [&](int t, int s, int buf, bool acc) {
        const uint32_t p_lo = p_lo0 + (2 * t + buf) * (P_TILE_BYTES >> 4), v_lo = v_lo0 + s * (KV_STAGE_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < FA_BN / 16; ++kk) {
          if (P_TMEM)
            umma_f16_ts(tmem_base + O_COL + t * D, tmem_base + S_COL + (2 * t + buf) * FA_BN + kk * 8, smem_desc_at(v_lo, HI, kk * 2048), idesc_pv, (acc || kk > 0) ? 1u : 0u);
          else
            umma_f16(tmem_base + O_COL + t * D, smem_desc_at(p_lo, HI, kk * 32), smem_desc_at(v_lo, HI, kk * 2048), idesc_pv, (acc || kk > 0) ? 1u : 0u);
        }
      };
      const int t = warp - 9;
      const int nt = n_t[t];
      mbar_wait(q_full, 0);
      // prologue: S(t,0) and S(t,1) are issued up front so the softmax warpgroup always has a block waiting
      for (int j0 = 0; j0 < 2 && j0 < n; ++j0) {
        mbar_wait(&k_full[j0], 0);
        tc_fence_after();
        if (j0 < nt) {
          issue_qk(t, j0, j0);
          umma_commit(&s_full[2 * t + j0]);
          umma_commit(&k_empty[j0]);
        } else {
          mbar_arrive(&k_empty[j0]);           // block not needed by this tile (causal): release our share of the stage
        }
      }
      for (int j = 0; j < n; ++j) {
        const int s = j % FA_STAGES, s2 = (j + 2) % FA_STAGES, buf = j & 1;
        mbar_wait(&v_full[s], (uint32_t)(j / FA_STAGES) & 1u);
        if (j + 2 < n) mbar_wait(&k_full[s2], (uint32_t)((j + 2) / FA_STAGES) & 1u);
        if (j < nt) {
          mbar_wait(&p_ready[2 * t + buf], (uint32_t)(j >> 1) & 1u);
          tc_fence_after();
          issue_pv(t, s, buf, j > 0);
          umma_commit(&pv_done[4 * t + (j & 3)]);
          umma_commit(&v_empty[s]);
        } else {
          mbar_arrive(&v_empty[s]);
        }
        if (j + 2 < n) {
          if (j + 2 < nt) {
            issue_qk(t, s2, buf);              // S[t][buf] was drained before p_ready(t,j) was signalled
            umma_commit(&s_full[2 * t + buf]);
            umma_commit(&k_empty[s2]);
          } else {
            mbar_arrive(&k_empty[s2]);
          }
        }
      }
    }
  } else {
    // ================================= softmax warpgroups ================================================================
    const int t = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const int q_idx = tile_row0[t] + row;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_addr0 = tmem_base + lane_base + S_COL + 2 * t * FA_BN;
    const uint32_t o_addr = tmem_base + lane_base + O_COL + t * D;
    uint8_t* p_row0 = smem_p + 2 * t * P_TILE_BYTES + row * 128;
    const int sw = row & 7;
    float m_ref = -INFINITY, l = 0.f;
    const int nb = n_t[t];
    const int lo = (p.row_lo != nullptr && q_idx < p.sq) ? p.row_lo[q_idx] : 0;   // first visible key of this row
    for (int j = 0; j < nb; ++j) {
      const int buf = j & 1;
      const uint32_t s_addr = s_addr0 + buf * FA_BN;
      uint8_t* p_row = p_row0 + buf * P_TILE_BYTES;
      mbar_wait(&s_full[2 * t + buf], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(s_addr, r0);
      tmem_ld_32x32b_x32(s_addr + 32, r1);
      tmem_ld_wait();
      const int kv0 = (j + jlo) * FA_BN;
      const int limit = p.causal ? min(p.sk - 1, q_idx + off) : p.sk - 1;   // last visible key of this row
      if (kv0 + FA_BN - 1 > limit) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (kv0 + c > limit) r0[c] = 0xff800000u;        // -inf
          if (kv0 + 32 + c > limit) r1[c] = 0xff800000u;
        }
      }
      if (kv0 < lo) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (kv0 + c < lo) r0[c] = 0xff800000u;
          if (kv0 + 32 + c < lo) r1[c] = 0xff800000u;
        }
      }
      // row max with 3-input FMNMX3 (sm_100) in 8 independent chains: 32 instructions for 64 scores, no long dependent chain
      float mxs[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = fmax3(__uint_as_float(r0[c]), __uint_as_float(r1[c]), __uint_as_float(r0[c + 8]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = fmax3(mxs[c], __uint_as_float(r1[c + 8]), __uint_as_float(r0[c + 16]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = fmax3(mxs[c], __uint_as_float(r1[c + 16]), __uint_as_float(r0[c + 24]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = fmaxf(mxs[c], __uint_as_float(r1[c + 24]));
      const float mx = fmaxf(fmax3(mxs[0], mxs[1], mxs[2]), fmaxf(fmax3(mxs[3], mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      const float mx_scaled = mx * p.scale_log2;
      float alpha = 1.f;
      if (mx_scaled > m_ref + FA_RESCALE_THRESHOLD) {
        alpha = fast_exp2(m_ref - mx_scaled);               // 0 on the first block (m_ref = -inf)
        m_ref = mx_scaled;
        l *= alpha;
      }
      const bool rescale = __any_sync(0xffffffffu, alpha != 1.f) && j > 0;
      // a row whose band starts in a later block has seen only -inf so far: exponentiate against 0 instead of -inf (−inf − (−inf) is NaN); every p is then 0
      const float m_sub = (m_ref == -INFINITY) ? 0.f : m_ref;
      // p = 2^(s*scale - m_ref); written as bf16 into the 128B-swizzled K-major tile the PV MMA reads as operand A
      float sums[8];
      uint32_t pw[32];   // the row's 64 probabilities as packed bf16 pairs (P_TMEM path)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t pk[4];
        float su = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = u * 8 + e * 2;
          const float a0 = __uint_as_float(c < 32 ? r0[c] : r1[c - 32]);
          const float a1 = __uint_as_float(c + 1 < 32 ? r0[c + 1] : r1[c + 1 - 32]);
          const float x0 = fmaf(a0, p.scale_log2, -m_sub), x1 = fmaf(a1, p.scale_log2, -m_sub);
          const float p0 = fast_exp2(x0);
          const float p1 = (FA_POLY_EXP2 && (e & 1)) ? poly_exp2(x1) : fast_exp2(x1);   // optionally every 4th exponential on the FMA pipe
          su += p0 + p1;
          __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
          pk[e] = *reinterpret_cast<uint32_t*>(&hb);
        }
        sums[u] = su;
        if (P_TMEM) {
#pragma unroll
          for (int e = 0; e < 4; ++e) pw[u * 4 + e] = pk[e];
        } else {
          *reinterpret_cast<uint4*>(p_row + ((u ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      if (P_TMEM) {
        tmem_st_32x32b_x32(s_addr, pw);   // P(j) overwrites columns 0..31 of S[t][buf]; this thread has already read its whole S row
        tmem_st_wait();
      }
      l += ((sums[0] + sums[1]) + (sums[2] + sums[3])) + ((sums[4] + sums[5]) + (sums[6] + sums[7]));
      if (rescale) {
        // rare path: wait until PV(t, j-1) is accumulated (PV(t, j) is not issued before p_ready(j)), then scale this row of O_t in place
        mbar_wait(&pv_done[4 * t + ((j - 1) & 3)], (uint32_t)((j - 1) >> 2) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_addr + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          tmem_st_32x32b_x32(o_addr + ch * 32, o);
        }
        tmem_st_wait();
      }
      if (!P_TMEM) fence_proxy_async();   // generic-proxy smem writes of P → visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[2 * t + buf]);     // one arrival per warp (4 per tile): per-thread arrivals serialise on the smem atomic unit
    }
    if (nb > 0) {
      mbar_wait(&pv_done[4 * t + ((nb - 1) & 3)], (uint32_t)((nb - 1) >> 2) & 1u);
      tc_fence_after();
      const float inv = l > 0.f ? 1.f / l : 0.f;
      const bool valid = q_idx < p.sq;
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)q_idx * p.o_pitch + ((size_t)bi * p.hq + h) * D;
#pragma unroll 1
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(o_addr + ch * 32, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(o[g * 16 + 2 * e]) * inv, __uint_as_float(o[g * 16 + 2 * e + 1]) * inv);
              v[e] = *reinterpret_cast<uint32_t*>(&hb);
            }
            st_global_v8(orow + ch * 32 + g * 16, v);
          }
        }
      }
      if (valid && p.lse != nullptr) p.lse[((size_t)bi * p.hq + h) * p.sq + q_idx] = m_ref * 0.6931471805599453f + logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int D, bool P_TMEM>
static int launch_fa_fwd(const void* q, const void* k, const void* v, FaParams p, long q_ss, long k_ss, long v_ss, cudaStream_t s) {
  constexpr int SMEM_BYTES = 2 * (FA_BM * D * 2) + 2 * FA_STAGES * (FA_BN * D * 2) + 4 * (FA_BM * 128) + 1024 + 256;
  CUtensorMap tq, tk, tv;
  bool ok = make_tmap_bf16_strided(&tq, q, p.sq, q_ss, q_ss * 2, 64, FA_BM);
  ok &= make_tmap_bf16_strided(&tk, k, p.sk, k_ss, k_ss * 2, 64, FA_BN);
  ok &= make_tmap_bf16_strided(&tv, v, p.sk, v_ss, v_ss * 2, 64, FA_BN);
  if (!ok) return -1;
  auto kern = fa_fwd_kernel<D, P_TMEM>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  dim3 grid((p.sq + 2 * FA_BM - 1) / (2 * FA_BM), p.hq, p.b);
  // mirrored pairing when the whole grid is at most ~1.5 waves and the mask is causal (otherwise every tile costs the same)
  p.pair_mode = (p.causal && p.row_lo == nullptr && p.sq == p.sk && (long)grid.x * grid.y * grid.z <= (long)num_sms() * 3 / 2 && grid.x > 1) ? 1 : 0;
  if (const char* e = getenv("MB200_FA_PAIR_MODE")) p.pair_mode = p.row_lo == nullptr ? atoi(e) : 0;
  kern<<<grid, FA_THREADS, SMEM_BYTES, s>>>(tq, tk, tv, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace mb200

using namespace mb200;

// Strides are in elements; *_ss = sequence stride (row pitch), *_sb = batch stride, *_sh = head stride; d is contiguous.
extern "C" int mb200_flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int sq, int sk, int b, int hq, int hk, int d, long q_ss,
                                    long q_sb, long q_sh, long k_ss, long k_sb, long k_sh, long v_ss, long v_sb, long v_sh, float scale, int causal,
                                    int variant, const int* row_lo, cudaStream_t s) {
  if ((d != 64 && d != 128) || hq % hk != 0) return -10;
  if ((q_ss | q_sb | q_sh | k_ss | k_sb | k_sh | v_ss | v_sb | v_sh) % 8 != 0) return -11;   // 16-byte alignment for TMA
  FaParams p;
  p.sq = sq; p.sk = sk; p.b = b; p.hq = hq; p.hk = hk; p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q_sb = q_sb; p.q_sh = q_sh; p.k_sb = k_sb; p.k_sh = k_sh; p.v_sb = v_sb; p.v_sh = v_sh;
  p.out = out; p.o_pitch = (long)b * hq * d; p.lse = lse; p.row_lo = row_lo;
  if (variant == 1) return d == 128 ? launch_fa_fwd<128, true>(q, k, v, p, q_ss, k_ss, v_ss, s) : launch_fa_fwd<64, true>(q, k, v, p, q_ss, k_ss, v_ss, s);
  return d == 128 ? launch_fa_fwd<128, false>(q, k, v, p, q_ss, k_ss, v_ss, s) : launch_fa_fwd<64, false>(q, k, v, p, q_ss, k_ss, v_ss, s);
}
