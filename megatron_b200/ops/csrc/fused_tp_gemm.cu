// Tensor-parallel GEMMs FUSED with their sequence-parallel collective over NVLink / NVSwitch (sm_100a).
//
// One persistent kernel = tcgen05 2-CTA GEMM clusters + a few "comm" clusters that move data between GPUs
// through peer-mapped symmetric memory while the tensor cores run.  No NCCL call, no separate collective
// launch, no grid-wide barrier: ordering is carried by per-chunk flags (st.release.sys / ld.acquire.sys).
//
//  mode AG  (all-gather → GEMM;  ColumnParallelLinear fwd, RowParallelLinear dgrad)
//     comm clusters : push this rank's activation shard, 256-row chunk by chunk, into EVERY rank's gathered
//                     buffer with `multimem.st` (NVSwitch replicates), then publish flag[rank][chunk] = epoch.
//     GEMM clusters : walk M-tiles in arrival order (own shard first, then chunk 0 of every peer, chunk 1, …);
//                     the TMA producer spins on the chunk flag before loading A tiles of that chunk.
//  mode RS  (GEMM → reduce-scatter;  RowParallelLinear fwd, ColumnParallelLinear dgrad)
//     GEMM clusters : walk M-tiles destination shard by destination shard (same order on all ranks); the
//                     epilogue writes the partial tile to the local symmetric buffer Y; when a 256-row block of
//                     Y is complete the last epilogue warp publishes flag_on_owner[chunk][rank] = epoch.
//     comm clusters : for each chunk of MY shard wait for all ranks' flags, then `multimem.ld_reduce` the block
//                     (fp32 accumulation in the switch) and store the reduced rows to the local output.
//
// Buffer reuse needs no extra barrier: buffers are double-buffered by the host and every op needs data from
// every peer, so a rank can only be two ops ahead of a peer that has finished reading the buffer being reused.
// Replaces TE userbuffers `ub_overlap_ag/rs` and the NCCL calls in tensor_parallel/layers.py.
#include <cstdio>
#include <mutex>

#include "gemm_sm100_device.cuh"

namespace mb200 {
using namespace ptx;

constexpr int MAX_TP = 8;
constexpr int CHUNK_ROWS = 256;   // = pair-tile rows
constexpr int MAX_CHUNKS = 64;    // chunks per rank shard
constexpr int AG_OFF = 0, RS_OFF = MAX_TP * MAX_CHUNKS, XAG_OFF = 2 * MAX_TP * MAX_CHUNKS;
constexpr int XAG_COUNTER = MAX_TP * MAX_CHUNKS;  // index into the local counters array
constexpr int AR_COUNTER = XAG_COUNTER + 8;        // + chunk: comm-warp counters of an all-reduce launch (counters array holds 1024 uint32)

struct FusedParams {
  GemmParams g;
  int mode;                 // 0 = AG, 1 = RS, 2 = AR (GEMM -> all-reduce, result in place in the symmetric Y on every rank)
  int rank, world;
  int chunks_per_rank;      // rows_per_rank / 256
  int comm_clusters;        // clusters [0, comm_clusters) run the collective (scheduled first), the rest run the GEMM
  uint32_t epoch;
  // AG
  const void* ag_src;       // local shard [rows_per_rank, K] bf16
  void* ag_dst_mc;          // multicast address of the gathered buffer (nullptr → P2P stores)
  void* ag_dst_peer[MAX_TP];
  // RS
  const void* rs_src_mc;    // multicast address of the partial-sum buffer Y (nullptr → P2P loads)
  const void* rs_src_peer[MAX_TP];
  void* rs_out;             // local output [rows_per_rank, N] bf16
  // piggy-back all-gather done by the comm clusters of an RS-mode launch (the wgrad operand of the same layer)
  const void* xag_src;      // local shard, xag_vec 16-byte vectors (0 → none)
  size_t xag_vec;
  void* xag_dst_mc;
  void* xag_dst_peer[MAX_TP];
  // flags (symmetric, uint32): AG  flags[AG_OFF + rank_of_shard*MAX_CHUNKS + chunk];  RS  flags[RS_OFF + chunk*MAX_TP + src_rank];
  // piggy-back AG  flags[XAG_OFF + src_rank]
  uint32_t* flags_peer[MAX_TP];
  uint32_t* counters;       // local: per M-tile completion counters (RS) / per chunk comm-CTA counters (AG)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Spin until *f has reached `epoch` (wrap-safe signed compare).  A peer that died or a protocol bug must not hang the box: after
// SPIN_TIMEOUT_NS the kernel reports which flag it was waiting for and traps (the host sees a launch failure, not a deadlock).
constexpr uint64_t SPIN_TIMEOUT_NS = 20ull * 1000000000ull;
__device__ __forceinline__ void spin_until(const uint32_t* f, uint32_t epoch, int what) {
  if ((int32_t)(ld_acquire_sys_u32(f) - epoch) >= 0) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t it = 0;
  while ((int32_t)(ld_acquire_sys_u32(f) - epoch) < 0) {
    if ((++it & 0x3FFu) == 0 && globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) {
      printf("[mb200 fused_tp_gemm] flag wait timed out: kind=%d flag=%p have=%u want=%u block=%d\n", what, (const void*)f, ld_acquire_sys_u32(f), epoch, (int)blockIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mm_st_v4(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 mm_ld_reduce_bf16x8(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// NVLink round trips are microseconds: keep U independent 16-byte transfers in flight per thread (Little's law), i.e. issue all
// loads of a batch before the first dependent store.  n = number of 16-byte vectors, tid/nthr = this thread's slot in the comm grid.
template <int U>
__device__ __forceinline__ void push_multicast(uint4* __restrict__ mc_dst, const uint4* __restrict__ src, size_t n, size_t tid, size_t nthr) {
  size_t i = tid;
  for (; i + (size_t)(U - 1) * nthr < n; i += (size_t)U * nthr) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src + i + (size_t)u * nthr);
#pragma unroll
    for (int u = 0; u < U; ++u) mm_st_v4(mc_dst + i + (size_t)u * nthr, v[u]);
  }
  for (; i < n; i += nthr) mm_st_v4(mc_dst + i, __ldg(src + i));
}
template <int U>
__device__ __forceinline__ void pull_reduce_multicast(uint4* __restrict__ out, const uint4* __restrict__ mc_src, size_t n, size_t tid, size_t nthr) {
  size_t i = tid;
  for (; i + (size_t)(U - 1) * nthr < n; i += (size_t)U * nthr) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = mm_ld_reduce_bf16x8(mc_src + i + (size_t)u * nthr);
#pragma unroll
    for (int u = 0; u < U; ++u) out[i + (size_t)u * nthr] = v[u];
  }
  for (; i < n; i += nthr) out[i] = mm_ld_reduce_bf16x8(mc_src + i);
}

template <int U>
__device__ __forceinline__ void reduce_broadcast_multicast(uint4* __restrict__ mc, size_t n, size_t tid, size_t nthr) {
  size_t i = tid;
  for (; i + (size_t)(U - 1) * nthr < n; i += (size_t)U * nthr) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = mm_ld_reduce_bf16x8(mc + i + (size_t)u * nthr);
#pragma unroll
    for (int u = 0; u < U; ++u) mm_st_v4(mc + i + (size_t)u * nthr, v[u]);
  }
  for (; i < n; i += nthr) mm_st_v4(mc + i, mm_ld_reduce_bf16x8(mc + i));
}

// position in the tile walk → M pair-tile index
__device__ __forceinline__ int walk_to_mblk(const FusedParams& p, int pos) {
  const int C = p.chunks_per_rank, W = p.world;
  if (p.mode == 0) {
    // AG: own shard first, then chunk c of every peer (in ring order) for c = 0..C-1  == arrival order
    if (pos < C) return p.rank * C + pos;
    const int q = pos - C;
    const int c = q / (W - 1), dr = q % (W - 1) + 1;
    return ((p.rank + dr) % W) * C + c;
  }
  // RS: destination shard 0 first on every rank, so owners can start reducing early
  return pos;
}

template <bool B_MN, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
fused_tp_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ Cptr, const __grid_constant__ FusedParams p) {
  constexpr int BNH = BN / 2;
  constexpr int B_STAGE_BYTES = BNH * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int PM = 2 * BM;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const int cluster_id = ((int)blockIdx.x >> 1) - p.comm_clusters;  // < 0: communication cluster
  const GemmParams& g = p.g;
  const int tiles_m = (g.M + PM - 1) / PM, tiles_n = (g.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;

  if (cluster_id < 0) {
    // =============================== communication clusters =================================================
    const int cta = (int)blockIdx.x;   // comm CTA index
    const int nctas = p.comm_clusters * 2;
    const int C = p.chunks_per_rank;
    if (p.mode == 0) {
      // ---- AG: push my shard chunk by chunk -------------------------------------------------------------
      const size_t row_bytes = (size_t)g.K * 2;
      const size_t chunk_vec = (size_t)CHUNK_ROWS * row_bytes / 16;
      // Work item = (chunk c, slice s): 8 rows of a 256-row chunk.  Every WARP walks its own items and fences its own stores, so
      // the (multi-microsecond) system fences of different warps overlap with other warps' traffic; the warp that completes a
      // chunk (32 slices) publishes the chunk flag on every rank.
      constexpr int SLICES = 32;
      const size_t slice_vec = chunk_vec / SLICES;
      const int n_warps = nctas * (NUM_THREADS / 32), gwarp = cta * (NUM_THREADS / 32) + warp;
      for (int item = gwarp; item < C * SLICES; item += n_warps) {
        const int c = item / SLICES, sl = item % SLICES;
        const size_t off = (size_t)c * chunk_vec + (size_t)sl * slice_vec;
        const uint4* src = reinterpret_cast<const uint4*>(p.ag_src) + off;
        const size_t dst_off = (size_t)p.rank * C * chunk_vec + off;
        if (p.ag_dst_mc != nullptr) {
          push_multicast<8>(reinterpret_cast<uint4*>(p.ag_dst_mc) + dst_off, src, slice_vec, (size_t)lane, 32);
        } else {
          for (size_t i = lane; i < slice_vec; i += 32) {
            const uint4 v = src[i];
            for (int d = 0; d < p.world; ++d) reinterpret_cast<uint4*>(p.ag_dst_peer[(p.rank + d) % p.world])[dst_off + i] = v;
          }
        }
        __threadfence_system();
        __syncwarp();
        uint32_t last = 0;
        if (lane == 0) last = (atomicAdd(&p.counters[c], 1u) == (uint32_t)SLICES - 1) ? 1u : 0u;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          if (lane == 0) p.counters[c] = 0;
          __threadfence();
          if (lane < p.world) st_release_sys_u32(p.flags_peer[lane] + AG_OFF + p.rank * MAX_CHUNKS + c, p.epoch);
        }
      }
    } else {
      // ---- RS: reduce the chunks of MY shard as they complete on all ranks ---------------------------------
      const size_t row_vec = (size_t)g.N * 2 / 16;
      const size_t chunk_vec = (size_t)CHUNK_ROWS * row_vec;
      if (p.xag_vec != 0) {
        // piggy-back all-gather (the wgrad operand) while the first output chunks are still being computed; warp-granular items as in AG mode
        constexpr size_t XSLICE = 4096;   // 64 KiB
        const int n_items = (int)((p.xag_vec + XSLICE - 1) / XSLICE);
        const int n_warps = nctas * (NUM_THREADS / 32), gwarp = cta * (NUM_THREADS / 32) + warp;
        const uint4* src0 = reinterpret_cast<const uint4*>(p.xag_src);
        const size_t base = (size_t)p.rank * p.xag_vec;
        for (int item = gwarp; item < n_items; item += n_warps) {
          const size_t off = (size_t)item * XSLICE;
          const size_t n = min(XSLICE, p.xag_vec - off);
          if (p.xag_dst_mc != nullptr) {
            push_multicast<8>(reinterpret_cast<uint4*>(p.xag_dst_mc) + base + off, src0 + off, n, (size_t)lane, 32);
          } else {
            for (size_t i = lane; i < n; i += 32) {
              const uint4 v = src0[off + i];
              for (int d = 0; d < p.world; ++d) reinterpret_cast<uint4*>(p.xag_dst_peer[(p.rank + d) % p.world])[base + off + i] = v;
            }
          }
          __threadfence_system();
          __syncwarp();
          uint32_t last = 0;
          if (lane == 0) last = (atomicAdd(&p.counters[XAG_COUNTER], 1u) == (uint32_t)n_items - 1) ? 1u : 0u;
          last = __shfl_sync(0xffffffffu, last, 0);
          if (last) {
            if (lane == 0) p.counters[XAG_COUNTER] = 0;
            __threadfence();
            if (lane < p.world) st_release_sys_u32(p.flags_peer[lane] + XAG_OFF + p.rank, p.epoch);
          }
        }
      }
      // Work item = (chunk c of MY shard, slice of 8 rows).  Warps walk their own items in chunk order: a warp waits only for the chunk
      // its next slice belongs to, and ~200 warps keep ~1.5 MB of in-switch reductions in flight (the NVLink round trip is microseconds).
      {
        constexpr int SLICES = 32;
        const size_t slice_vec = chunk_vec / SLICES;
        const int n_warps = nctas * (NUM_THREADS / 32), gwarp = cta * (NUM_THREADS / 32) + warp;
        for (int item = gwarp; item < C * SLICES; item += n_warps) {
          const int c = item / SLICES, sl = item % SLICES;
          if (lane < p.world) {
            spin_until(p.flags_peer[p.rank] + RS_OFF + c * MAX_TP + lane, p.epoch, 1);
          }
          __syncwarp();
          const size_t src_off = ((size_t)p.rank * C + c) * chunk_vec + (size_t)sl * slice_vec;
          uint4* out = reinterpret_cast<uint4*>(p.rs_out) + (size_t)c * chunk_vec + (size_t)sl * slice_vec;
          if (p.mode == 2) {
            // all-reduce: the reduced rows go straight back into EVERY rank's Y (in place: only this warp reads these elements, and each
            // element's broadcast store depends on its own reduction load), then the chunk is announced like an all-gather chunk
            if (p.rs_src_mc != nullptr) {
              reduce_broadcast_multicast<16>(reinterpret_cast<uint4*>(const_cast<void*>(p.rs_src_mc)) + src_off, slice_vec, (size_t)lane, 32);
            } else {
              for (size_t i = lane; i < slice_vec; i += 32) {
                float acc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = 0.f;
                for (int d = 0; d < p.world; ++d) {
                  const uint4 v = reinterpret_cast<const uint4*>(p.rs_src_peer[(p.rank + d) % p.world])[src_off + i];
                  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
                  for (int k = 0; k < 8; ++k) acc[k] += __bfloat162float(h[k]);
                }
                uint4 o;
                __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int k = 0; k < 4; ++k) ob[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
                for (int d = 0; d < p.world; ++d) reinterpret_cast<uint4*>(const_cast<void*>(p.rs_src_peer[(p.rank + d) % p.world]))[src_off + i] = o;
              }
            }
            __threadfence_system();
            __syncwarp();
            uint32_t last = 0;
            if (lane == 0) last = (atomicAdd(&p.counters[AR_COUNTER + c], 1u) == (uint32_t)SLICES - 1) ? 1u : 0u;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
              if (lane == 0) p.counters[AR_COUNTER + c] = 0;
              __threadfence();
              if (lane < p.world) st_release_sys_u32(p.flags_peer[lane] + AG_OFF + p.rank * MAX_CHUNKS + c, p.epoch);
            }
            continue;
          }
          if (p.rs_src_mc != nullptr) {
            pull_reduce_multicast<16>(out, reinterpret_cast<const uint4*>(p.rs_src_mc) + src_off, slice_vec, (size_t)lane, 32);
          } else {
            for (size_t i = lane; i < slice_vec; i += 32) {
              float acc[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] = 0.f;
              for (int d = 0; d < p.world; ++d) {
                const uint4 v = reinterpret_cast<const uint4*>(p.rs_src_peer[(p.rank + d) % p.world])[src_off + i];
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += __bfloat162float(h[k]);
              }
              uint4 o;
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int k = 0; k < 4; ++k) ob[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
              out[i] = o;
            }
          }
        }
      }
      if (p.mode == 2) {
        // the kernel may not complete before every owner has written its reduced chunks into my Y: comm warps share the (rank, chunk) flags
        const int n_warps = nctas * (NUM_THREADS / 32), gwarp = cta * (NUM_THREADS / 32) + warp;
        for (int item = gwarp * 32 + lane; item < p.world * C; item += n_warps * 32) {
          const int r = item / C, c = item % C;
          spin_until(p.flags_peer[p.rank] + AG_OFF + r * MAX_CHUNKS + c, p.epoch, 3);
        }
      }
      if (p.xag_vec != 0 && cta == 0 && threadIdx.x < 32) {
        // the kernel may not complete before every peer's piggy-back shard has landed here
        if (lane < p.world) {
          spin_until(p.flags_peer[p.rank] + XAG_OFF + lane, p.epoch, 2);
        }
      }
    }
    // comm clusters take no part in the GEMM clusters' TMEM / cluster barriers: each CTA pair is its own cluster
    return;
  }

  // ===================================== GEMM clusters (2-CTA tcgen05) ==========================================
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const bool leader = cta_rank == 0;
  const int num_clusters = ((int)gridDim.x >> 1) - p.comm_clusters;
  const int k_blocks = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_holder);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_pos, n_blk;
        tile_coords(tile, tiles_m, tiles_n, g.group_m, m_pos, n_blk);   // groups of walk positions sweep N together: weight tiles stay in L2
        const int m_blk = walk_to_mblk(p, m_pos);
        if (p.mode == 0) {
          // wait until the 256-row chunk holding this tile's A rows has landed in the local gathered buffer
          const int r = m_blk / p.chunks_per_rank, c = m_blk % p.chunks_per_rank;
          spin_until(p.flags_peer[p.rank] + AG_OFF + r * MAX_CHUNKS + c, p.epoch, 0);
          fence_proxy_async_global();  // peer (generic-proxy) writes → our TMA (async-proxy) reads
        }
        const int m0 = m_blk * PM + (int)cta_rank * BM;
        const int n0 = n_blk * BN + (int)cta_rank * BNH;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES); else mbar_arrive_remote(&full_bar[stage], 0);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
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
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PM, BN, false, B_MN);
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
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, BK * 128, 1024) : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16_2sm(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
    constexpr int CH = BN / 32 / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_pos, n_blk;
      tile_coords(tile, tiles_m, tiles_n, g.group_m, m_pos, n_blk);
      const int m_blk = walk_to_mblk(p, m_pos);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<false>(Cptr, g, tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16), m_blk * PM + (int)cta_rank * BM + ew * 32 + lane, n_blk * BN, half * CH,
                           (half + 1) * CH, lane, &tmem_empty[acc], !leader);
      if (p.mode >= 1) {
        // publish: this warp's part of tile (m_blk, n_blk) is in Y.  The last of tiles_n * 16 warp-parts of the
        // 256-row block tells the owner rank that its chunk is complete on this rank.
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          const uint32_t total = (uint32_t)tiles_n * 2u * EPI_WARPS;
          if (atomicAdd(&p.counters[m_blk], 1u) == total - 1) {
            p.counters[m_blk] = 0;
            __threadfence_system();
            const int owner = m_blk / p.chunks_per_rank, c = m_blk % p.chunks_per_rank;
            st_release_sys_u32(p.flags_peer[owner] + RS_OFF + c * MAX_TP + p.rank, p.epoch);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

template <bool B_MN, int BN>
static int launch_fused(const void* A, const void* B, void* C, FusedParams p, int comm_clusters, cudaStream_t s) {
  constexpr int BNH = BN / 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BNH * BK * 2;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ta, tb;
  bool ok = make_tmap_bf16(&ta, A, p.g.M, p.g.K, BK, BM);
  ok &= B_MN ? make_tmap_bf16(&tb, B, p.g.K, p.g.N, 64, BK) : make_tmap_bf16(&tb, B, p.g.N, p.g.K, BK, BNH);
  if (!ok) return -1;
  auto kern = fused_tp_gemm_kernel<B_MN, BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -3;
    configured = true;
  }
  const int total_clusters = num_sms() / 2;
  const int tiles = ((p.g.M + 2 * BM - 1) / (2 * BM)) * ((p.g.N + BN - 1) / BN);
  int gemm_clusters = total_clusters - comm_clusters;
  if (gemm_clusters > tiles) gemm_clusters = tiles;
  p.comm_clusters = comm_clusters;
  kern<<<(gemm_clusters + comm_clusters) * 2, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, C, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace mb200

using namespace mb200;

// mode 0 (AG): A = local gathered buffer [M, K] (this kernel fills it), ag_src = local shard, C = output [M, N] bf16.
// mode 1 (RS): A = local input [M, K], C = Y partial buffer [M, N] bf16 in symmetric memory, rs_out = [M/world, N].
// b_layout: 0 → B[N, K] (K-major), 1 → B[K, N] (N-major).  Flags arrays are uint32 in symmetric memory.
extern "C" int mb200_fused_tp_gemm(int mode, const void* A, const void* B, void* C, int M, int N, int K, int b_layout, int rank, int world, uint32_t epoch,
                                   const void* ag_src, int64_t ag_dst_mc, const int64_t* ag_dst_peer, int64_t rs_src_mc, const int64_t* rs_src_peer, void* rs_out,
                                   const void* xag_src, int64_t xag_bytes, int64_t xag_dst_mc, const int64_t* xag_dst_peer, const int64_t* flags_peer, void* counters,
                                   int comm_clusters, cudaStream_t s) {
  if (world > MAX_TP || M % (world * CHUNK_ROWS) != 0 || (M / world / CHUNK_ROWS) > MAX_CHUNKS) return -10;
  if (K % 8 != 0 || N % 8 != 0) return -11;
  FusedParams p;
  p.g.M = M; p.g.N = N; p.g.K = K; p.g.ldc = N; p.g.accumulate = 0;
  p.g.group_m = 8 < M / (2 * BM) ? 8 : (M / (2 * BM) > 0 ? M / (2 * BM) : 1);
  p.mode = mode; p.rank = rank; p.world = world; p.chunks_per_rank = M / world / CHUNK_ROWS; p.epoch = epoch; p.comm_clusters = 0;
  p.ag_src = ag_src; p.ag_dst_mc = reinterpret_cast<void*>(ag_dst_mc); p.rs_src_mc = reinterpret_cast<const void*>(rs_src_mc); p.rs_out = rs_out;
  p.xag_src = xag_src; p.xag_vec = xag_src ? (size_t)xag_bytes / 16 : 0; p.xag_dst_mc = reinterpret_cast<void*>(xag_dst_mc);
  if (xag_bytes % 16 != 0) return -12;
  for (int i = 0; i < MAX_TP; ++i) {
    p.xag_dst_peer[i] = (i < world && xag_dst_peer) ? reinterpret_cast<void*>(xag_dst_peer[i]) : nullptr;
    p.ag_dst_peer[i] = (i < world && ag_dst_peer) ? reinterpret_cast<void*>(ag_dst_peer[i]) : nullptr;
    p.rs_src_peer[i] = (i < world && rs_src_peer) ? reinterpret_cast<const void*>(rs_src_peer[i]) : nullptr;
    p.flags_peer[i] = i < world ? reinterpret_cast<uint32_t*>(flags_peer[i]) : nullptr;
  }
  p.counters = reinterpret_cast<uint32_t*>(counters);
  if (comm_clusters < 1) comm_clusters = 1;
  const bool small_n = N <= 128;
  if (b_layout == 0) return small_n ? launch_fused<false, 128>(A, B, C, p, comm_clusters, s) : launch_fused<false, 256>(A, B, C, p, comm_clusters, s);
  return small_n ? launch_fused<true, 128>(A, B, C, p, comm_clusters, s) : launch_fused<true, 256>(A, B, C, p, comm_clusters, s);
}
