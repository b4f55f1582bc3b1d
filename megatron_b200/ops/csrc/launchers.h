// C launcher declarations shared by the kernel translation units and bindings.cpp (host-only safe).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb200 {
enum DType : int { kF32 = 0, kBF16 = 1, kF16 = 2 };
}

// ---- C launchers (implemented in the .cu files, called from bindings.cpp) ---------------
extern "C" {
void mb200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int H, float eps, int zero_centered, int dtype, cudaStream_t s);
void mb200_rmsnorm_bwd(const void* gy, const void* x, const void* w, const float* rstd, void* gx, float* gw_partial, void* gw, int rows, int H,
                       int zero_centered, int dtype, int nblocks, cudaStream_t s);
void mb200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mu, float* rstd, int rows, int H, float eps, int zero_centered,
                         int dtype, cudaStream_t s);
void mb200_layernorm_bwd(const void* gy, const void* x, const void* w, const float* mu, const float* rstd, void* gx, float* partial, void* gw, void* gb,
                         int rows, int H, int zero_centered, int dtype, int nblocks, cudaStream_t s);
void mb200_swiglu_fwd(const void* y, const void* bias, const float* probs, void* out, long rows, int F, int dtype, cudaStream_t s);
void mb200_swiglu_bwd(const void* g, const void* y, const void* bias, const float* probs, void* dy, float* dprobs, long rows, int F, int dtype, cudaStream_t s);
void mb200_rope(const void* t, const float* freqs, void* out, int S, int B, int Hh, int D, int Drot, float mscale, int conj, int dtype, cudaStream_t s);
void mb200_ce_stats(const void* logits, const long* target, float* stats, int rows, int V, long vocab_start, int dtype, cudaStream_t s);
void mb200_ce_bwd(void* logits, const long* target, const float* lse, const float* gloss, int rows, int V, long vocab_start, int dtype, cudaStream_t s);
void mb200_multi_l2norm(const void* const* ptrs, const long* sizes, const int* dtypes, int n, float* partial, float* out, int nblocks, cudaStream_t s);
void mb200_multi_scale(void* const* ptrs, const long* sizes, const int* dtypes, int n, const float* scale, int nblocks, cudaStream_t s);
void mb200_multi_adam(float* const* p32, const void* const* grads, float* const* m, float* const* v, void* const* lowp, const long* sizes,
                      const int* gdtypes, const int* ldtypes, int n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, int adamw,
                      const float* grad_scale, int nblocks, cudaStream_t s);
void mb200_nvl_barrier(const int64_t* ptrs, const int64_t* flags, int rank, int world, uint32_t epoch, int slot, cudaStream_t s);
void mb200_nvl_allgather(const int64_t* ptrs, const int64_t* flags, int64_t mc, const void* src, size_t dst_off, size_t shard_bytes, int rank, int world,
                         uint32_t epoch, void* ctrl, int slot, int nblocks, cudaStream_t s);
void mb200_nvl_reducescatter(const int64_t* ptrs, const int64_t* flags, int64_t mc, size_t src_off, void* out, size_t shard_elems, float scale, int dtype,
                             int rank, int world, uint32_t epoch, void* ctrl, int slot, int trailing, int nblocks, cudaStream_t s);
void mb200_nvl_allreduce(const int64_t* ptrs, const int64_t* flags, int64_t mc, size_t off, size_t elems, float scale, int dtype, int rank, int world,
                         uint32_t epoch, void* ctrl, int slot, int nblocks, cudaStream_t s);
int mb200_gemm_bf16_v(const void* A, const void* B, void* C, int M, int N, int K, int layout, int accumulate, int c_dtype, int variant, cudaStream_t s);
int mb200_fused_tp_gemm(int mode, const void* A, const void* B, void* C, int M, int N, int K, int b_layout, int rank, int world, uint32_t epoch,
                        const void* ag_src, int64_t ag_dst_mc, const int64_t* ag_dst_peer, int64_t rs_src_mc, const int64_t* rs_src_peer, void* rs_out,
                        const void* xag_src, int64_t xag_bytes, int64_t xag_dst_mc, const int64_t* xag_dst_peer, const int64_t* flags_peer, void* counters,
                        int comm_clusters, cudaStream_t s);
int mb200_flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int sq, int sk, int b, int hq, int hk, int d, long q_ss, long q_sb,
                         long q_sh, long k_ss, long k_sb, long k_sh, long v_ss, long v_sb, long v_sh, float scale, int causal, int variant, const int* row_lo, cudaStream_t s);
void mb200_batched_copy(const void* tasks_dev, const void* chunk_prefix_dev, int ntasks, unsigned long long total_chunks, int nblocks, cudaStream_t s);
int mb200_grouped_gemm_bf16(const void* a, const void* b, void* c, const int* offsets, int E, int dim_n, int dim_k, int mode, int accumulate, int c_dtype,
                            void* maps_dev, cudaStream_t s);
int mb200_gemm_fp8_nt(const void* A, const void* B, void* C, int M, int N, int K, int a_fmt, int b_fmt, float alpha, const float* alpha_dev, cudaStream_t s);
void mb200_moe_gather_rows(const void* in, void* out, const int64_t* src, const float* scale, int64_t n_out, int hidden, cudaStream_t s);
void mb200_moe_combine_rows(const void* in, void* out, const int64_t* pos, const float* w, int64_t n_tokens, int topk, int hidden, cudaStream_t s);
int mb200_moe_topk_router(const float* logits, const float* expert_bias, int T, int E, int topk, int score_fn, int renormalize, float scaling, float* probs, int64_t* ids,
                          uint8_t* routing_map, int* tokens_per_expert, cudaStream_t s);
void mb200_moe_push_rows(const void* in, const int64_t* src_row, const int32_t* dst_rank, const int64_t* dst_slot, const int64_t* peer_ptrs, int world, size_t dst_off_bytes,
                         int64_t n_pairs, int row_bytes, cudaStream_t s);
void mb200_moe_pull_rows(void* out, const int32_t* rank, const int64_t* slot, const float* w, const int64_t* peer_ptrs, int world, size_t src_off_bytes, int64_t n_out,
                         int topk, int row_bytes, int raw, cudaStream_t s);
int mb200_flash_attn_bwd(const void* q, const void* k, const void* v, const void* dout, const void* out, const float* lse, float* vec, void* dq, void* dk, void* dv,
                         void* scratch, int split_heads, int sq, int sk, int b, int hq, int hk, int d, long q_ss, long q_sb, long q_sh, long k_ss, long k_sb, long k_sh,
                         long v_ss, long v_sb, long v_sh, long do_ss, long do_sb, long do_sh, long o_ss, long o_sb, long o_sh, float scale, int causal, const int* row_lo,
                         const int* col_hi, cudaStream_t s);
void mb200_indices_to_multihot(const int64_t* idx, const float* probs, uint8_t* map, float* probs_out, long T, int k, int E, cudaStream_t s);
void mb200_multihot_probs_grad(const int64_t* idx, const float* g_in, float* g_out, long T, int k, int E, int scatter, cudaStream_t s);
void mb200_multihot_to_indices(const uint8_t* map, const float* probs, int64_t* idx, float* probs_out, long T, int k, int E, cudaStream_t s);
void mb200_pad_routing_map(const uint8_t* in, uint8_t* out, long T, int E, int multiple, cudaStream_t s);
void mb200_moe_aux_loss_fwd(const float* probs, const float* tpe, float* partial, int nblocks, float* loss, long T, int E, float coeff, cudaStream_t s);
void mb200_moe_aux_loss_bwd(const float* tpe, const float* gloss, float coeff, float* gprobs, long T, int E, cudaStream_t s);
int mb200_mla_rope_inplace(const void* src, void* x, const float* ang, const int64_t* pos, long rows, int H, int nope, int emb, int batch, float mscale, int interleaved, int inverse, int dtype,
                           cudaStream_t s);
int mb200_mla_kv_split(const void* a, const void* b, const float* ang, const int64_t* pos, void* o0, void* o1, long rows, int H, int kd, int vd, int emb, int batch, float mscale,
                       int interleaved, int backward, int dtype, cudaStream_t s);
int mb200_paged_stash(void* flat, void* pages, const int64_t* page_ids, const int64_t* num_tokens, long t_max, long row_bytes, int page_size, int pop, cudaStream_t s);
void mb200_spec_verify(const int64_t* draft_tokens, const float* draft_probs, const float* target_probs, const float* u_accept, const float* u_sample, int64_t* n_accepted,
                       int64_t* next_token, int B, int k, int V, cudaStream_t s);
long mb200_bias_dropout_add_draws(long numel, int dtype, int* grid_out);
int mb200_bias_dropout_add(const void* x, const void* bias, const void* residual, void* y, long numel, int H, float p, unsigned long long seed, unsigned long long offset,
                           int backward, int dtype, cudaStream_t s);
size_t mb200_flash_attn_bwd_scratch_bytes(int sq, int sk, int b, int hq, int hk, int split_heads);
int mb200_flash_attn_bwd_split_heads(int sk, int b, int hq, int hk);
void mb200_paged_kv_append(const void* k_new, const void* v_new, void* k_pool, void* v_pool, const int32_t* block_table, const int32_t* positions, int B, int table_width,
                           int block_size, int hk, int d, cudaStream_t s);
int mb200_paged_decode(const void* q, const void* k_pool, const void* v_pool, const int32_t* block_table, const int32_t* lengths, float* o_part, float* ml_part, void* out,
                       int B, int hq, int hk, int d, int table_width, int block_size, float scale, int nsplit, int tokens_per_split, cudaStream_t s);
int mb200_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int layout, int accumulate, int c_dtype, cudaStream_t s);
void mb200_add_rmsnorm_fwd(const void* x, const void* res, const void* w, void* y, void* res_out, float* rstd, int rows, int H, float eps, int zc, int dtype, cudaStream_t s);
void mb200_add_rmsnorm_bwd(const void* gy, const void* gres, const void* h, const void* w, const float* rstd, void* gx, float* partial, void* gw, int rows, int H, int zc,
                           int dtype, int nblocks, cudaStream_t s);
void mb200_rope_pos(const void* t, const float* freqs, const int* pos, void* out, long tokens, int Hh, int D, int Drot, float mscale, int conj, int dtype, cudaStream_t s);
void mb200_rope_qkv(const void* qkv, const float* freqs, void* out, int S, int B, int NG, int QPG, int D, int Drot, float mscale, int conj, int dtype, cudaStream_t s);
int mb200_conv1d_fwd(const void* x, const void* w, const void* bias, const void* left, void* y, long rows, int d, int l, int K, int act, int dtype, cudaStream_t s);
int mb200_conv1d_bwd(const void* gy, const void* x, const void* w, const void* bias, const void* left, void* gx, void* gleft, float* gw_acc, float* gb_acc, long rows,
                     int d, int l, int K, int act, int dtype, cudaStream_t s);
void mb200_ssd_state_fwd(const float* states, const float* decay, const float* init, float* prev, float* fin, int b, int c, int h, int E, cudaStream_t s);
void mb200_ssd_state_bwd(const float* g_prev, const float* g_fin, const float* prev, const float* decay, float* g_states, float* g_init, float* g_decay, int b, int c,
                         int h, int E, cudaStream_t s);
void mb200_ssd_step(float* state, const void* x, const float* dt, const float* A, const void* B, const void* C, const float* D, void* y, int b, int h, int g, int p, int n,
                    int dtype, cudaStream_t s);
void mb200_mxfp8_quant(const void* x, void* q, void* sf, long rows, int K, cudaStream_t s);
void mb200_mxfp8_dequant(const void* q, const void* sf, void* out, long rows, int K, cudaStream_t s);
void mb200_softmax_fwd(const void* x, const void* mask, void* y, long rows, int h, int sq, int sk, float scale, int causal, int dtype, cudaStream_t s);
void mb200_softmax_bwd(const void* gy, const void* y, void* gx, long rows, int sk, float scale, int dtype, cudaStream_t s);
void mb200_act_fwd(const void* x, void* y, long rows, int F, int mode, int dtype, cudaStream_t s);
void mb200_act_bwd(const void* g, const void* x, void* gx, long rows, int F, int mode, int dtype, cudaStream_t s);
void mb200_nvfp4_quant(const void* x, const float* tscale, void* q, void* sf, long rows, int K, cudaStream_t s);
int mb200_gemm_mxfp8_nt(const void* A, const void* B, const void* sfa, const void* sfb, void* C, int M, int N, int K, int tile, cudaStream_t s);
int mb200_gemm_nvfp4_nt(const void* A, const void* B, const void* sfa, const void* sfb, void* C, int M, int N, int K, float alpha, const float* alpha_dev, cudaStream_t s);
}
