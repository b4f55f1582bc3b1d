// PyTorch bindings for the megatron_b200 sm_100a kernels.
// Kernels live in plain .cu translation units with C launchers (common.cuh); this file only
// validates tensors, allocates outputs and forwards raw pointers + the current stream.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <unordered_map>
#include <vector>

#include <ATen/cuda/CUDAGeneratorImpl.h>
#include <ATen/cuda/CUDAGraphsUtils.cuh>

#include "launchers.h"

namespace {

using at::Tensor;
using c10::optional;

int dtype_code(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return mb200::kF32;
    case at::kBFloat16: return mb200::kBF16;
    case at::kHalf: return mb200::kF16;
    default: TORCH_CHECK(false, "megatron_b200: unsupported dtype ", t.scalar_type());
  }
}
int vec_elems(const Tensor& t) { return t.scalar_type() == at::kFloat ? 4 : 8; }
cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }
void check_cuda_contig(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
void check_aligned16(const Tensor& t, const char* name) {
  TORCH_CHECK((reinterpret_cast<uintptr_t>(t.data_ptr()) & 15) == 0, name, " must be 16-byte aligned");
}
int persistent_blocks() {
  static int n = 0;
  if (n == 0) n = at::cuda::getCurrentDeviceProperties()->multiProcessorCount * 4;
  return n;
}

// ---- norms -------------------------------------------------------------------------------------
std::vector<Tensor> rmsnorm_fwd(const Tensor& x, const Tensor& w, double eps, bool zero_centered) {
  check_cuda_contig(x, "x"); check_cuda_contig(w, "weight");
  TORCH_CHECK(x.dim() == 2 && x.size(1) == w.numel() && x.scalar_type() == w.scalar_type());
  TORCH_CHECK(x.size(1) % vec_elems(x) == 0, "hidden size must be a multiple of ", vec_elems(x));
  check_aligned16(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x);
  auto rstd = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  mb200_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), rstd.data_ptr<float>(), (int)x.size(0), (int)x.size(1), (float)eps, zero_centered,
                    dtype_code(x), cur_stream());
  return {y, rstd};
}

std::vector<Tensor> rmsnorm_bwd(const Tensor& gy, const Tensor& x, const Tensor& w, const Tensor& rstd, bool zero_centered) {
  check_cuda_contig(gy, "gy"); check_cuda_contig(x, "x");
  TORCH_CHECK(gy.sizes() == x.sizes() && gy.scalar_type() == x.scalar_type());
  c10::cuda::CUDAGuard g(x.device());
  const int rows = (int)x.size(0), H = (int)x.size(1);
  const int nblocks = std::min(rows, persistent_blocks());
  auto gx = at::empty_like(x);
  auto gw = at::empty_like(w);
  auto partial = at::empty({nblocks, 2, H}, x.options().dtype(at::kFloat));
  mb200_rmsnorm_bwd(gy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr<float>(), gx.data_ptr(), partial.data_ptr<float>(), gw.data_ptr(), rows, H,
                    zero_centered, dtype_code(x), nblocks, cur_stream());
  return {gx, gw};
}

std::vector<Tensor> layernorm_fwd(const Tensor& x, const Tensor& w, const optional<Tensor>& b, double eps, bool zero_centered) {
  check_cuda_contig(x, "x"); check_cuda_contig(w, "weight");
  TORCH_CHECK(x.dim() == 2 && x.size(1) == w.numel() && x.scalar_type() == w.scalar_type());
  TORCH_CHECK(x.size(1) % vec_elems(x) == 0);
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x);
  auto mu = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  auto rstd = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  mb200_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.has_value() ? b->data_ptr() : nullptr, y.data_ptr(), mu.data_ptr<float>(), rstd.data_ptr<float>(),
                      (int)x.size(0), (int)x.size(1), (float)eps, zero_centered, dtype_code(x), cur_stream());
  return {y, mu, rstd};
}

std::vector<Tensor> layernorm_bwd(const Tensor& gy, const Tensor& x, const Tensor& w, const Tensor& mu, const Tensor& rstd, bool zero_centered) {
  check_cuda_contig(gy, "gy"); check_cuda_contig(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  const int rows = (int)x.size(0), H = (int)x.size(1);
  const int nblocks = std::min(rows, persistent_blocks());
  auto gx = at::empty_like(x);
  auto gw = at::empty_like(w);
  auto gb = at::empty_like(w);
  auto partial = at::empty({nblocks, 2, H}, x.options().dtype(at::kFloat));
  mb200_layernorm_bwd(gy.data_ptr(), x.data_ptr(), w.data_ptr(), mu.data_ptr<float>(), rstd.data_ptr<float>(), gx.data_ptr(), partial.data_ptr<float>(),
                      gw.data_ptr(), gb.data_ptr(), rows, H, zero_centered, dtype_code(x), nblocks, cur_stream());
  return {gx, gw, gb};
}

// ---- activations ---------------------------------------------------------------------------------
Tensor swiglu_fwd(const Tensor& y, const optional<Tensor>& bias, const optional<Tensor>& probs) {
  check_cuda_contig(y, "y");
  TORCH_CHECK(y.dim() == 2 && y.size(1) % 2 == 0);
  const int F = (int)(y.size(1) / 2);
  TORCH_CHECK(F % vec_elems(y) == 0, "ffn size must be a multiple of ", vec_elems(y));
  c10::cuda::CUDAGuard g(y.device());
  auto out = at::empty({y.size(0), F}, y.options());
  Tensor pf;
  if (probs.has_value()) pf = probs->to(at::kFloat).contiguous();
  mb200_swiglu_fwd(y.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, probs.has_value() ? pf.data_ptr<float>() : nullptr, out.data_ptr(),
                   (long)y.size(0), F, dtype_code(y), cur_stream());
  return out;
}

std::vector<optional<Tensor>> swiglu_bwd(const Tensor& gout, const Tensor& y, const optional<Tensor>& bias, const optional<Tensor>& probs) {
  check_cuda_contig(gout, "g"); check_cuda_contig(y, "y");
  const int F = (int)(y.size(1) / 2);
  c10::cuda::CUDAGuard g(y.device());
  auto dy = at::empty_like(y);
  Tensor pf, dprobs;
  if (probs.has_value()) {
    pf = probs->to(at::kFloat).contiguous();
    dprobs = at::empty({y.size(0)}, y.options().dtype(at::kFloat));
  }
  mb200_swiglu_bwd(gout.data_ptr(), y.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, probs.has_value() ? pf.data_ptr<float>() : nullptr,
                   dy.data_ptr(), probs.has_value() ? dprobs.data_ptr<float>() : nullptr, (long)y.size(0), F, dtype_code(y), cur_stream());
  optional<Tensor> dp;
  if (probs.has_value()) dp = dprobs.to(probs->scalar_type());
  return {dy, dp};
}

Tensor rope_fwd(const Tensor& t, const Tensor& freqs, double mscale, bool conj) {
  check_cuda_contig(t, "t"); check_cuda_contig(freqs, "freqs");
  TORCH_CHECK(t.dim() == 4 && freqs.dim() == 2 && freqs.scalar_type() == at::kFloat);
  const int S = (int)t.size(0), B = (int)t.size(1), Hh = (int)t.size(2), D = (int)t.size(3), Drot = (int)freqs.size(1);
  TORCH_CHECK(freqs.size(0) >= S, "rope: freqs has fewer positions than the sequence");
  const int vn = vec_elems(t);
  TORCH_CHECK(Drot <= D && (Drot / 2) % vn == 0 && (D - Drot) % vn == 0, "rope: head dim / rotary dim not vectorisable");
  c10::cuda::CUDAGuard g(t.device());
  auto out = at::empty_like(t);
  mb200_rope(t.data_ptr(), freqs.data_ptr<float>(), out.data_ptr(), S, B, Hh, D, Drot, (float)mscale, conj, dtype_code(t), cur_stream());
  return out;
}

// ---- cross entropy -------------------------------------------------------------------------------
Tensor ce_stats(const Tensor& logits, const Tensor& target, int64_t vocab_start) {
  check_cuda_contig(logits, "logits"); check_cuda_contig(target, "target");
  TORCH_CHECK(logits.dim() == 2 && target.scalar_type() == at::kLong && target.numel() == logits.size(0));
  TORCH_CHECK((logits.size(1) * logits.element_size()) % 16 == 0, "vocab shard row must be 16-byte aligned");
  c10::cuda::CUDAGuard g(logits.device());
  auto stats = at::empty({3, logits.size(0)}, logits.options().dtype(at::kFloat));
  mb200_ce_stats(logits.data_ptr(), target.data_ptr<long>(), stats.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1), vocab_start,
                 dtype_code(logits), cur_stream());
  return stats;
}

Tensor ce_bwd(Tensor logits, const Tensor& target, const Tensor& lse, const Tensor& gloss, int64_t vocab_start) {
  check_cuda_contig(logits, "logits");
  TORCH_CHECK(lse.scalar_type() == at::kFloat && gloss.scalar_type() == at::kFloat && lse.is_contiguous() && gloss.is_contiguous());
  c10::cuda::CUDAGuard g(logits.device());
  mb200_ce_bwd(logits.data_ptr(), target.data_ptr<long>(), lse.data_ptr<float>(), gloss.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1),
               vocab_start, dtype_code(logits), cur_stream());
  return logits;
}

// ---- multi-tensor ----------------------------------------------------------------------------------
// Device-side metadata (pointer tables, sizes, chunk prefix) is cached per distinct tensor list:
// optimizer buffers are persistent, so after the first step a launch costs no H2D traffic.
constexpr long kChunk = 8192;

struct MetaCache {
  std::unordered_map<uint64_t, std::pair<std::vector<int64_t>, Tensor>> map;
  Tensor get(const std::vector<int64_t>& host, const at::Device& dev) {
    uint64_t h = 1469598103934665603ull;
    for (auto v : host) { h ^= (uint64_t)v; h *= 1099511628211ull; }
    auto it = map.find(h);
    if (it != map.end() && it->second.first == host) return it->second.second;
    auto cpu = at::empty({(long)host.size()}, at::TensorOptions().dtype(at::kLong).pinned_memory(true));
    std::memcpy(cpu.data_ptr(), host.data(), host.size() * sizeof(int64_t));
    auto devt = cpu.to(dev, /*non_blocking=*/false);
    if (map.size() > 4096) map.clear();
    map[h] = {host, devt};
    return devt;
  }
};
MetaCache& meta_cache() { static MetaCache c; return c; }

// layout (int64 words): [n_lists * n pointers][n sizes][n+1 prefix][n_dt * n dtype codes packed 2/word as int32]
struct Packed { Tensor t; int n; };

Packed pack_meta(const std::vector<std::vector<Tensor>>& lists, const std::vector<int>& dtype_lists) {
  const int n = (int)lists[0].size();
  std::vector<int64_t> host;
  host.reserve(lists.size() * n + 2 * n + 1 + (dtype_lists.size() * n + 1) / 2 + 2);
  for (auto& l : lists) {
    TORCH_CHECK((int)l.size() == n, "multi-tensor lists differ in length");
    for (auto& t : l) host.push_back((int64_t)t.data_ptr());
  }
  for (auto& t : lists[0]) host.push_back(t.numel());
  int64_t acc = 0;
  host.push_back(0);
  for (auto& t : lists[0]) { acc += (t.numel() + kChunk - 1) / kChunk; host.push_back(acc); }
  std::vector<int32_t> dts;
  for (int li : dtype_lists) for (auto& t : lists[li]) dts.push_back(dtype_code(t));
  if (dts.size() % 2) dts.push_back(0);
  for (size_t i = 0; i < dts.size(); i += 2) host.push_back((int64_t)(uint32_t)dts[i] | ((int64_t)(uint32_t)dts[i + 1] << 32));
  return {meta_cache().get(host, lists[0][0].device()), n};
}

Tensor multi_l2norm(const std::vector<Tensor>& tensors) {
  TORCH_CHECK(!tensors.empty());
  for (auto& t : tensors) check_cuda_contig(t, "tensor");
  c10::cuda::CUDAGuard g(tensors[0].device());
  auto pk = pack_meta({tensors}, {0});
  const int n = pk.n;
  const int64_t* base = pk.t.data_ptr<int64_t>();
  const int nblocks = persistent_blocks();
  auto partial = at::empty({nblocks}, tensors[0].options().dtype(at::kFloat));
  auto out = at::empty({}, tensors[0].options().dtype(at::kFloat));
  mb200_multi_l2norm((const void* const*)base, (const long*)(base + n), (const int*)(base + 3 * n + 1), n, partial.data_ptr<float>(),
                     out.data_ptr<float>(), nblocks, cur_stream());
  return out;
}

void multi_scale(const std::vector<Tensor>& tensors, const Tensor& scale) {
  if (tensors.empty()) return;
  c10::cuda::CUDAGuard g(tensors[0].device());
  auto pk = pack_meta({tensors}, {0});
  const int n = pk.n;
  const int64_t* base = pk.t.data_ptr<int64_t>();
  mb200_multi_scale((void* const*)base, (const long*)(base + n), (const int*)(base + 3 * n + 1), n, scale.data_ptr<float>(), persistent_blocks(),
                    cur_stream());
}

void multi_adam(const std::vector<Tensor>& p32, const std::vector<Tensor>& grads, const std::vector<Tensor>& m, const std::vector<Tensor>& v,
                const std::vector<Tensor>& lowp, double lr, double b1, double b2, double eps, double wd, int64_t step, bool adamw,
                const Tensor& grad_scale) {
  if (p32.empty()) return;
  for (size_t i = 0; i < p32.size(); ++i) {
    TORCH_CHECK(p32[i].scalar_type() == at::kFloat && m[i].scalar_type() == at::kFloat && v[i].scalar_type() == at::kFloat,
                "master weights and Adam moments must be fp32");
    TORCH_CHECK(p32[i].is_contiguous() && grads[i].is_contiguous() && m[i].is_contiguous() && v[i].is_contiguous() && lowp[i].is_contiguous());
    TORCH_CHECK(p32[i].numel() == grads[i].numel() && p32[i].numel() == lowp[i].numel());
  }
  c10::cuda::CUDAGuard g(p32[0].device());
  // pointer lists: 0 p32, 1 grads, 2 m, 3 v, 4 lowp; sizes; prefix; dtype lists: grads(1), lowp(4)
  auto pk = pack_meta({p32, grads, m, v, lowp}, {1, 4});
  const int n = pk.n;
  const int64_t* base = pk.t.data_ptr<int64_t>();
  const int64_t* sizes = base + 5 * n;
  const int* dts = (const int*)(base + 5 * n + n + (n + 1));
  const double bc1 = 1.0 - std::pow(b1, (double)step), bc2 = 1.0 - std::pow(b2, (double)step);
  mb200_multi_adam((float* const*)base, (const void* const*)(base + n), (float* const*)(base + 2 * n), (float* const*)(base + 3 * n),
                   (void* const*)(base + 4 * n), (const long*)sizes, dts, dts + n, n, (float)lr, (float)b1, (float)b2, (float)eps, (float)wd,
                   (float)bc1, (float)bc2, adamw, grad_scale.data_ptr<float>(), persistent_blocks(), cur_stream());
}

// ---- GEMM -------------------------------------------------------------------------------------------
void gemm_bf16(const Tensor& a, const Tensor& b, Tensor c, int64_t layout, bool accumulate, int64_t variant) {
  check_cuda_contig(a, "a"); check_cuda_contig(b, "b"); check_cuda_contig(c, "c");
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16, "gemm_bf16 expects bf16 operands");
  TORCH_CHECK(c.scalar_type() == at::kBFloat16 || c.scalar_type() == at::kFloat, "gemm_bf16 output must be bf16 or fp32");
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && c.dim() == 2);
  int64_t M, N, K;
  if (layout == 0) { M = a.size(0); K = a.size(1); N = b.size(0); TORCH_CHECK(b.size(1) == K); }
  else if (layout == 1) { M = a.size(0); K = a.size(1); N = b.size(1); TORCH_CHECK(b.size(0) == K); }
  else { K = a.size(0); M = a.size(1); N = b.size(1); TORCH_CHECK(b.size(0) == K); }
  TORCH_CHECK(c.size(0) == M && c.size(1) == N, "gemm_bf16: output shape mismatch");
  check_aligned16(a, "a"); check_aligned16(b, "b"); check_aligned16(c, "c");
  TORCH_CHECK((a.size(1) % 8) == 0 && (b.size(1) % 8) == 0, "gemm_bf16: row pitch must be a multiple of 16 bytes");
  c10::cuda::CUDAGuard g(a.device());
  const int rc = mb200_gemm_bf16_v(a.data_ptr(), b.data_ptr(), c.data_ptr(), (int)M, (int)N, (int)K, (int)layout, accumulate ? 1 : 0, dtype_code(c),
                                   (int)variant, cur_stream());
  TORCH_CHECK(rc == 0, "mb200_gemm_bf16 failed with code ", rc);
}

// ---- NVLink collectives over symmetric memory ------------------------------------------------------
// `ptrs` / `flags`: per-rank peer-mapped addresses (python ints); `mc`: multicast address or 0.
#ifdef MB200_HAVE_NVLINK_COLLECTIVES
void nvl_barrier(const std::vector<int64_t>& ptrs, const std::vector<int64_t>& flags, int64_t rank, int64_t epoch, int64_t slot) {
  mb200_nvl_barrier(ptrs.data(), flags.data(), (int)rank, (int)ptrs.size(), (uint32_t)epoch, (int)slot, cur_stream());
}
void nvl_allgather(const std::vector<int64_t>& ptrs, const std::vector<int64_t>& flags, int64_t mc, const Tensor& src, int64_t dst_off_bytes, int64_t rank,
                   int64_t epoch, const Tensor& ctrl, int64_t slot, int64_t nblocks) {
  check_cuda_contig(src, "src");
  check_aligned16(src, "src");
  const size_t bytes = (size_t)src.numel() * src.element_size();
  TORCH_CHECK(bytes % 16 == 0 && dst_off_bytes % 16 == 0, "all-gather shard must be a multiple of 16 bytes");
  c10::cuda::CUDAGuard g(src.device());
  mb200_nvl_allgather(ptrs.data(), flags.data(), mc, src.data_ptr(), (size_t)dst_off_bytes, bytes, (int)rank, (int)ptrs.size(), (uint32_t)epoch,
                      ctrl.data_ptr(), (int)slot, (int)nblocks, cur_stream());
}
void nvl_reducescatter(const std::vector<int64_t>& ptrs, const std::vector<int64_t>& flags, int64_t mc, int64_t src_off_bytes, Tensor out, double scale,
                       int64_t rank, int64_t epoch, const Tensor& ctrl, int64_t slot, bool trailing, int64_t nblocks) {
  check_cuda_contig(out, "out");
  check_aligned16(out, "out");
  TORCH_CHECK(out.scalar_type() == at::kBFloat16 || out.scalar_type() == at::kFloat, "reduce-scatter supports bf16/fp32");
  TORCH_CHECK(((size_t)out.numel() * out.element_size()) % 16 == 0 && src_off_bytes % 16 == 0);
  c10::cuda::CUDAGuard g(out.device());
  mb200_nvl_reducescatter(ptrs.data(), flags.data(), mc, (size_t)src_off_bytes, out.data_ptr(), (size_t)out.numel(), (float)scale, dtype_code(out), (int)rank,
                          (int)ptrs.size(), (uint32_t)epoch, ctrl.data_ptr(), (int)slot, trailing ? 1 : 0, (int)nblocks, cur_stream());
}
void nvl_allreduce(const std::vector<int64_t>& ptrs, const std::vector<int64_t>& flags, int64_t mc, int64_t off_bytes, int64_t elems, int64_t dtype, double scale,
                   int64_t rank, int64_t epoch, const Tensor& ctrl, int64_t slot, int64_t nblocks) {
  c10::cuda::CUDAGuard g(ctrl.device());
  mb200_nvl_allreduce(ptrs.data(), flags.data(), mc, (size_t)off_bytes, (size_t)elems, (float)scale, (int)dtype, (int)rank, (int)ptrs.size(), (uint32_t)epoch,
                      ctrl.data_ptr(), (int)slot, (int)nblocks, cur_stream());
}
#endif

#ifdef MB200_HAVE_FLASH_ATTN_SM100
// q [sq, b, hq, d], k/v [sk, b, hk, d] (any s/b/h strides, d contiguous) -> (out [sq, b, hq, d], lse [b, hq, sq] fp32)
// row_lo (optional int32 [sq]): first visible key of every query, monotone non-decreasing — the band mask of sliding windows / packed sequences (with b == 1)
std::vector<Tensor> flash_attn_fwd(const Tensor& q, const Tensor& k, const Tensor& v, bool causal, double scale, int64_t variant, const c10::optional<Tensor>& row_lo) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && k.scalar_type() == at::kBFloat16 && v.scalar_type() == at::kBFloat16, "flash_attn_fwd: bf16 CUDA tensors");
  TORCH_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4 && q.stride(3) == 1 && k.stride(3) == 1 && v.stride(3) == 1, "flash_attn_fwd: [s, b, h, d] with contiguous d");
  TORCH_CHECK(((uintptr_t)q.data_ptr() | (uintptr_t)k.data_ptr() | (uintptr_t)v.data_ptr()) % 16 == 0, "flash_attn_fwd: 16-byte aligned tensors");
  c10::cuda::CUDAGuard g(q.device());
  const int sq = (int)q.size(0), b = (int)q.size(1), hq = (int)q.size(2), d = (int)q.size(3), sk = (int)k.size(0), hk = (int)k.size(2);
  TORCH_CHECK(!causal || sk >= sq, "flash_attn_fwd: causal needs sk >= sq");
  auto out = at::empty({sq, b, hq, d}, q.options());
  auto lse = at::empty({b, hq, sq}, q.options().dtype(at::kFloat));
  const int* lo = nullptr;
  if (row_lo.has_value() && row_lo->defined()) {
    TORCH_CHECK(causal && row_lo->is_cuda() && row_lo->scalar_type() == at::kInt && row_lo->is_contiguous() && row_lo->numel() == sq, "flash_attn_fwd: row_lo must be int32 [sq] on the GPU, with a causal mask");
    lo = row_lo->data_ptr<int>();
  }
  const int rc = mb200_flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), sq, sk, b, hq, hk, d, q.stride(0), q.stride(1),
                                      q.stride(2), k.stride(0), k.stride(1), k.stride(2), v.stride(0), v.stride(1), v.stride(2), (float)scale, causal ? 1 : 0,
                                      (int)variant, lo, cur_stream());
  TORCH_CHECK(rc == 0, "flash_attn_fwd failed with code ", rc);
  return {out, lse};
}
#endif

#ifdef MB200_HAVE_FLASH_ATTN_BWD_SM100
// go, q, o [sq,b,hq,128]; k, v [sk,b,hk,128]; lse [b,hq,sq] fp32 -> (dq, dk, dv) bf16.  split_heads: -1 = decide from the grid size.
// Heads are processed in chunks so that the bf16 dS scratch ([heads, sk, sq]) stays under max_scratch_mb.
std::vector<Tensor> flash_attn_bwd(const Tensor& go, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& o, const Tensor& lse, bool causal, double scale,
                                   int64_t split_heads, int64_t max_scratch_mb, const c10::optional<Tensor>& row_lo, const c10::optional<Tensor>& col_hi) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && go.scalar_type() == at::kBFloat16 && o.scalar_type() == at::kBFloat16, "flash_attn_bwd: bf16 CUDA tensors");
  TORCH_CHECK(q.stride(3) == 1 && k.stride(3) == 1 && v.stride(3) == 1 && go.stride(3) == 1 && o.stride(3) == 1, "flash_attn_bwd: contiguous head dim");
  TORCH_CHECK(lse.is_contiguous() && lse.scalar_type() == at::kFloat, "flash_attn_bwd: fp32 [b,h,s] log-sum-exp");
  c10::cuda::CUDAGuard g(q.device());
  const int sq = (int)q.size(0), b = (int)q.size(1), hq = (int)q.size(2), d = (int)q.size(3), sk = (int)k.size(0), hk = (int)k.size(2);
  auto dq = at::empty({sq, b, hq, d}, q.options());
  auto dk = at::empty({sk, b, hk, d}, q.options());
  auto dv = at::empty({sk, b, hk, d}, q.options());
  auto delta = at::empty({b, hq, (sq + 63) / 64, 128}, q.options().dtype(at::kFloat));   // per 64-query block: lse*log2e | rowsum(dO o O)*scale
  const int *lo = nullptr, *hi = nullptr;
  if (row_lo.has_value() && row_lo->defined()) {
    TORCH_CHECK(col_hi.has_value() && col_hi->defined(), "flash_attn_bwd: row_lo and col_hi come together");
    TORCH_CHECK(causal && row_lo->scalar_type() == at::kInt && col_hi->scalar_type() == at::kInt && row_lo->is_contiguous() && col_hi->is_contiguous() && row_lo->numel() == sq && col_hi->numel() == sk,
                "flash_attn_bwd: row_lo int32 [sq], col_hi int32 [sk], causal mask");
    lo = row_lo->data_ptr<int>();
    hi = col_hi->data_ptr<int>();
  }
  const int sh = split_heads < 0 ? mb200_flash_attn_bwd_split_heads(sk, b, hq, hk) : (int)split_heads;
  const size_t bytes = mb200_flash_attn_bwd_scratch_bytes(sq, sk, b, hq, hk, sh);
  auto scratch = at::empty({(int64_t)bytes}, q.options().dtype(at::kByte));
  const int rc = mb200_flash_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), go.data_ptr(), o.data_ptr(), lse.data_ptr<float>(), delta.data_ptr<float>(), dq.data_ptr(),
                                      dk.data_ptr(), dv.data_ptr(), scratch.data_ptr(), sh, sq, sk, b, hq, hk, d, q.stride(0), q.stride(1), q.stride(2), k.stride(0),
                                      k.stride(1), k.stride(2), v.stride(0), v.stride(1), v.stride(2), go.stride(0), go.stride(1), go.stride(2), o.stride(0), o.stride(1),
                                      o.stride(2), (float)scale, causal ? 1 : 0, lo, hi, cur_stream());
  TORCH_CHECK(rc == 0, "flash_attn_bwd failed with code ", rc);
  return {dq, dk, dv};
}
#endif

// dst's storage becomes an alias of src's storage (reference N4 ``share_storage``, tensor_parallel/random.py:41-80): lets
// CheckpointWithoutOutput hand a recomputed activation back to every tensor that still references the discarded one.
void share_storage(Tensor dst, const Tensor& src) {
  TORCH_CHECK(dst.nbytes() <= src.storage().nbytes(), "share_storage: source storage too small");
  dst.set_(src.storage(), dst.storage_offset(), dst.sizes(), dst.strides());
}

#ifdef MB200_HAVE_ROUTING_KERNELS
// ---- router bookkeeping + MLA rotary (routing_kernels.cu) ---------------------------------------------------------------------------------------------
std::vector<Tensor> indices_to_multihot(const Tensor& idx, const Tensor& probs, int64_t E) {
  check_cuda_contig(idx, "indices"); check_cuda_contig(probs, "probs");
  TORCH_CHECK(idx.dim() == 2 && idx.scalar_type() == at::kLong && probs.scalar_type() == at::kFloat && probs.sizes() == idx.sizes(), "indices_to_multihot: int64 [T, k] + fp32 [T, k]");
  c10::cuda::CUDAGuard g(idx.device());
  auto map = at::empty({idx.size(0), E}, idx.options().dtype(at::kBool));
  auto out = at::empty({idx.size(0), E}, probs.options());
  mb200_indices_to_multihot(idx.data_ptr<int64_t>(), probs.data_ptr<float>(), (uint8_t*)map.data_ptr(), out.data_ptr<float>(), idx.size(0), (int)idx.size(1), (int)E, cur_stream());
  return {map, out};
}
// scatter = false: g [T, E] -> [T, k]; true: g [T, k] -> [T, E]
Tensor multihot_probs_grad(const Tensor& idx, const Tensor& g, int64_t E, bool scatter) {
  check_cuda_contig(idx, "indices"); check_cuda_contig(g, "grad");
  TORCH_CHECK(idx.scalar_type() == at::kLong && g.scalar_type() == at::kFloat && g.size(0) == idx.size(0) && g.size(1) == (scatter ? idx.size(1) : E));
  c10::cuda::CUDAGuard gd(idx.device());
  auto out = at::empty({idx.size(0), scatter ? E : idx.size(1)}, g.options());
  mb200_multihot_probs_grad(idx.data_ptr<int64_t>(), g.data_ptr<float>(), out.data_ptr<float>(), idx.size(0), (int)idx.size(1), (int)E, scatter ? 1 : 0, cur_stream());
  return out;
}
std::vector<Tensor> multihot_to_indices(const Tensor& map, const Tensor& probs, int64_t k) {
  check_cuda_contig(map, "map"); check_cuda_contig(probs, "probs");
  TORCH_CHECK(map.dim() == 2 && map.scalar_type() == at::kBool && probs.scalar_type() == at::kFloat && probs.sizes() == map.sizes(), "multihot_to_indices: bool [T, E] + fp32 [T, E]");
  c10::cuda::CUDAGuard g(map.device());
  auto idx = at::empty({map.size(0), k}, map.options().dtype(at::kLong));
  auto out = at::empty({map.size(0), k}, probs.options());
  mb200_multihot_to_indices((const uint8_t*)map.data_ptr(), probs.data_ptr<float>(), idx.data_ptr<int64_t>(), out.data_ptr<float>(), map.size(0), (int)k, (int)map.size(1), cur_stream());
  return {idx, out};
}
Tensor pad_routing_map(const Tensor& map, int64_t multiple) {
  check_cuda_contig(map, "map");
  TORCH_CHECK(map.dim() == 2 && map.scalar_type() == at::kBool && multiple > 0, "pad_routing_map: bool [T, E]");
  c10::cuda::CUDAGuard g(map.device());
  auto out = at::empty_like(map);
  mb200_pad_routing_map((const uint8_t*)map.data_ptr(), (uint8_t*)out.data_ptr(), map.size(0), (int)map.size(1), (int)multiple, cur_stream());
  return out;
}
Tensor moe_aux_loss_fwd(const Tensor& probs, const Tensor& tpe, double coeff) {
  check_cuda_contig(probs, "probs"); check_cuda_contig(tpe, "tokens_per_expert");
  TORCH_CHECK(probs.dim() == 2 && probs.scalar_type() == at::kFloat && tpe.scalar_type() == at::kFloat && tpe.numel() == probs.size(1), "moe_aux_loss: fp32 probs [T, E], fp32 tokens_per_expert [E]");
  c10::cuda::CUDAGuard g(probs.device());
  const int nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(296, probs.size(0) / 8));
  auto partial = at::empty({nblocks}, probs.options());
  auto loss = at::empty({}, probs.options());
  mb200_moe_aux_loss_fwd(probs.data_ptr<float>(), tpe.data_ptr<float>(), partial.data_ptr<float>(), nblocks, loss.data_ptr<float>(), probs.size(0), (int)probs.size(1), (float)coeff, cur_stream());
  return loss;
}
Tensor moe_aux_loss_bwd(const Tensor& tpe, const Tensor& gloss, double coeff, int64_t T) {
  check_cuda_contig(tpe, "tokens_per_expert");
  TORCH_CHECK(tpe.scalar_type() == at::kFloat && gloss.scalar_type() == at::kFloat && gloss.numel() == 1 && gloss.is_cuda());
  c10::cuda::CUDAGuard g(tpe.device());
  auto gp = at::empty({T, tpe.numel()}, tpe.options());
  mb200_moe_aux_loss_bwd(tpe.data_ptr<float>(), gloss.data_ptr<float>(), (float)coeff, gp.data_ptr<float>(), T, (int)tpe.numel(), cur_stream());
  return gp;
}
// x [..., H, nope + emb] contiguous, rotated IN PLACE; ang fp32 [positions, emb]; pos (optional int64 [rows]) else position = row / batch
// out-of-place (returns a new tensor: untouched channels copied in the same pass) unless `inplace`
Tensor mla_rope_inplace(Tensor x, const Tensor& ang, const c10::optional<Tensor>& pos, int64_t nope, int64_t emb, int64_t batch, double mscale, bool interleaved, bool inverse,
                        bool inplace) {
  check_cuda_contig(x, "x"); check_cuda_contig(ang, "angles");
  TORCH_CHECK(x.dim() >= 2 && x.size(-1) == nope + emb && ang.scalar_type() == at::kFloat && ang.size(-1) == emb, "mla_rope_inplace: x [..., H, nope + emb], angles fp32 [positions, emb]");
  const int H = (int)x.size(-2);
  const long rows = x.numel() / (H * (nope + emb));
  const int64_t* pp = nullptr;
  if (pos.has_value() && pos->defined()) {
    TORCH_CHECK(pos->is_cuda() && pos->scalar_type() == at::kLong && pos->is_contiguous() && pos->numel() == rows, "mla_rope_inplace: int64 positions, one per row");
    pp = pos->data_ptr<int64_t>();
  } else {
    TORCH_CHECK(ang.numel() / emb >= (rows + batch - 1) / batch, "mla_rope_inplace: not enough angle rows");
  }
  c10::cuda::CUDAGuard g(x.device());
  Tensor out = inplace ? x : at::empty_like(x);
  const int rc = mb200_mla_rope_inplace(x.data_ptr(), out.data_ptr(), ang.data_ptr<float>(), pp, rows, H, (int)nope, (int)emb, (int)batch, (float)mscale, interleaved, inverse,
                                        dtype_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "mla_rope_inplace: rotary dim must be even and <= 64");
  return out;
}
// forward: (kv [rows.., H, kd + vd], k_pe [rows.., emb]) -> (key [.., H, kd + emb], value [.., H, vd]); backward: (dkey, dvalue) -> (dkv, dk_pe)
std::vector<Tensor> mla_kv_split(const Tensor& a, const Tensor& b, const c10::optional<Tensor>& ang, const c10::optional<Tensor>& pos, int64_t kd, int64_t vd, int64_t emb, int64_t batch,
                                 double mscale, bool interleaved, bool backward) {
  check_cuda_contig(a, "a"); check_cuda_contig(b, "b");
  TORCH_CHECK(a.scalar_type() == b.scalar_type() && a.dim() >= 2);
  const int H = (int)a.size(-2);
  const long rows = a.numel() / (H * a.size(-1));
  if (!backward) { TORCH_CHECK(a.size(-1) == kd + vd && b.numel() == rows * emb, "mla_kv_split: kv [.., H, kd + vd], k_pe [.., emb]"); }
  else { TORCH_CHECK(a.size(-1) == kd + emb && b.size(-1) == vd && b.numel() == rows * H * vd, "mla_kv_split backward: dkey [.., H, kd + emb], dvalue [.., H, vd]"); }
  const float* ap = nullptr;
  const int64_t* pp = nullptr;
  if (ang.has_value() && ang->defined()) {
    TORCH_CHECK(ang->is_cuda() && ang->scalar_type() == at::kFloat && ang->is_contiguous() && ang->size(-1) == emb);
    ap = ang->data_ptr<float>();
  }
  if (pos.has_value() && pos->defined()) {
    TORCH_CHECK(pos->is_cuda() && pos->scalar_type() == at::kLong && pos->is_contiguous() && pos->numel() == rows);
    pp = pos->data_ptr<int64_t>();
  }
  c10::cuda::CUDAGuard g(a.device());
  auto lead = a.sizes().vec();
  lead.pop_back();
  auto s0 = lead, s1 = lead;
  s0.push_back(backward ? kd + vd : kd + emb);
  if (backward) { s1.pop_back(); s1.push_back(1); s1.push_back(emb); } else { s1.push_back(vd); }
  auto o0 = at::empty(s0, a.options()), o1 = at::empty(s1, a.options());
  const int rc = mb200_mla_kv_split(a.data_ptr(), b.data_ptr(), ap, pp, o0.data_ptr(), o1.data_ptr(), rows, H, (int)kd, (int)vd, (int)emb, (int)batch, (float)mscale, interleaved,
                                    backward, dtype_code(a), cur_stream());
  TORCH_CHECK(rc == 0, "mla_kv_split: rotary dim must be even and <= 256");
  return {o0, o1};
}
#endif

#ifdef MB200_HAVE_MISC_KERNELS
// ---- paged stash / speculative verify / bias-dropout-add (misc_kernels.cu) ------------------------------------------------------------------------------
// flat [T_max, H] <-> pages [(P + 1) * page_size, H]; page_ids int64 [ceil(T_max / page_size)] (scratch id in unused slots); num_tokens int64 device scalar
void paged_stash(Tensor flat, Tensor pages, const Tensor& page_ids, const Tensor& num_tokens, int64_t page_size, bool pop) {
  check_cuda_contig(flat, "flat"); check_cuda_contig(pages, "pages"); check_cuda_contig(page_ids, "page_ids");
  TORCH_CHECK(flat.dim() == 2 && flat.scalar_type() == pages.scalar_type() && page_ids.scalar_type() == at::kLong && num_tokens.is_cuda() && num_tokens.scalar_type() == at::kLong &&
              num_tokens.numel() == 1, "paged_stash: flat [T, H], int64 page ids, int64 device token count");
  TORCH_CHECK(page_ids.numel() * page_size >= flat.size(0), "paged_stash: page table too short");
  check_aligned16(flat, "flat"); check_aligned16(pages, "pages");
  c10::cuda::CUDAGuard g(flat.device());
  const int rc = mb200_paged_stash(flat.data_ptr(), pages.data_ptr(), page_ids.data_ptr<int64_t>(), num_tokens.data_ptr<int64_t>(), flat.size(0), flat.size(1) * flat.element_size(),
                                   (int)page_size, pop ? 1 : 0, cur_stream());
  TORCH_CHECK(rc == 0, "paged_stash: row size must be a multiple of 16 bytes");
}
// -> (n_accepted int64 [B], next_token int64 [B])
std::vector<Tensor> spec_verify(const Tensor& draft_tokens, const c10::optional<Tensor>& draft_probs, const Tensor& target_probs, const Tensor& u_accept, const Tensor& u_sample) {
  check_cuda_contig(draft_tokens, "draft_tokens"); check_cuda_contig(target_probs, "target_probs"); check_cuda_contig(u_accept, "u_accept"); check_cuda_contig(u_sample, "u_sample");
  const int B = (int)draft_tokens.size(0), k = (int)draft_tokens.size(1), V = (int)target_probs.size(2);
  TORCH_CHECK(draft_tokens.scalar_type() == at::kLong && target_probs.scalar_type() == at::kFloat && target_probs.size(0) == B && target_probs.size(1) == k + 1 &&
              u_accept.scalar_type() == at::kFloat && u_accept.numel() == (int64_t)B * k && u_sample.scalar_type() == at::kFloat && u_sample.numel() == B, "spec_verify: shapes");
  const float* dp = nullptr;
  if (draft_probs.has_value() && draft_probs->defined()) {
    check_cuda_contig(*draft_probs, "draft_probs");
    TORCH_CHECK(draft_probs->scalar_type() == at::kFloat && draft_probs->numel() == (int64_t)B * k * V);
    dp = draft_probs->data_ptr<float>();
  }
  c10::cuda::CUDAGuard g(draft_tokens.device());
  auto n = at::empty({B}, draft_tokens.options()), nxt = at::empty({B}, draft_tokens.options());
  mb200_spec_verify(draft_tokens.data_ptr<int64_t>(), dp, target_probs.data_ptr<float>(), u_accept.data_ptr<float>(), u_sample.data_ptr<float>(), n.data_ptr<int64_t>(),
                    nxt.data_ptr<int64_t>(), B, k, V, cur_stream());
  return {n, nxt};
}
// y = residual + dropout(x + bias, p); the mask is regenerated in the backward from (seed, offset) taken from the current CUDA generator here
std::tuple<Tensor, int64_t, int64_t> bias_dropout_add_fwd(const Tensor& x, const c10::optional<Tensor>& bias, const Tensor& residual, double p) {
  check_cuda_contig(x, "x"); check_cuda_contig(residual, "residual");
  TORCH_CHECK(x.sizes() == residual.sizes() && x.scalar_type() == residual.scalar_type(), "bias_dropout_add: x and residual must match");
  check_aligned16(x, "x"); check_aligned16(residual, "residual");
  const int H = (int)x.size(-1);
  const void* bp = nullptr;
  if (bias.has_value() && bias->defined()) {
    check_cuda_contig(*bias, "bias");
    TORCH_CHECK(bias->numel() == H && bias->scalar_type() == x.scalar_type(), "bias_dropout_add: bias [H] in x's dtype");
    check_aligned16(*bias, "bias");
    bp = bias->data_ptr();
  }
  c10::cuda::CUDAGuard g(x.device());
  uint64_t seed = 0, offset = 0;
  if (p > 0.0) {
    int grid;
    const long draws = mb200_bias_dropout_add_draws(x.numel(), dtype_code(x), &grid);
    auto gen = at::get_generator_or_default<at::CUDAGeneratorImpl>(c10::nullopt, at::cuda::detail::getDefaultCUDAGenerator());
    std::lock_guard<std::mutex> lock(gen->mutex_);
    at::PhiloxCudaState st = gen->philox_cuda_state((uint64_t)draws);
    TORCH_CHECK(!st.captured_, "bias_dropout_add: not capturable (the Python wrapper uses the eager path under CUDA-graph capture)");
    seed = st.seed_.val;
    offset = st.offset_.val;
  }
  auto y = at::empty_like(x);
  const int rc = mb200_bias_dropout_add(x.data_ptr(), bp, residual.data_ptr(), y.data_ptr(), x.numel(), H, (float)p, seed, offset, 0, dtype_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "bias_dropout_add: sizes must be multiples of the 16-byte vector");
  return {y, (int64_t)seed, (int64_t)offset};
}
Tensor bias_dropout_add_bwd(const Tensor& gy, double p, int64_t seed, int64_t offset) {
  check_cuda_contig(gy, "gy");
  c10::cuda::CUDAGuard g(gy.device());
  auto gx = at::empty_like(gy);
  const int rc = mb200_bias_dropout_add(gy.data_ptr(), nullptr, nullptr, gx.data_ptr(), gy.numel(), (int)gy.size(-1), (float)p, (uint64_t)seed, (uint64_t)offset, 1, dtype_code(gy),
                                        cur_stream());
  TORCH_CHECK(rc == 0, "bias_dropout_add_bwd: sizes must be multiples of the 16-byte vector");
  return gx;
}
#endif

#ifdef MB200_HAVE_PAGED_ATTENTION
// k_new, v_new [B, hk, d] -> pools [num_blocks, block_size, hk, d] (one layer) at block_table[b, positions[b] / block_size], offset positions[b] % block_size
void paged_kv_append(const Tensor& k_new, const Tensor& v_new, Tensor k_pool, Tensor v_pool, const Tensor& block_table, const Tensor& positions) {
  TORCH_CHECK(k_new.is_cuda() && k_new.scalar_type() == at::kBFloat16 && k_pool.scalar_type() == at::kBFloat16 && k_new.is_contiguous() && v_new.is_contiguous(), "paged_kv_append: contiguous bf16 CUDA k/v");
  TORCH_CHECK(k_pool.is_contiguous() && v_pool.is_contiguous() && k_pool.dim() == 4, "paged_kv_append: pools [num_blocks, block_size, hk, d] contiguous");
  TORCH_CHECK(block_table.scalar_type() == at::kInt && positions.scalar_type() == at::kInt && block_table.is_contiguous() && positions.is_contiguous(), "paged_kv_append: int32 block table / positions");
  c10::cuda::CUDAGuard g(k_new.device());
  const int B = (int)k_new.size(0), hk = (int)k_new.size(1), d = (int)k_new.size(2);
  TORCH_CHECK((hk * d) % 8 == 0 && k_pool.size(2) == hk && k_pool.size(3) == d);
  mb200_paged_kv_append(k_new.data_ptr(), v_new.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), block_table.data_ptr<int32_t>(), positions.data_ptr<int32_t>(), B,
                        (int)block_table.size(1), (int)k_pool.size(1), hk, d, cur_stream());
}

// q [B, hq, d] -> out [B, hq, d]: attention of one new token per request over its pages (lengths include the new token)
// fixed_tokens_per_split > 0: split boundaries at fixed absolute positions irrespective of the batch (batch-invariant mode: a request's result is then
// bitwise independent of what it is co-scheduled with); 0: size the splits for two waves of CTAs
Tensor paged_decode(const Tensor& q, const Tensor& k_pool, const Tensor& v_pool, const Tensor& block_table, const Tensor& lengths, double scale, int64_t max_len,
                    int64_t fixed_tokens_per_split) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && q.is_contiguous() && k_pool.is_contiguous() && v_pool.is_contiguous() && k_pool.scalar_type() == at::kBFloat16, "paged_decode: contiguous bf16 CUDA tensors");
  TORCH_CHECK(block_table.scalar_type() == at::kInt && lengths.scalar_type() == at::kInt && block_table.is_contiguous() && lengths.is_contiguous(), "paged_decode: int32 block table / lengths");
  c10::cuda::CUDAGuard g(q.device());
  const int B = (int)q.size(0), hq = (int)q.size(1), d = (int)q.size(2), hk = (int)k_pool.size(2), bs = (int)k_pool.size(1);
  // enough CTAs for two waves; a split is a whole number of 128-token steps (4 warps x 32 tokens)
  int nsplit = (int)((2 * 148 + (int64_t)B * hk - 1) / ((int64_t)B * hk));
  const int max_steps = (int)((max_len + 127) / 128);
  nsplit = std::max(1, std::min(nsplit, max_steps));
  int tokens_per_split = ((max_steps + nsplit - 1) / nsplit) * 128;
  if (fixed_tokens_per_split > 0) tokens_per_split = (int)((fixed_tokens_per_split + 127) / 128 * 128);
  nsplit = (int)((max_len + tokens_per_split - 1) / tokens_per_split);
  nsplit = std::max(1, nsplit);
  auto o_part = at::empty({(int64_t)B * hq * nsplit * d}, q.options().dtype(at::kFloat));
  auto ml_part = at::empty({(int64_t)B * hq * nsplit * 2}, q.options().dtype(at::kFloat));
  auto out = at::empty_like(q);
  const int rc = mb200_paged_decode(q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), block_table.data_ptr<int32_t>(), lengths.data_ptr<int32_t>(), o_part.data_ptr<float>(),
                                    ml_part.data_ptr<float>(), out.data_ptr(), B, hq, hk, d, (int)block_table.size(1), bs, (float)scale, nsplit, tokens_per_split, cur_stream());
  TORCH_CHECK(rc == 0, "paged_decode: unsupported head dim / GQA ratio (d in {64,128}, hq/hk in {1,2,4,8}); code ", rc);
  return out;
}
#endif

#ifdef MB200_HAVE_RUNTIME_NATIVE
// tasks: int64 [n, 3] rows of (src_ptr, dst_ptr, nbytes) on the host; copied to the device and executed by ONE kernel
void batched_copy(const Tensor& tasks_cpu, int64_t nblocks) {
  TORCH_CHECK(!tasks_cpu.is_cuda() && tasks_cpu.scalar_type() == at::kLong && tasks_cpu.dim() == 2 && tasks_cpu.size(1) == 3, "batched_copy: int64 [n, 3] CPU tensor");
  const int n = (int)tasks_cpu.size(0);
  if (n == 0) return;
  auto t = tasks_cpu.contiguous();
  const int64_t* a = t.data_ptr<int64_t>();
  auto prefix = at::empty({n}, t.options());
  int64_t* pf = prefix.data_ptr<int64_t>();
  unsigned long long total = 0;
  for (int i = 0; i < n; ++i) {
    pf[i] = (int64_t)total;
    total += ((unsigned long long)a[3 * i + 2] + 65535ull) / 65536ull;
  }
  auto dev = at::Device(at::kCUDA, c10::cuda::current_device());
  auto td = t.to(dev, /*non_blocking=*/false);
  auto pd = prefix.to(dev, false);
  mb200_batched_copy(td.data_ptr(), pd.data_ptr(), n, total, (int)nblocks, cur_stream());
  c10::cuda::CUDACachingAllocator::recordStream(td.storage().data_ptr(), c10::cuda::getCurrentCUDAStream());
  c10::cuda::CUDACachingAllocator::recordStream(pd.storage().data_ptr(), c10::cuda::getCurrentCUDAStream());
}
#endif

#ifdef MB200_HAVE_GROUPED_GEMM_SM100
// mode 0: (x [T,K], w [E,N,K]) -> out [T,N];  mode 1: (gy [T,N], w [E,N,K]) -> out [T,K];  mode 2: (gy [T,N], x [T,K]) -> out = gw [E,N,K]
void grouped_gemm_bf16(const Tensor& a, const Tensor& b, Tensor out, const Tensor& offsets_cpu, int64_t mode, bool accumulate) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && a.is_contiguous() && b.is_contiguous() && out.is_contiguous(),
              "grouped_gemm_bf16: contiguous bf16 CUDA operands");
  TORCH_CHECK(!offsets_cpu.is_cuda() && offsets_cpu.scalar_type() == at::kInt, "grouped_gemm_bf16: offsets must be a CPU int32 tensor");
  c10::cuda::CUDAGuard g(a.device());
  const int E = (int)offsets_cpu.numel() - 1;
  const Tensor& w = mode == 2 ? out : b;
  TORCH_CHECK(w.dim() == 3 && w.size(0) == E, "grouped_gemm_bf16: weight-shaped operand must be [E, N, K]");
  const int dim_n = (int)w.size(1), dim_k = (int)w.size(2);
  const int c_dtype = out.scalar_type() == at::kFloat ? 0 : 1;
  TORCH_CHECK(mode == 2 || c_dtype == 1, "grouped_gemm_bf16: activations are bf16");
  auto maps = at::empty({(int64_t)(2 * E * 128 + 64)}, a.options().dtype(at::kByte));
  void* md = reinterpret_cast<void*>(((uintptr_t)maps.data_ptr() + 63) & ~(uintptr_t)63);
  const int rc = mb200_grouped_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), offsets_cpu.data_ptr<int>(), E, dim_n, dim_k, (int)mode, accumulate ? 1 : 0, c_dtype,
                                         md, cur_stream());
  TORCH_CHECK(rc == 0, "grouped_gemm_bf16 failed with code ", rc);
}
#endif

#ifdef MB200_HAVE_GEMM_FP8_SM100
// a [M,K], b [N,K]: torch.float8_e4m3fn / float8_e5m2 -> bf16 [M,N] = alpha * a · bᵀ
Tensor gemm_fp8_nt(const Tensor& a, const Tensor& b, double alpha, const c10::optional<Tensor>& alpha_dev) {
  auto fmt = [](const Tensor& t) {
    TORCH_CHECK(t.scalar_type() == at::kFloat8_e4m3fn || t.scalar_type() == at::kFloat8_e5m2, "gemm_fp8_nt: float8_e4m3fn / float8_e5m2 operands");
    return t.scalar_type() == at::kFloat8_e4m3fn ? 0 : 1;
  };
  TORCH_CHECK(a.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.is_contiguous() && b.is_contiguous() && a.size(1) == b.size(1), "gemm_fp8_nt: a [M,K], b [N,K] contiguous");
  c10::cuda::CUDAGuard g(a.device());
  auto c = at::empty({a.size(0), b.size(0)}, a.options().dtype(at::kBFloat16));
  const int rc = mb200_gemm_fp8_nt(a.data_ptr(), b.data_ptr(), c.data_ptr(), (int)a.size(0), (int)b.size(0), (int)a.size(1), fmt(a), fmt(b), (float)alpha,
                                   alpha_dev.has_value() ? alpha_dev->data_ptr<float>() : nullptr, cur_stream());
  TORCH_CHECK(rc == 0, "gemm_fp8_nt failed with code ", rc);
  return c;
}
#endif

#ifdef MB200_HAVE_MOE_KERNELS
// out[i] = scale[i] * in[src[i]]   (bf16 rows, hidden % 8 == 0)
Tensor moe_gather_rows(const Tensor& in, const Tensor& src, const c10::optional<Tensor>& scale) {
  TORCH_CHECK(in.is_cuda() && in.scalar_type() == at::kBFloat16 && in.dim() == 2 && in.is_contiguous() && in.size(1) % 8 == 0, "moe_gather_rows: contiguous bf16 [n, h], h % 8 == 0");
  TORCH_CHECK(src.scalar_type() == at::kLong && src.is_contiguous(), "moe_gather_rows: int64 indices");
  c10::cuda::CUDAGuard g(in.device());
  auto out = at::empty({src.numel(), in.size(1)}, in.options());
  mb200_moe_gather_rows(in.data_ptr(), out.data_ptr(), src.data_ptr<int64_t>(), scale.has_value() ? scale->data_ptr<float>() : nullptr, src.numel(), (int)in.size(1), cur_stream());
  return out;
}
// out[t] = sum_k w[t,k] * in[pos[t,k]]   (pos < 0 skipped)
Tensor moe_combine_rows(const Tensor& in, const Tensor& pos, const c10::optional<Tensor>& w) {
  TORCH_CHECK(in.is_cuda() && in.scalar_type() == at::kBFloat16 && in.dim() == 2 && in.is_contiguous() && in.size(1) % 8 == 0, "moe_combine_rows: contiguous bf16 [n, h], h % 8 == 0");
  TORCH_CHECK(pos.scalar_type() == at::kLong && pos.dim() == 2 && pos.is_contiguous(), "moe_combine_rows: int64 [T, k] positions");
  c10::cuda::CUDAGuard g(in.device());
  auto out = at::empty({pos.size(0), in.size(1)}, in.options());
  mb200_moe_combine_rows(in.data_ptr(), out.data_ptr(), pos.data_ptr<int64_t>(), w.has_value() ? w->data_ptr<float>() : nullptr, pos.size(0), (int)pos.size(1), (int)in.size(1),
                         cur_stream());
  return out;
}
// rows of `in` (any dtype, row size multiple of 16 bytes) -> slot dst_slot[i] of peer dst_rank[i]'s buffer at byte offset dst_off
void moe_push_rows(const Tensor& in, const Tensor& src_row, const Tensor& dst_rank, const Tensor& dst_slot, const std::vector<int64_t>& peer_ptrs, int64_t dst_off_bytes) {
  TORCH_CHECK(in.is_cuda() && in.dim() == 2 && in.is_contiguous(), "moe_push_rows: contiguous 2-D CUDA tensor");
  const int row_bytes = (int)(in.size(1) * in.element_size());
  TORCH_CHECK(row_bytes % 16 == 0 && dst_off_bytes % 16 == 0, "moe_push_rows: 16-byte rows");
  TORCH_CHECK(src_row.scalar_type() == at::kLong && dst_slot.scalar_type() == at::kLong && dst_rank.scalar_type() == at::kInt, "moe_push_rows: index dtypes");
  c10::cuda::CUDAGuard g(in.device());
  mb200_moe_push_rows(in.data_ptr(), src_row.data_ptr<int64_t>(), dst_rank.data_ptr<int>(), dst_slot.data_ptr<int64_t>(), peer_ptrs.data(), (int)peer_ptrs.size(),
                      (size_t)dst_off_bytes, src_row.numel(), row_bytes, cur_stream());
}
// out[t] = sum_k w[t,k] * peer[rank[t,k]].buf[slot[t,k]] (bf16 rows, fp32 accumulate); raw=true: k = 1, bitwise copy of any dtype
Tensor moe_pull_rows(const Tensor& rank, const Tensor& slot, const c10::optional<Tensor>& w, const std::vector<int64_t>& peer_ptrs, int64_t src_off_bytes, int64_t row_elems,
                     at::ScalarType dtype, bool raw) {
  TORCH_CHECK(rank.is_cuda() && rank.scalar_type() == at::kInt && slot.scalar_type() == at::kLong && rank.is_contiguous() && slot.is_contiguous(), "moe_pull_rows: index dtypes");
  c10::cuda::CUDAGuard g(rank.device());
  const int64_t n_out = slot.size(0);
  const int topk = slot.dim() == 2 ? (int)slot.size(1) : 1;
  auto out = at::empty({n_out, row_elems}, rank.options().dtype(dtype));
  const int row_bytes = (int)(row_elems * out.element_size());
  TORCH_CHECK(row_bytes % 16 == 0 && src_off_bytes % 16 == 0, "moe_pull_rows: 16-byte rows");
  TORCH_CHECK(raw || dtype == at::kBFloat16, "moe_pull_rows: weighted combine needs bf16 rows");
  mb200_moe_pull_rows(out.data_ptr(), rank.data_ptr<int>(), slot.data_ptr<int64_t>(), w.has_value() ? w->data_ptr<float>() : nullptr, peer_ptrs.data(), (int)peer_ptrs.size(),
                      (size_t)src_off_bytes, n_out, topk, row_bytes, raw ? 1 : 0, cur_stream());
  return out;
}
// logits fp32 [T, E] -> (probs [T,k] fp32, ids [T,k] int64, routing_map [T,E] bool, tokens_per_expert [E] int32)
std::vector<Tensor> moe_topk_router(const Tensor& logits, const c10::optional<Tensor>& expert_bias, int64_t topk, int64_t score_fn, bool renormalize, double scaling) {
  TORCH_CHECK(logits.is_cuda() && logits.scalar_type() == at::kFloat && logits.dim() == 2 && logits.is_contiguous(), "moe_topk_router: contiguous fp32 [T, E] logits");
  c10::cuda::CUDAGuard g(logits.device());
  const int T = (int)logits.size(0), E = (int)logits.size(1);
  auto probs = at::empty({T, topk}, logits.options());
  auto ids = at::empty({T, topk}, logits.options().dtype(at::kLong));
  auto map = at::zeros({T, E}, logits.options().dtype(at::kBool));
  auto tpe = at::zeros({E}, logits.options().dtype(at::kInt));
  const int rc = mb200_moe_topk_router(logits.data_ptr<float>(), expert_bias.has_value() ? expert_bias->data_ptr<float>() : nullptr, T, E, (int)topk, (int)score_fn,
                                       renormalize ? 1 : 0, (float)scaling, probs.data_ptr<float>(), ids.data_ptr<int64_t>(), reinterpret_cast<uint8_t*>(map.data_ptr()),
                                       tpe.data_ptr<int>(), cur_stream());
  TORCH_CHECK(rc == 0, "moe_topk_router failed with code ", rc);
  return {probs, ids, map, tpe};
}
#endif

// Attention backward through the cuDNN library (consumes OUR forward's out + log-sum-exp); [s, b, h, d] tensors in and out.
std::vector<Tensor> attn_bwd_cudnn(const Tensor& go, const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& o, const Tensor& lse, bool causal, double scale) {
  c10::cuda::CUDAGuard g(q.device());
  auto P = [](const Tensor& t) { return t.permute({1, 2, 0, 3}); };
  auto seed = at::zeros({}, q.options().dtype(at::kLong));
  auto r = at::_scaled_dot_product_cudnn_attention_backward(P(go), P(q), P(k), P(v), P(o), lse, seed, seed, Tensor(), Tensor(), Tensor(), q.size(0), k.size(0), 0.0,
                                                            causal, scale);
  return {std::get<0>(r).permute({2, 0, 1, 3}), std::get<1>(r).permute({2, 0, 1, 3}), std::get<2>(r).permute({2, 0, 1, 3})};
}

#ifdef MB200_HAVE_FUSED_TP_GEMM
// mode 0: AG(a_shard) -> a_full (symmetric, filled in-kernel) ; c = a_full @ op(b).   mode 1: c (symmetric partial) = a @ op(b) ; rs_out = RS(c).
void fused_tp_gemm(int64_t mode, const Tensor& a, const Tensor& b, Tensor c, int64_t b_layout, int64_t rank, int64_t epoch, const Tensor& ag_src, int64_t ag_dst_mc,
                   const std::vector<int64_t>& ag_dst_peer, int64_t rs_src_mc, const std::vector<int64_t>& rs_src_peer, Tensor rs_out,
                   const Tensor& xag_src, int64_t xag_dst_mc, const std::vector<int64_t>& xag_dst_peer, const std::vector<int64_t>& flags_peer, Tensor counters,
                   int64_t comm_clusters) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kBFloat16, "fused_tp_gemm: bf16 CUDA tensors");
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && c.is_contiguous(), "fused_tp_gemm: contiguous operands");
  c10::cuda::CUDAGuard g(a.device());
  const int M = (int)a.size(0), K = (int)a.size(1);
  const int N = (int)(b_layout == 0 ? b.size(0) : b.size(1));
  TORCH_CHECK((b_layout == 0 ? b.size(1) : b.size(0)) == K, "fused_tp_gemm: K mismatch");
  const int world = (int)flags_peer.size();
  const bool have_xag = mode == 1 && xag_src.defined() && xag_src.numel() > 0 && !xag_dst_peer.empty();
  const int rc = mb200_fused_tp_gemm((int)mode, a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, (int)b_layout, (int)rank, world, (uint32_t)epoch,
                                     mode == 0 ? ag_src.data_ptr() : nullptr, ag_dst_mc, ag_dst_peer.empty() ? nullptr : ag_dst_peer.data(), rs_src_mc,
                                     rs_src_peer.empty() ? nullptr : rs_src_peer.data(), mode == 1 ? rs_out.data_ptr() : nullptr,
                                     have_xag ? xag_src.data_ptr() : nullptr, have_xag ? (int64_t)(xag_src.numel() * xag_src.element_size()) : 0, xag_dst_mc,
                                     xag_dst_peer.empty() ? nullptr : xag_dst_peer.data(), flags_peer.data(), counters.data_ptr(), (int)comm_clusters, cur_stream());
  TORCH_CHECK(rc == 0, "fused_tp_gemm failed with code ", rc);
}
#endif

// ---- second kernel batch (extra_kernels.cu + fused residual norm) ---------------------------------------------------------------
std::vector<Tensor> add_rmsnorm_fwd(const Tensor& x, const Tensor& res, const Tensor& w, double eps, bool zero_centered) {
  check_cuda_contig(x, "x"); check_cuda_contig(res, "residual"); check_cuda_contig(w, "weight");
  TORCH_CHECK(x.dim() == 2 && x.sizes() == res.sizes() && x.size(1) == w.numel() && x.scalar_type() == w.scalar_type() && x.scalar_type() == res.scalar_type());
  TORCH_CHECK(x.size(1) % vec_elems(x) == 0, "hidden size must be a multiple of ", vec_elems(x));
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x), h = at::empty_like(x);
  auto rstd = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  mb200_add_rmsnorm_fwd(x.data_ptr(), res.data_ptr(), w.data_ptr(), y.data_ptr(), h.data_ptr(), rstd.data_ptr<float>(), (int)x.size(0), (int)x.size(1), (float)eps,
                        zero_centered, dtype_code(x), cur_stream());
  return {y, h, rstd};
}

std::vector<Tensor> add_rmsnorm_bwd(const Tensor& gy, const c10::optional<Tensor>& gres, const Tensor& h, const Tensor& w, const Tensor& rstd, bool zero_centered) {
  check_cuda_contig(gy, "gy"); check_cuda_contig(h, "h");
  TORCH_CHECK(gy.sizes() == h.sizes() && gy.scalar_type() == h.scalar_type());
  if (gres.has_value()) { check_cuda_contig(*gres, "gres"); TORCH_CHECK(gres->sizes() == h.sizes() && gres->scalar_type() == h.scalar_type()); }
  c10::cuda::CUDAGuard g(h.device());
  const int rows = (int)h.size(0), H = (int)h.size(1);
  const int nblocks = std::min(rows, persistent_blocks());
  auto gx = at::empty_like(h);
  auto gw = at::empty_like(w);
  auto partial = at::empty({nblocks, 2, H}, h.options().dtype(at::kFloat));
  mb200_add_rmsnorm_bwd(gy.data_ptr(), gres.has_value() ? gres->data_ptr() : nullptr, h.data_ptr(), w.data_ptr(), rstd.data_ptr<float>(), gx.data_ptr(),
                        partial.data_ptr<float>(), gw.data_ptr(), rows, H, zero_centered, dtype_code(h), nblocks, cur_stream());
  return {gx, gw};
}

#ifdef MB200_HAVE_EXTRA_KERNELS
// t [T, H, D] packed tokens, pos int32 [T]
Tensor rope_pos(const Tensor& t, const Tensor& freqs, const Tensor& pos, double mscale, bool conj) {
  check_cuda_contig(t, "t"); check_cuda_contig(freqs, "freqs"); check_cuda_contig(pos, "pos");
  TORCH_CHECK(t.dim() == 3 && freqs.dim() == 2 && freqs.scalar_type() == at::kFloat && pos.scalar_type() == at::kInt && pos.numel() == t.size(0));
  const int Hh = (int)t.size(1), D = (int)t.size(2), Drot = (int)freqs.size(1), vn = vec_elems(t);
  TORCH_CHECK(Drot <= D && (Drot / 2) % vn == 0 && (D - Drot) % vn == 0, "rope: head dim / rotary dim not vectorisable");
  c10::cuda::CUDAGuard g(t.device());
  auto out = at::empty_like(t);
  mb200_rope_pos(t.data_ptr(), freqs.data_ptr<float>(), pos.data_ptr<int>(), out.data_ptr(), t.size(0), Hh, D, Drot, (float)mscale, conj, dtype_code(t), cur_stream());
  return out;
}

// qkv [S, B, NG, (QPG+2)*D]; rotates q and k heads; in place when `inplace`
Tensor rope_qkv(Tensor qkv, const Tensor& freqs, int64_t qpg, int64_t D, double mscale, bool conj, bool inplace) {
  check_cuda_contig(qkv, "qkv"); check_cuda_contig(freqs, "freqs");
  TORCH_CHECK(qkv.dim() == 4 && qkv.size(3) == (qpg + 2) * D && freqs.dim() == 2 && freqs.scalar_type() == at::kFloat && freqs.size(0) >= qkv.size(0));
  const int Drot = (int)freqs.size(1), vn = vec_elems(qkv);
  TORCH_CHECK(Drot <= D && (Drot / 2) % vn == 0 && (D - Drot) % vn == 0, "rope: head dim / rotary dim not vectorisable");
  c10::cuda::CUDAGuard g(qkv.device());
  Tensor out = inplace ? qkv : at::empty_like(qkv);
  mb200_rope_qkv(qkv.data_ptr(), freqs.data_ptr<float>(), out.data_ptr(), (int)qkv.size(0), (int)qkv.size(1), (int)qkv.size(2), (int)qpg, (int)D, Drot, (float)mscale,
                 conj, dtype_code(qkv), cur_stream());
  return out;
}

// x [b, d, l]; w [d, K]; left [b, d, K-1]
Tensor conv1d_fwd(const Tensor& x, const Tensor& w, const c10::optional<Tensor>& bias, const c10::optional<Tensor>& left, bool silu) {
  check_cuda_contig(x, "x"); check_cuda_contig(w, "w");
  TORCH_CHECK(x.dim() == 3 && w.dim() == 2 && w.size(0) == x.size(1) && w.scalar_type() == x.scalar_type());
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x);
  const int rc = mb200_conv1d_fwd(x.data_ptr(), w.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, left.has_value() ? left->data_ptr() : nullptr, y.data_ptr(),
                                  x.size(0) * x.size(1), (int)x.size(1), (int)x.size(2), (int)w.size(1), silu, dtype_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "conv1d_fwd: kernel width must be 2..4");
  return y;
}

std::vector<Tensor> conv1d_bwd(const Tensor& gy, const Tensor& x, const Tensor& w, const c10::optional<Tensor>& bias, const c10::optional<Tensor>& left, bool silu) {
  check_cuda_contig(gy, "gy"); check_cuda_contig(x, "x");
  c10::cuda::CUDAGuard g(x.device());
  auto gx = at::empty_like(x);
  auto gw = at::zeros({w.size(0), w.size(1)}, x.options().dtype(at::kFloat));
  auto gb = at::zeros({w.size(0)}, x.options().dtype(at::kFloat));
  Tensor gleft = left.has_value() ? at::empty_like(*left) : Tensor();
  const int rc = mb200_conv1d_bwd(gy.data_ptr(), x.data_ptr(), w.data_ptr(), bias.has_value() ? bias->data_ptr() : nullptr, left.has_value() ? left->data_ptr() : nullptr,
                                  gx.data_ptr(), left.has_value() ? gleft.data_ptr() : nullptr, gw.data_ptr<float>(), gb.data_ptr<float>(), x.size(0) * x.size(1),
                                  (int)x.size(1), (int)x.size(2), (int)w.size(1), silu, dtype_code(x), cur_stream());
  TORCH_CHECK(rc == 0, "conv1d_bwd: kernel width must be 2..4");
  return {gx, gw, gb, gleft.defined() ? gleft : at::empty({0}, x.options())};
}

// states [b, c, h, p, n] fp32; decay [b, h, c] fp32; init [b, h, p, n] fp32 (optional) -> prev [b, c, h, p, n], final [b, h, p, n]
std::vector<Tensor> ssd_state_fwd(const Tensor& states, const Tensor& decay, const c10::optional<Tensor>& init) {
  check_cuda_contig(states, "states"); check_cuda_contig(decay, "decay");
  TORCH_CHECK(states.dim() == 5 && states.scalar_type() == at::kFloat && decay.scalar_type() == at::kFloat);
  const int b = (int)states.size(0), c = (int)states.size(1), h = (int)states.size(2), E = (int)(states.size(3) * states.size(4));
  c10::cuda::CUDAGuard g(states.device());
  auto prev = at::empty_like(states);
  auto fin = at::empty({b, h, states.size(3), states.size(4)}, states.options());
  mb200_ssd_state_fwd(states.data_ptr<float>(), decay.data_ptr<float>(), init.has_value() ? init->data_ptr<float>() : nullptr, prev.data_ptr<float>(), fin.data_ptr<float>(),
                      b, c, h, E, cur_stream());
  return {prev, fin};
}

std::vector<Tensor> ssd_state_bwd(const Tensor& g_prev, const c10::optional<Tensor>& g_fin, const Tensor& prev, const Tensor& decay) {
  check_cuda_contig(g_prev, "g_prev"); check_cuda_contig(prev, "prev");
  const int b = (int)prev.size(0), c = (int)prev.size(1), h = (int)prev.size(2), E = (int)(prev.size(3) * prev.size(4));
  c10::cuda::CUDAGuard g(prev.device());
  auto g_states = at::empty_like(prev);
  auto g_init = at::empty({b, h, prev.size(3), prev.size(4)}, prev.options());
  auto g_decay = at::empty({b, h, c}, prev.options());
  mb200_ssd_state_bwd(g_prev.data_ptr<float>(), g_fin.has_value() ? g_fin->data_ptr<float>() : nullptr, prev.data_ptr<float>(), decay.data_ptr<float>(),
                      g_states.data_ptr<float>(), g_init.data_ptr<float>(), g_decay.data_ptr<float>(), b, c, h, E, cur_stream());
  return {g_states, g_init, g_decay};
}

// state [b, h, p, n] fp32 (updated in place); x [b, h, p]; dt [b, h] fp32; A [h] fp32; B, C [b, g, n]; D [h] fp32 optional
Tensor ssd_step(Tensor state, const Tensor& x, const Tensor& dt, const Tensor& A, const Tensor& B, const Tensor& C, const c10::optional<Tensor>& D) {
  check_cuda_contig(state, "state"); check_cuda_contig(x, "x"); check_cuda_contig(B, "B"); check_cuda_contig(C, "C");
  TORCH_CHECK(state.scalar_type() == at::kFloat && dt.scalar_type() == at::kFloat && A.scalar_type() == at::kFloat && dt.is_contiguous() && A.is_contiguous());
  TORCH_CHECK(B.scalar_type() == x.scalar_type() && C.scalar_type() == x.scalar_type());
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x);
  mb200_ssd_step(state.data_ptr<float>(), x.data_ptr(), dt.data_ptr<float>(), A.data_ptr<float>(), B.data_ptr(), C.data_ptr(), D.has_value() ? D->data_ptr<float>() : nullptr,
                 y.data_ptr(), (int)x.size(0), (int)x.size(1), (int)B.size(1), (int)x.size(2), (int)B.size(2), dtype_code(x), cur_stream());
  return y;
}

// x [rows, K] bf16 -> (q uint8 [rows, K] (e4m3 bits), sf uint8 [rows, K/32] (e8m0))
std::vector<Tensor> mxfp8_quant(const Tensor& x) {
  check_cuda_contig(x, "x");
  TORCH_CHECK(x.dim() == 2 && x.scalar_type() == at::kBFloat16 && x.size(1) % 32 == 0, "mxfp8_quant: bf16 [rows, K], K % 32 == 0");
  c10::cuda::CUDAGuard g(x.device());
  auto q = at::empty({x.size(0), x.size(1)}, x.options().dtype(at::kByte));
  auto sf = at::empty({x.size(0), x.size(1) / 32}, x.options().dtype(at::kByte));
  mb200_mxfp8_quant(x.data_ptr(), q.data_ptr(), sf.data_ptr(), x.size(0), (int)x.size(1), cur_stream());
  return {q, sf};
}

// x [rows, K] bf16, tscale fp32 device scalar -> (packed e2m1 uint8 [rows, K/2], ue4m3 block scales uint8 [rows, K/16])
std::vector<Tensor> nvfp4_quant(const Tensor& x, const Tensor& tscale) {
  check_cuda_contig(x, "x");
  TORCH_CHECK(x.dim() == 2 && x.scalar_type() == at::kBFloat16 && x.size(1) % 16 == 0, "nvfp4_quant: bf16 [rows, K], K % 16 == 0");
  TORCH_CHECK(tscale.is_cuda() && tscale.scalar_type() == at::kFloat && tscale.numel() == 1, "nvfp4_quant: tscale must be a 1-element fp32 CUDA tensor");
  c10::cuda::CUDAGuard g(x.device());
  auto q = at::empty({x.size(0), x.size(1) / 2}, x.options().dtype(at::kByte));
  auto sf = at::empty({x.size(0), x.size(1) / 16}, x.options().dtype(at::kByte));
  mb200_nvfp4_quant(x.data_ptr(), tscale.data_ptr<float>(), q.data_ptr(), sf.data_ptr(), x.size(0), (int)x.size(1), cur_stream());
  return {q, sf};
}

// x [b, h, sq, sk]; mask uint8 [b, 1, sq, sk] (optional)
Tensor softmax_fwd(const Tensor& x, const c10::optional<Tensor>& mask, double scale, bool causal) {
  check_cuda_contig(x, "x");
  TORCH_CHECK(x.dim() == 4, "softmax_fwd: [b, h, sq, sk]");
  if (mask.has_value()) { check_cuda_contig(*mask, "mask"); TORCH_CHECK(mask->scalar_type() == at::kByte && mask->numel() == x.size(0) * x.size(2) * x.size(3), "softmax_fwd: mask must be uint8 [b, 1, sq, sk]"); }
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty_like(x);
  mb200_softmax_fwd(x.data_ptr(), mask.has_value() ? mask->data_ptr() : nullptr, y.data_ptr(), x.size(0) * x.size(1) * x.size(2), (int)x.size(1), (int)x.size(2), (int)x.size(3),
                    (float)scale, causal, dtype_code(x), cur_stream());
  return y;
}
Tensor softmax_bwd(const Tensor& gy, const Tensor& y, double scale) {
  check_cuda_contig(gy, "gy"); check_cuda_contig(y, "y");
  c10::cuda::CUDAGuard g(y.device());
  auto gx = at::empty_like(y);
  mb200_softmax_bwd(gy.data_ptr(), y.data_ptr(), gx.data_ptr(), y.numel() / y.size(-1), (int)y.size(-1), (float)scale, dtype_code(y), cur_stream());
  return gx;
}
// mode 0: squared ReLU on [rows, F]; mode 1: quick-GeGLU, x [rows, 2F] -> [rows, F]
Tensor act_fwd(const Tensor& x, int64_t mode) {
  check_cuda_contig(x, "x");
  TORCH_CHECK(x.dim() == 2, "act_fwd: 2-D input");
  const int64_t F = mode == 0 ? x.size(1) : x.size(1) / 2;
  TORCH_CHECK(F % vec_elems(x) == 0, "act_fwd: width must be a multiple of ", vec_elems(x));
  c10::cuda::CUDAGuard g(x.device());
  auto y = at::empty({x.size(0), F}, x.options());
  mb200_act_fwd(x.data_ptr(), y.data_ptr(), x.size(0), (int)F, (int)mode, dtype_code(x), cur_stream());
  return y;
}
Tensor act_bwd(const Tensor& g_, const Tensor& x, int64_t mode) {
  check_cuda_contig(g_, "g"); check_cuda_contig(x, "x");
  const int64_t F = mode == 0 ? x.size(1) : x.size(1) / 2;
  c10::cuda::CUDAGuard g(x.device());
  auto gx = at::empty_like(x);
  mb200_act_bwd(g_.data_ptr(), x.data_ptr(), gx.data_ptr(), x.size(0), (int)F, (int)mode, dtype_code(x), cur_stream());
  return gx;
}

Tensor mxfp8_dequant(const Tensor& q, const Tensor& sf) {
  check_cuda_contig(q, "q"); check_cuda_contig(sf, "sf");
  TORCH_CHECK(q.dim() == 2 && q.scalar_type() == at::kByte && sf.scalar_type() == at::kByte && sf.size(0) == q.size(0) && sf.size(1) * 32 == q.size(1));
  c10::cuda::CUDAGuard g(q.device());
  auto out = at::empty({q.size(0), q.size(1)}, q.options().dtype(at::kBFloat16));
  mb200_mxfp8_dequant(q.data_ptr(), sf.data_ptr(), out.data_ptr(), q.size(0), (int)q.size(1), cur_stream());
  return out;
}
#endif

#ifdef MB200_HAVE_GEMM_MXFP8_SM100
// a [M,K], b [N,K] uint8 (e4m3 bits); sfa / sfb: swizzled E8M0 scale atoms [ceil(rows/128), K/128, 512] -> bf16 [M,N]
Tensor gemm_mxfp8_nt(const Tensor& a, const Tensor& sfa, const Tensor& b, const Tensor& sfb, int64_t tile) {
  TORCH_CHECK(a.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.is_contiguous() && b.is_contiguous() && a.size(1) == b.size(1) && a.scalar_type() == at::kByte &&
                  b.scalar_type() == at::kByte, "gemm_mxfp8_nt: a [M,K], b [N,K] contiguous uint8");
  const int64_t M = a.size(0), N = b.size(0), K = a.size(1);
  TORCH_CHECK(K % 128 == 0, "gemm_mxfp8_nt: K must be a multiple of 128");
  TORCH_CHECK(sfa.is_contiguous() && sfb.is_contiguous() && sfa.scalar_type() == at::kByte && sfb.scalar_type() == at::kByte &&
                  sfa.numel() == ((M + 127) / 128) * (K / 128) * 512 && sfb.numel() == ((N + 127) / 128) * (K / 128) * 512, "gemm_mxfp8_nt: scale atoms have the wrong size");
  c10::cuda::CUDAGuard g(a.device());
  auto c = at::empty({M, N}, a.options().dtype(at::kBFloat16));
  const int rc = mb200_gemm_mxfp8_nt(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), c.data_ptr(), (int)M, (int)N, (int)K, (int)tile, cur_stream());
  TORCH_CHECK(rc == 0, "gemm_mxfp8_nt failed with code ", rc);
  return c;
}
#endif

#ifdef MB200_HAVE_GEMM_NVFP4_SM100
// a [M,K/2], b [N,K/2] uint8 (two e2m1 codes per byte); sfa / sfb: swizzled UE4M3 scale atoms [ceil(rows/128), K/64, 512]; alpha_dev: product of the tensor scales
Tensor gemm_nvfp4_nt(const Tensor& a, const Tensor& sfa, const Tensor& b, const Tensor& sfb, double alpha, const c10::optional<Tensor>& alpha_dev) {
  TORCH_CHECK(a.is_cuda() && a.dim() == 2 && b.dim() == 2 && a.is_contiguous() && b.is_contiguous() && a.size(1) == b.size(1) && a.scalar_type() == at::kByte &&
                  b.scalar_type() == at::kByte, "gemm_nvfp4_nt: a [M,K/2], b [N,K/2] contiguous uint8");
  const int64_t M = a.size(0), N = b.size(0), K = a.size(1) * 2;
  TORCH_CHECK(K % 256 == 0, "gemm_nvfp4_nt: K must be a multiple of 256");
  TORCH_CHECK(sfa.is_contiguous() && sfb.is_contiguous() && sfa.scalar_type() == at::kByte && sfb.scalar_type() == at::kByte &&
                  sfa.numel() == ((M + 127) / 128) * (K / 64) * 512 && sfb.numel() == ((N + 127) / 128) * (K / 64) * 512, "gemm_nvfp4_nt: scale atoms have the wrong size");
  c10::cuda::CUDAGuard g(a.device());
  auto c = at::empty({M, N}, a.options().dtype(at::kBFloat16));
  const int rc = mb200_gemm_nvfp4_nt(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), c.data_ptr(), (int)M, (int)N, (int)K, (float)alpha,
                                     alpha_dev.has_value() ? alpha_dev->data_ptr<float>() : nullptr, cur_stream());
  TORCH_CHECK(rc == 0, "gemm_nvfp4_nt failed with code ", rc);
  return c;
}
#endif

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
#ifdef MB200_HAVE_GEMM_NVFP4_SM100
  m.def("gemm_nvfp4_nt", &gemm_nvfp4_nt);
#endif
#ifdef MB200_HAVE_GEMM_MXFP8_SM100
  m.def("gemm_mxfp8_nt", &gemm_mxfp8_nt, pybind11::arg("a"), pybind11::arg("sfa"), pybind11::arg("b"), pybind11::arg("sfb"), pybind11::arg("tile") = 0);
#endif
  m.def("add_rmsnorm_fwd", &add_rmsnorm_fwd);
  m.def("add_rmsnorm_bwd", &add_rmsnorm_bwd);
#ifdef MB200_HAVE_EXTRA_KERNELS
  m.def("rope_pos", &rope_pos);
  m.def("rope_qkv", &rope_qkv);
  m.def("conv1d_fwd", &conv1d_fwd);
  m.def("conv1d_bwd", &conv1d_bwd);
  m.def("ssd_state_fwd", &ssd_state_fwd);
  m.def("ssd_state_bwd", &ssd_state_bwd);
  m.def("ssd_step", &ssd_step);
  m.def("mxfp8_quant", &mxfp8_quant);
  m.def("mxfp8_dequant", &mxfp8_dequant);
  m.def("nvfp4_quant", &nvfp4_quant);
  m.def("softmax_fwd", &softmax_fwd);
  m.def("softmax_bwd", &softmax_bwd);
  m.def("act_fwd", &act_fwd);
  m.def("act_bwd", &act_bwd);
#endif
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("layernorm_fwd", &layernorm_fwd);
  m.def("layernorm_bwd", &layernorm_bwd);
  m.def("swiglu_fwd", &swiglu_fwd);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("rope_fwd", &rope_fwd);
  m.def("ce_stats", &ce_stats);
  m.def("ce_bwd", &ce_bwd);
  m.def("multi_l2norm", &multi_l2norm);
  m.def("multi_scale", &multi_scale);
  m.def("multi_adam", &multi_adam);
  m.def("gemm_bf16", &gemm_bf16, pybind11::arg("a"), pybind11::arg("b"), pybind11::arg("c"), pybind11::arg("layout"), pybind11::arg("accumulate"),
        pybind11::arg("variant") = 0);
#ifdef MB200_HAVE_NVLINK_COLLECTIVES
  m.def("nvl_barrier", &nvl_barrier);
  m.def("nvl_allgather", &nvl_allgather);
  m.def("nvl_reducescatter", &nvl_reducescatter);
  m.def("nvl_allreduce", &nvl_allreduce);
#endif
  m.def("attn_bwd_cudnn", &attn_bwd_cudnn);
  m.def("share_storage", &share_storage);
#ifdef MB200_HAVE_FLASH_ATTN_BWD_SM100
  m.def("flash_attn_bwd", &flash_attn_bwd, pybind11::arg("go"), pybind11::arg("q"), pybind11::arg("k"), pybind11::arg("v"), pybind11::arg("o"), pybind11::arg("lse"),
        pybind11::arg("causal"), pybind11::arg("scale"), pybind11::arg("split_heads") = -1, pybind11::arg("max_scratch_mb") = 0, pybind11::arg("row_lo") = pybind11::none(),
        pybind11::arg("col_hi") = pybind11::none());
#endif
#ifdef MB200_HAVE_ROUTING_KERNELS
  m.def("indices_to_multihot", &indices_to_multihot);
  m.def("multihot_probs_grad", &multihot_probs_grad);
  m.def("multihot_to_indices", &multihot_to_indices);
  m.def("pad_routing_map", &pad_routing_map);
  m.def("moe_aux_loss_fwd", &moe_aux_loss_fwd);
  m.def("moe_aux_loss_bwd", &moe_aux_loss_bwd);
  m.def("mla_rope_inplace", &mla_rope_inplace);
  m.def("mla_kv_split", &mla_kv_split);
#endif
#ifdef MB200_HAVE_MISC_KERNELS
  m.def("paged_stash", &paged_stash);
  m.def("spec_verify", &spec_verify);
  m.def("bias_dropout_add_fwd", &bias_dropout_add_fwd);
  m.def("bias_dropout_add_bwd", &bias_dropout_add_bwd);
#endif
#ifdef MB200_HAVE_PAGED_ATTENTION
  m.def("paged_kv_append", &paged_kv_append);
  m.def("paged_decode", &paged_decode, pybind11::arg("q"), pybind11::arg("k_pool"), pybind11::arg("v_pool"), pybind11::arg("block_table"), pybind11::arg("lengths"),
        pybind11::arg("scale"), pybind11::arg("max_len"), pybind11::arg("fixed_tokens_per_split") = 0);
#endif
#ifdef MB200_HAVE_MOE_KERNELS
  m.def("moe_gather_rows", &moe_gather_rows);
  m.def("moe_combine_rows", &moe_combine_rows);
  m.def("moe_topk_router", &moe_topk_router);
  m.def("moe_push_rows", &moe_push_rows);
  m.def("moe_pull_rows", &moe_pull_rows);
#endif
#ifdef MB200_HAVE_GEMM_FP8_SM100
  m.def("gemm_fp8_nt", &gemm_fp8_nt);
#endif
#ifdef MB200_HAVE_GROUPED_GEMM_SM100
  m.def("grouped_gemm_bf16", &grouped_gemm_bf16);
#endif
#ifdef MB200_HAVE_RUNTIME_NATIVE
  m.def("batched_copy", &batched_copy);
#endif
#ifdef MB200_HAVE_FLASH_ATTN_SM100
  m.def("flash_attn_fwd", &flash_attn_fwd, pybind11::arg("q"), pybind11::arg("k"), pybind11::arg("v"), pybind11::arg("causal"), pybind11::arg("scale"),
        pybind11::arg("variant") = 0, pybind11::arg("row_lo") = pybind11::none());
#endif
#ifdef MB200_HAVE_FUSED_TP_GEMM
  m.def("fused_tp_gemm", &fused_tp_gemm);
#endif
}
