"""Python surface of the second kernel batch (``csrc/extra_kernels.cu`` + the residual-fused norm in ``csrc/norm.cu``).

Every op has a PyTorch fp32 reference path (CPU tensors, and the oracle of the GPU tests) and an autograd wrapper around the CUDA
kernels.  Like the rest of ``ops``, a CUDA tensor with a missing extension raises instead of silently falling back."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _count, _use_cuda, ext
from . import reference as ref


# ------------------------------------------------------------------------------------------------ residual add + RMSNorm
class _AddRMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, w, eps, zero_centered):
        shape = x.shape
        x2, r2 = x.reshape(-1, shape[-1]), residual.reshape(-1, shape[-1])
        if _use_cuda(x):
            y, h, rstd = ext().add_rmsnorm_fwd(x2.contiguous(), r2.contiguous(), w, eps, zero_centered)
            _count()
        else:
            h = (x2.float() + r2.float()).to(x.dtype)
            y, rstd = ref.rms_norm_fwd(h, w, eps, zero_centered)
        ctx.save_for_backward(h, w, rstd)
        ctx.zero_centered, ctx.shape = zero_centered, shape
        return y.view(shape), h.view(shape)

    @staticmethod
    def backward(ctx, gy, gh):
        h, w, rstd = ctx.saved_tensors
        g2 = gy.reshape(h.shape)
        gh2 = gh.reshape(h.shape) if gh is not None else None
        if _use_cuda(g2):
            gx, gw = ext().add_rmsnorm_bwd(g2.contiguous(), gh2.contiguous() if gh2 is not None else None, h, w, rstd, ctx.zero_centered)
            _count(2)
        else:
            gx, gw = ref.rms_norm_bwd(g2, h, w, rstd, ctx.zero_centered)
            if gh2 is not None:
                gx = (gx.float() + gh2.float()).to(gx.dtype)
        gx = gx.view(ctx.shape)
        return gx, gx, gw, None, None


def add_rms_norm(x, residual, weight, eps: float = 1e-5, zero_centered_gamma: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``h = x + residual``; returns ``(rmsnorm(h) * weight, h)`` in one pass over the data (and one pass in backward, where the gradient
    arriving through ``h`` is added inside the norm-backward kernel)."""
    return _AddRMSNormFn.apply(x, residual, weight, eps, zero_centered_gamma)


# ------------------------------------------------------------------------------------------------ RoPE variants
def positions_from_cu_seqlens(cu_seqlens: torch.Tensor, total: int) -> torch.Tensor:
    """int32 position of every packed token inside its own sequence (vectorised: no host loop over sequences)."""
    cu = cu_seqlens.to(torch.int64)
    tok = torch.arange(total, device=cu.device)
    seq = torch.searchsorted(cu[1:].contiguous(), tok, right=True).clamp(max=cu.numel() - 2)
    return (tok - cu[seq]).to(torch.int32)


class _RoPEThdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, freqs, pos, mscale):
        ctx.save_for_backward(freqs, pos)
        ctx.mscale = mscale
        return _rope_pos(t, freqs, pos, mscale, False)

    @staticmethod
    def backward(ctx, g):
        freqs, pos = ctx.saved_tensors
        return _rope_pos(g, freqs, pos, ctx.mscale, True), None, None, None


def _rope_pos(t, freqs, pos, mscale, conj):
    f2 = freqs.reshape(freqs.shape[0], -1).float().contiguous()
    if _use_cuda(t):
        out = ext().rope_pos(t.contiguous(), f2, pos.contiguous(), float(mscale), conj)
        _count()
        return out
    fr = f2[pos.long()][:, None, :]                        # [T, 1, d_rot]
    return ref.rope_fwd(t.unsqueeze(1), fr.unsqueeze(1), False, mscale, conj=conj).squeeze(1)


def apply_rope_thd(t: torch.Tensor, cu_seqlens: torch.Tensor, freqs: torch.Tensor, mscale: float = 1.0) -> torch.Tensor:
    """Rotary embedding for packed sequences: ``t [T, h, d]``, every sequence restarts at position 0 (reference: fused RoPE ``thd`` format)."""
    pos = positions_from_cu_seqlens(cu_seqlens, t.shape[0])
    return _RoPEThdFn.apply(t, freqs, pos, mscale)


class _RoPEQKVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, freqs, qpg, d, mscale):
        ctx.save_for_backward(freqs)
        ctx.qpg, ctx.d, ctx.mscale = qpg, d, mscale
        return _rope_qkv(qkv, freqs, qpg, d, mscale, False)

    @staticmethod
    def backward(ctx, g):
        (freqs,) = ctx.saved_tensors
        return _rope_qkv(g, freqs, ctx.qpg, ctx.d, ctx.mscale, True), None, None, None, None


def _rope_qkv(qkv, freqs, qpg, d, mscale, conj):
    f2 = freqs.reshape(freqs.shape[0], -1).float().contiguous()
    if _use_cuda(qkv):
        out = ext().rope_qkv(qkv.contiguous(), f2, qpg, d, float(mscale), conj, False)
        _count()
        return out
    s, b, ng, _ = qkv.shape
    x = qkv.reshape(s, b, ng, qpg + 2, d)
    rot = ref.rope_fwd(x[:, :, :, : qpg + 1].reshape(s, b, ng * (qpg + 1), d), f2[:s, None, None, :], False, mscale, conj=conj).reshape(s, b, ng, qpg + 1, d)
    return torch.cat([rot, x[:, :, :, qpg + 1 :]], dim=3).reshape(qkv.shape)


def apply_rope_qkv(mixed_qkv: torch.Tensor, freqs: torch.Tensor, queries_per_group: int, head_dim: int, mscale: float = 1.0) -> torch.Tensor:
    """RoPE on the query and key heads of the mixed QKV projection output ``[s, b, ng, (qpg + 2) * d]`` in ONE kernel; the value heads pass
    through.  q / k / v are then strided views of the result (the flash-attention kernels take strided operands), so the three split copies
    of the unfused path disappear."""
    return _RoPEQKVFn.apply(mixed_qkv, freqs, queries_per_group, head_dim, mscale)


# ------------------------------------------------------------------------------------------------ causal conv1d
def _conv1d_ref(x, w, bias, left, silu):
    d, k = w.shape
    lf = left if left is not None else x.new_zeros(x.shape[0], d, k - 1)
    xp = torch.cat([lf.to(x.dtype), x], dim=-1)
    y = torch.nn.functional.conv1d(xp.float(), w.float().unsqueeze(1), bias.float() if bias is not None else None, groups=d)
    return (torch.nn.functional.silu(y) if silu else y).to(x.dtype)


class _CausalConv1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, left, silu):
        x = x.contiguous()
        ctx.save_for_backward(x, w, bias, left)
        ctx.silu = silu
        y = ext().conv1d_fwd(x, w.contiguous(), bias, left.contiguous() if left is not None else None, silu)
        _count()
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, bias, left = ctx.saved_tensors
        gx, gw, gb, gleft = ext().conv1d_bwd(gy.contiguous(), x, w.contiguous(), bias, left.contiguous() if left is not None else None, ctx.silu)
        _count()
        return gx, gw.to(w.dtype), (gb.to(bias.dtype) if bias is not None else None), (gleft if left is not None else None), None


def causal_conv1d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, initial_state: Optional[torch.Tensor] = None, silu: bool = True):
    """Depthwise causal convolution ``x [b, d, l]``, ``weight [d, k]`` (k ≤ 4), optional carried-in window ``[b, d, k-1]``, fused SiLU."""
    if _use_cuda(x) and 2 <= weight.shape[1] <= 4 and weight.dtype == x.dtype:
        return _CausalConv1dFn.apply(x, weight, bias, initial_state, silu)
    return _conv1d_ref(x, weight, bias, initial_state, silu)


# ------------------------------------------------------------------------------------------------ SSD state passing / step
def _state_passing_ref(states, decay, init):
    b, c, h, p, n = states.shape
    s = init.float() if init is not None else states.new_zeros(b, h, p, n)
    prev = []
    for z in range(c):
        prev.append(s)
        s = torch.exp(decay[:, :, z]).view(b, h, 1, 1) * s + states[:, z]
    return torch.stack(prev, dim=1), s


class _StatePassingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, states, decay, init):
        states, decay = states.contiguous(), decay.contiguous()
        prev, fin = ext().ssd_state_fwd(states, decay, init.contiguous() if init is not None else None)
        _count()
        ctx.save_for_backward(prev, decay)
        ctx.has_init = init is not None
        return prev, fin

    @staticmethod
    def backward(ctx, g_prev, g_fin):
        prev, decay = ctx.saved_tensors
        g_states, g_init, g_decay = ext().ssd_state_bwd(g_prev.contiguous(), g_fin.contiguous() if g_fin is not None else None, prev, decay)
        _count()
        return g_states, g_decay, (g_init if ctx.has_init else None)


def ssd_state_passing(states: torch.Tensor, chunk_log_decay: torch.Tensor, initial_state: Optional[torch.Tensor] = None):
    """Inter-chunk recurrence of the Mamba-2 SSD scan: ``S_{z+1} = exp(decay_z) · S_z + states_z``.
    ``states [b, c, h, p, n]`` fp32, ``chunk_log_decay [b, h, c]`` → (state entering each chunk ``[b, c, h, p, n]``, final state ``[b, h, p, n]``).
    O(c) sequential scan in one kernel instead of the O(c²) segment-sum matmul."""
    if _use_cuda(states):
        return _StatePassingFn.apply(states.float(), chunk_log_decay.float(), initial_state.float() if initial_state is not None else None)
    return _state_passing_ref(states.float(), chunk_log_decay.float(), initial_state)


def ssd_step(state: torch.Tensor, x, dt, A, B, C, D=None):
    """Decode-time state update, in place on ``state [b, h, p, n]`` (fp32).  Returns ``y [b, h, p]``."""
    if _use_cuda(x) and state.dtype == torch.float32:
        y = ext().ssd_step(state, x.contiguous(), dt.float().contiguous(), A.float().contiguous(), B.to(x.dtype).contiguous(), C.to(x.dtype).contiguous(),
                           D.float().contiguous() if D is not None else None)
        _count()
        return y
    h, g = x.shape[1], B.shape[1]
    Bf, Cf = B.float().repeat_interleave(h // g, dim=1), C.float().repeat_interleave(h // g, dim=1)
    dA = torch.exp(A.float().view(1, h) * dt.float())
    state.mul_(dA[..., None, None]).add_(torch.einsum("bh,bhn,bhp->bhpn", dt.float(), Bf, x.float()))
    y = torch.einsum("bhpn,bhn->bhp", state.float(), Cf)
    if D is not None:
        y = y + x.float() * D.float().view(1, h, 1)
    return y.to(x.dtype)


# ------------------------------------------------------------------------------------------------ MXFP8
E4M3_MAX_EXP = 8


def mxfp8_quantize_reference(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    rows, K = x.shape
    xb = x.float().view(rows, K // 32, 32)
    amax = xb.abs().amax(-1)
    e = torch.where(amax > 0, torch.floor(torch.log2(amax.clamp(min=1e-38))) - E4M3_MAX_EXP, torch.full_like(amax, -127.0)).clamp(-127, 127)
    q = (xb * torch.exp2(-e).unsqueeze(-1)).clamp(-448, 448).to(torch.float8_e4m3fn)
    return q.view(rows, K).view(torch.uint8), (e + 127).to(torch.uint8)


def mxfp8_quantize(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``x [rows, K]`` → (E4M3 payload as uint8 ``[rows, K]``, E8M0 block scales uint8 ``[rows, K/32]``) — OCP MXFP8, 1×32 blocks along K."""
    if _use_cuda(x):
        q, sf = ext().mxfp8_quant(x.to(torch.bfloat16).contiguous())
        _count()
        return q, sf
    return mxfp8_quantize_reference(x)


def mxfp8_dequantize(q: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    if _use_cuda(q):
        out = ext().mxfp8_dequant(q.contiguous(), sf.contiguous())
        _count()
        return out
    rows, K = q.shape
    v = q.view(torch.float8_e4m3fn).float().view(rows, K // 32, 32)
    return (v * torch.exp2(sf.float() - 127).unsqueeze(-1)).view(rows, K).to(torch.bfloat16)


def mxfp8_swizzle_scales(sf: torch.Tensor) -> torch.Tensor:
    """``sf [rows, K/32]`` (uint8 E8M0) → scale atoms ``[ceil(rows/128), K/128, 512]`` in the layout ``tcgen05.cp`` moves into tensor memory:
    inside an atom the byte of (row r, k-group k) sits at ``(r % 32) * 16 + (r // 32) * 4 + k``.  Rows are padded to a multiple of 128 with scale 0."""
    rows, kb = sf.shape
    assert kb % 4 == 0, "K must be a multiple of 128"
    R = (rows + 127) // 128
    if R * 128 != rows:
        sf = torch.cat([sf, sf.new_zeros(R * 128 - rows, kb)], dim=0)
    return sf.view(R, 4, 32, kb // 4, 4).permute(0, 3, 2, 1, 4).contiguous().view(R, kb // 4, 512)


def gemm_mxfp8_nt(a_q: torch.Tensor, a_sf: torch.Tensor, b_q: torch.Tensor, b_sf: torch.Tensor, tile: int = 0) -> torch.Tensor:
    """``C[M,N] (bf16) = dequant(A) · dequant(B)ᵀ`` with MXFP8 operands (payload uint8 ``[rows, K]`` + E8M0 scales ``[rows, K/32]``): block-scaled
    ``tcgen05.mma kind::mxf8f6f4`` on CUDA, dequantise-then-matmul reference elsewhere."""
    if _use_cuda(a_q):
        out = ext().gemm_mxfp8_nt(a_q.contiguous(), mxfp8_swizzle_scales(a_sf), b_q.contiguous(), mxfp8_swizzle_scales(b_sf), tile)
        _count()
        return out
    return (mxfp8_dequantize(a_q, a_sf).float() @ mxfp8_dequantize(b_q, b_sf).float().t()).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ NVFP4
def nvfp4_pack(codes: torch.Tensor) -> torch.Tensor:
    """One E2M1 code per byte ``[rows, K]`` → two per byte ``[rows, K/2]`` (element 2i in the low nibble)."""
    return (codes[:, 0::2] | (codes[:, 1::2] << 4)).contiguous()


def nvfp4_unpack(packed: torch.Tensor) -> torch.Tensor:
    out = torch.empty(packed.shape[0], packed.shape[1] * 2, dtype=torch.uint8, device=packed.device)
    out[:, 0::2], out[:, 1::2] = packed & 0xF, packed >> 4
    return out


def nvfp4_quantize(x: torch.Tensor):
    """``x [rows, K]`` → ``(packed codes uint8 [rows, K/2], block scales float8_e4m3fn [rows, K/16], tensor scale fp32 [1])``: the CUDA quantiser
    (one pass, no host sync: the tensor amax stays on the device) — same numerics as ``core.fp4_utils.quantize_nvfp4`` up to tie rounding."""
    tscale = (x.detach().abs().amax().float().clamp(min=1e-12) / (6.0 * 448.0)).reshape(1)
    if _use_cuda(x):
        q, sf = ext().nvfp4_quant(x.to(torch.bfloat16).contiguous(), tscale)
        _count()
        return q, sf.view(torch.float8_e4m3fn), tscale
    from ..core.fp4_utils import quantize_nvfp4

    codes, bscale, t = quantize_nvfp4(x)
    return nvfp4_pack(codes), bscale, t.reshape(1)


def gemm_nvfp4_nt(a_codes: torch.Tensor, a_bscale: torch.Tensor, a_tscale, b_codes: torch.Tensor, b_bscale: torch.Tensor, b_tscale) -> torch.Tensor:
    """``C[M,N] (bf16) = dequant(A) · dequant(B)ᵀ`` for NVFP4 operands (codes uint8 — one per byte ``[rows, K]`` as produced by ``core.fp4_utils.quantize_nvfp4`` or
    packed ``[rows, K/2]`` as produced by ``nvfp4_quantize`` — block scales ``float8_e4m3fn [rows, K/16]``, fp32 tensor scale): block-scaled
    ``tcgen05.mma kind::mxf4nvf4`` on CUDA (K % 256 == 0), dequantise-then-matmul elsewhere."""
    K = a_bscale.shape[1] * 16
    a_packed, b_packed = a_codes.shape[1] * 2 == K, b_codes.shape[1] * 2 == K
    if _use_cuda(a_codes) and K % 256 == 0 and hasattr(ext(), "gemm_nvfp4_nt"):
        alpha = (torch.as_tensor(a_tscale, device=a_codes.device).float() * torch.as_tensor(b_tscale, device=a_codes.device).float()).reshape(1)
        out = ext().gemm_nvfp4_nt(a_codes.contiguous() if a_packed else nvfp4_pack(a_codes), mxfp8_swizzle_scales(a_bscale.view(torch.uint8)),
                                  b_codes.contiguous() if b_packed else nvfp4_pack(b_codes), mxfp8_swizzle_scales(b_bscale.view(torch.uint8)), 1.0, alpha)
        _count()
        return out
    from ..core.fp4_utils import dequantize_nvfp4

    ac = nvfp4_unpack(a_codes) if a_packed else a_codes
    bc = nvfp4_unpack(b_codes) if b_packed else b_codes
    return (dequantize_nvfp4(ac, a_bscale, a_tscale, torch.float32) @ dequantize_nvfp4(bc, b_bscale, b_tscale, torch.float32).t()).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ scale-mask-softmax, squared ReLU, quick-GeGLU
class _ScaledMaskedSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask, scale, causal):
        y = ext().softmax_fwd(x.contiguous(), mask, float(scale), causal)
        _count()
        ctx.save_for_backward(y)
        ctx.scale = float(scale)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gx = ext().softmax_bwd(gy.contiguous(), y, ctx.scale)
        _count()
        return gx, None, None, None


def scaled_masked_softmax(x: torch.Tensor, mask: Optional[torch.Tensor], scale: float = 1.0, causal: bool = False) -> torch.Tensor:
    """``softmax(scale · x + mask)`` over the last dim of ``[b, h, sq, sk]`` (``mask`` bool ``[b, 1, sq, sk]``, True = masked; ``causal`` = bottom-right aligned
    triangle) — the unfused attention path's kernel (reference ``fused_softmax.py`` / the ``scaled_*_softmax_cuda`` extensions)."""
    if _use_cuda(x) and x.dim() == 4 and (mask is None or (mask.dim() == 4 and mask.shape[1] == 1 and mask.shape[0] == x.shape[0])):
        m8 = None if mask is None else mask.to(torch.uint8).expand(x.shape[0], 1, x.shape[2], x.shape[3]).contiguous()
        return _ScaledMaskedSoftmaxFn.apply(x, m8, scale, causal)
    return ref.scaled_masked_softmax(x, mask, scale, causal=causal)


class _ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, mode):
        ctx.save_for_backward(x2)
        ctx.mode = mode
        y = ext().act_fwd(x2, mode)
        _count()
        return y

    @staticmethod
    def backward(ctx, g):
        (x2,) = ctx.saved_tensors
        gx = ext().act_bwd(g.contiguous(), x2, ctx.mode)
        _count()
        return gx, None


def squared_relu(x: torch.Tensor) -> torch.Tensor:
    if _use_cuda(x) and x.shape[-1] % 8 == 0:
        return _ActFn.apply(x.reshape(-1, x.shape[-1]).contiguous(), 0).view(x.shape)
    return torch.pow(torch.relu(x), 2)


def quick_geglu(y: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[..., 2F] → [..., F]``: ``a · σ(1.702 a) · b`` with ``(a, b) = chunk(y + bias)``."""
    yb = y if bias is None else y + bias
    if _use_cuda(yb) and (yb.shape[-1] // 2) % 8 == 0:
        return _ActFn.apply(yb.reshape(-1, yb.shape[-1]).contiguous(), 1).view(*yb.shape[:-1], yb.shape[-1] // 2)
    a, b = yb.chunk(2, dim=-1)
    return a * torch.sigmoid(1.702 * a) * b
