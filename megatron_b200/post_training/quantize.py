"""Post-training quantisation (reference ``megatron/post_training/`` + ``core/post_training/modelopt`` — ModelOpt PTQ specs; implemented natively here).

``calibrate`` runs a few batches and records the activation absolute maximum at the input of every (matched) linear layer; ``quantize_model`` then turns those
layers into inference layers whose WEIGHTS are stored quantised (packed payload + scales — the tensors the block-scaled kernels consume) and whose
activations are quantised on the fly:

    format        weights                      activations                GEMM
    ------        -------                      -----------                ----
    fp8           E4M3, per-tensor scale       E4M3, calibrated scale     tcgen05 kind::f8f6f4
    mxfp8         E4M3 + E8M0 / 32             same, dynamic              tcgen05 kind::mxf8f6f4.block_scale
    nvfp4         E2M1 + UE4M3 / 16 + fp32     same, dynamic              tcgen05 kind::mxf4nvf4.block_scale
    w4a16         NVFP4 weights only           bf16                       dequantise + bf16 GEMM (memory-bound decode)

Per-layer choice goes through the same glob matchers as ``core/quantization`` (first match wins, ``none`` keeps a layer in bf16 — typically the output
head and the first / last blocks).  ``export_quantized_state_dict`` returns the packed tensors for deployment / checkpointing."""
from __future__ import annotations

import fnmatch
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

FORMATS = ("fp8", "mxfp8", "nvfp4", "w4a16", "none")


@dataclass
class PTQConfig:
    default: str = "mxfp8"
    matchers: List[Tuple[str, str]] = field(default_factory=lambda: [("*output_layer*", "none")])     # (glob over module path, format)

    def format_for(self, path: str) -> str:
        for pat, fmt in self.matchers:
            if fnmatch.fnmatch(path, pat):
                return fmt
        return self.default


def _linears(model):
    from ..core.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear

    for name, m in model.named_modules():
        if isinstance(m, (ColumnParallelLinear, RowParallelLinear, torch.nn.Linear)) and getattr(m, "weight", None) is not None and m.weight.dim() == 2:
            yield name, m


@torch.no_grad()
def calibrate(model, batches: Iterable, forward_fn: Callable, config: Optional[PTQConfig] = None) -> Dict[str, float]:
    """Max-calibration: ``amax[path]`` of the input activations of every linear that will be quantised."""
    config = config or PTQConfig()
    amax: Dict[str, float] = {}
    hs = []
    for name, m in _linears(model):
        if config.format_for(name) == "none":
            continue

        def obs(mod, args, _n=name):
            amax[_n] = max(amax.get(_n, 0.0), float(args[0].detach().abs().amax()))

        hs.append(m.register_forward_pre_hook(obs))
    was = model.training
    model.eval()
    try:
        for b in batches:
            forward_fn(model, b)
    finally:
        model.train(was)
        for h in hs:
            h.remove()
    return amax


class QuantizedLinearState:
    """What replaces ``weight`` after PTQ."""

    def __init__(self, fmt: str, weight: torch.Tensor, act_amax: Optional[float]):
        from .. import ops
        from ..core.fp4_utils import quantize_nvfp4
        from ..core.fp8_utils import E4M3, FP8_MAX

        self.fmt, self.out_features, self.in_features, self.dtype = fmt, weight.shape[0], weight.shape[1], weight.dtype
        w = weight.detach()
        if fmt == "fp8":
            s = FP8_MAX[E4M3] / w.abs().amax().float().clamp(min=1e-12)
            self.tensors = {"weight_q": (w.float() * s).clamp(-FP8_MAX[E4M3], FP8_MAX[E4M3]).to(E4M3), "weight_scale_inv": (1.0 / s).reshape(1)}
            a = FP8_MAX[E4M3] / max(act_amax or 1.0, 1e-12)
            self.tensors["act_scale"] = torch.tensor([a], dtype=torch.float32, device=w.device)
        elif fmt == "mxfp8":
            q, sf = ops.mxfp8_quantize(w.to(torch.bfloat16))
            self.tensors = {"weight_q": q, "weight_sf": sf}
        elif fmt in ("nvfp4", "w4a16"):
            codes, bs, ts = quantize_nvfp4(w)
            self.tensors = {"weight_q": ops.nvfp4_pack(codes), "weight_sf": bs, "weight_tscale": ts.reshape(1)}
        else:
            raise ValueError(fmt)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors.values())

    def matmul(self, x: torch.Tensor) -> torch.Tensor:
        """``x [..., K] @ Wᵀ`` → ``[..., N]`` in x's dtype."""
        from .. import ops
        from ..core.fp4_utils import dequantize_nvfp4
        from ..core.fp8_utils import E4M3, FP8_MAX, _gemm_nt

        t = self.tensors
        x2 = x.reshape(-1, x.shape[-1])
        if self.fmt == "fp8":
            xq = (x2.float() * t["act_scale"]).clamp(-FP8_MAX[E4M3], FP8_MAX[E4M3]).to(E4M3)        # static (calibrated) activation scale: no amax pass at run time
            y = _gemm_nt(xq, 1.0 / t["act_scale"], t["weight_q"], t["weight_scale_inv"])
        elif self.fmt == "mxfp8":
            K = x2.shape[1]
            pad = (-K) % 128
            xq, xsf = ops.mxfp8_quantize(torch.nn.functional.pad(x2, (0, pad)).to(torch.bfloat16) if pad else x2.to(torch.bfloat16))
            if pad:
                wq = torch.nn.functional.pad(t["weight_q"], (0, pad))
                wsf = torch.nn.functional.pad(t["weight_sf"], (0, pad // 32))
            else:
                wq, wsf = t["weight_q"], t["weight_sf"]
            y = ops.gemm_mxfp8_nt(xq, xsf, wq, wsf)
        elif self.fmt == "nvfp4" and x2.shape[1] % 256 == 0:
            y = ops.gemm_nvfp4_nt(*ops.nvfp4_quantize(x2), t["weight_q"], t["weight_sf"], t["weight_tscale"])
        else:       # w4a16 (and nvfp4 with an unaligned K): weight-only
            w = dequantize_nvfp4(ops.nvfp4_unpack(t["weight_q"]), t["weight_sf"], t["weight_tscale"], x.dtype)
            y = x2 @ w.t()
        return y.to(x.dtype).view(*x.shape[:-1], self.out_features)


def quantize_model(model, config: Optional[PTQConfig] = None, act_amax: Optional[Dict[str, float]] = None) -> Dict[str, QuantizedLinearState]:
    """In place: matched linears lose their bf16 ``weight`` Parameter (memory is released) and compute through ``QuantizedLinearState``.  TP collectives of
    Column / Row parallel layers are kept (the replacement forward calls the same mappings).  Returns ``{path: state}``."""
    from ..core.tensor_parallel import mappings as mp
    from ..core.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
    from ..core.utils import get_pg_size

    config = config or PTQConfig()
    states: Dict[str, QuantizedLinearState] = {}
    for name, m in list(_linears(model)):
        fmt = config.format_for(name)
        if fmt == "none":
            continue
        st = QuantizedLinearState(fmt, m.weight, (act_amax or {}).get(name))
        states[name] = st
        bias = getattr(m, "bias", None)
        del m._parameters["weight"]
        m.quant_state = st

        if isinstance(m, ColumnParallelLinear):
            def fwd(input_, weight=None, runtime_gather_output=None, _m=m, _st=st):
                x = mp.gather_from_sequence_parallel_region(input_, tensor_parallel_output_grad=False, group=_m.tp_group) if _m.sequence_parallel else input_
                y = _st.matmul(x)
                b = _m.bias if (_m.bias is not None and not _m.skip_bias_add) else None
                y = y + b if b is not None else y
                gather = _m.gather_output if runtime_gather_output is None else runtime_gather_output
                if gather and get_pg_size(_m.tp_group) > 1:
                    y = mp.gather_from_tensor_model_parallel_region(y, group=_m.tp_group)
                return y, (_m.bias if _m.skip_bias_add else None)
        elif isinstance(m, RowParallelLinear):
            def fwd(input_, _m=m, _st=st):
                y = _st.matmul(input_)
                if get_pg_size(_m.tp_group) > 1:
                    y = (mp.reduce_scatter_to_sequence_parallel_region if _m.sequence_parallel else mp.reduce_from_tensor_model_parallel_region)(y, group=_m.tp_group)
                if not _m.skip_bias_add:
                    return (y + _m.bias if _m.bias is not None else y), None
                return y, _m.bias
        else:
            def fwd(input_, _m=m, _st=st, _b=bias):
                y = _st.matmul(input_)
                return y + _b if _b is not None else y
        m.forward = fwd
    return states


def export_quantized_state_dict(model, states: Dict[str, QuantizedLinearState]) -> Dict[str, torch.Tensor]:
    out = {k: v for k, v in model.state_dict().items() if isinstance(v, torch.Tensor)}
    for path, st in states.items():
        for tname, t in st.tensors.items():
            out[f"{path}.{tname}"] = t
        out[f"{path}.quant_format"] = torch.tensor([FORMATS.index(st.fmt)])
    return out
