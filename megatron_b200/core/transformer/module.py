"""Base module (reference ``transformer/module.py:67,479``)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..transformer.transformer_config import TransformerConfig

_FLOAT_TYPES = (torch.float32,)
_HALF_TYPES = (torch.float16,)
_BF16_TYPES = (torch.bfloat16,)


def param_is_not_shared(param):
    return not getattr(param, "shared", False)


class MegatronModule(torch.nn.Module):
    """``nn.Module`` that knows how to describe its parameters as a sharded state dict."""

    def __init__(self, config: TransformerConfig):
        super().__init__()
        self.config = config

    def state_dict_for_save_checkpoint(self, prefix: str = "", keep_vars: bool = False):
        return self.state_dict(prefix=prefix, keep_vars=keep_vars)

    def sharded_state_dict(self, prefix: str = "", sharded_offsets: Tuple[Tuple[int, int, int]] = (), metadata: Optional[dict] = None):
        """Default: own tensors are replicated across TP; recurse into children."""
        from .utils import make_sharded_tensors_for_checkpoint, sharded_state_dict_default

        sd = {}
        self._save_to_state_dict(sd, "", keep_vars=True)
        out = make_sharded_tensors_for_checkpoint(sd, prefix, sharded_offsets=sharded_offsets, tp_group=getattr(self, "tp_group", None))
        for name, child in self.named_children():
            out.update(sharded_state_dict_default(child, f"{prefix}{name}.", sharded_offsets, metadata))
        return out

    def set_is_first_microbatch(self):
        for m in self.modules():
            if hasattr(m, "is_first_microbatch"):
                m.is_first_microbatch = True

    def set_symmetric_ar(self, set_to=None):
        pass


class GraphableMegatronModule(MegatronModule):
    """Module whose forward may be captured in a CUDA graph (reference :168)."""

    def __init__(self, config, vp_stage=None):
        super().__init__(config)
        self.vp_stage = vp_stage
        self.cudagraph_manager = None
        if getattr(config, "cuda_graph_impl", "none") == "local" or config.enable_cuda_graph:
            from .cuda_graphs import CudaGraphManager

            self.cudagraph_manager = CudaGraphManager(config, vp_stage=vp_stage)

    def __call__(self, *args, **kwargs):
        if self.cudagraph_manager is not None and self.cudagraph_manager.should_graph(self, args, kwargs):
            return self.cudagraph_manager(self, args, kwargs)
        return super().__call__(*args, **kwargs)

    def _eager_forward(self, *args, **kwargs):
        return torch.nn.Module.__call__(self, *args, **kwargs)


def _convert(val, fn):
    if isinstance(val, (tuple, list)):
        return type(val)(_convert(v, fn) for v in val)
    return fn(val)


def fp32_to_float16(val, float16_convertor):
    def f(v):
        return float16_convertor(v) if isinstance(v, torch.Tensor) and v.dtype in _FLOAT_TYPES else v

    return _convert(val, f)


def float16_to_fp32(val):
    def f(v):
        return v.float() if isinstance(v, torch.Tensor) and v.dtype in (_HALF_TYPES + _BF16_TYPES) else v

    return _convert(val, f)


class Float16Module(MegatronModule):
    """Casts the wrapped module to fp16/bf16 and its pipeline inputs/outputs accordingly."""

    def __init__(self, config: TransformerConfig, module: torch.nn.Module):
        super().__init__(config)
        self.fp16, self.bf16 = config.fp16, config.bf16
        self.vp_stage = getattr(module, "vp_stage", None)
        if self.fp16:
            self.add_module("module", module.half())
            self.float16_convertor = lambda v: v.half()
        elif self.bf16:
            self.add_module("module", module.bfloat16())
            self.float16_convertor = lambda v: v.bfloat16()
        else:
            raise Exception("Float16Module needs fp16 or bf16")

    def set_input_tensor(self, input_tensor):
        return self.module.set_input_tensor(input_tensor)

    def forward(self, *inputs, fp32_output=True, **kwargs):
        from .. import parallel_state as ps

        if ps.is_pipeline_first_stage(ignore_virtual=False, vp_stage=self.vp_stage):
            inputs = fp32_to_float16(inputs, self.float16_convertor)
        out = self.module(*inputs, **kwargs)
        if ps.is_pipeline_last_stage(ignore_virtual=False, vp_stage=self.vp_stage) and fp32_output:
            out = float16_to_fp32(out)
        return out

    def state_dict(self, destination=None, prefix="", keep_vars=False):
        return self.module.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.module.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)

    def sharded_state_dict(self, prefix="", *args, **kwargs):
        return self.module.sharded_state_dict(prefix, *args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        self.module.load_state_dict(state_dict, strict=strict)
