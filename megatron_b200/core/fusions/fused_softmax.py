"""scale + mask + softmax (reference ``fusions/fused_softmax.py:11-359``).  On the flash path this op does not exist as a separate
kernel (it is the TMEM→register stage of ``csrc/flash_attn_sm100.cu``); this module serves the unfused / arbitrary-mask path."""
import torch

from ...ops import reference as ref
from ..transformer.enums import AttnMaskType


class FusedScaleMaskSoftmax(torch.nn.Module):
    def __init__(self, input_in_fp16=False, input_in_bf16=True, attn_mask_type=AttnMaskType.padding, scaled_masked_softmax_fusion=True, mask_func=None,
                 softmax_in_fp32=True, scale=None, window_size=None):
        super().__init__()
        self.attn_mask_type, self.scale, self.softmax_in_fp32 = attn_mask_type, scale, softmax_in_fp32

    def forward(self, input: torch.Tensor, mask, softmax_offset=None):
        causal = self.attn_mask_type == AttnMaskType.causal and mask is None
        if softmax_offset is not None:  # "softmax-one" / sink variant: an extra per-head logit joins the denominator
            x = input.float() * (self.scale or 1.0)
            if mask is not None:
                x = x.masked_fill(mask.bool(), float("-inf"))
            sink = softmax_offset.reshape(1, -1, 1, 1).float().expand(x.shape[0], -1, x.shape[2], 1)
            p = torch.softmax(torch.cat([x, sink], dim=-1), dim=-1)[..., :-1]
            return p.to(input.dtype)
        from ... import ops

        return ops.scaled_masked_softmax(input, None if mask is None else mask.bool(), self.scale or 1.0, causal=causal)
