"""TransformerEngine adapter (reference ``extensions/transformer_engine.py``, 3.7 kLoC of wrappers around TE modules).

Nothing here imports TE.  The names the reference's specs use are bound to this framework's own layers, whose hot paths are the
in-tree tcgen05 kernels: fused norm+linear is ``ColumnParallelLinear`` with the RMSNorm kernel in front (``fused_residual_rmsnorm``),
``TEDotProductAttention`` is the flash-attention kernel path of ``DotProductAttention``, grouped linears are ``GroupedMLP``'s grouped
GEMM, fp8 recipes are ``core/fp8_utils``.  ``TELayerNormColumnParallelLinear`` is a real norm+linear module with TE's parameter names, so a spec or checkpoint written against the
TE names keeps its normalisation and its keys."""
from ..tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
from ..transformer.dot_product_attention import DotProductAttention
from ..transformer.torch_norm import WrappedTorchNorm

HAVE_TE = False
TEColumnParallelLinear = ColumnParallelLinear
TERowParallelLinear = RowParallelLinear


class LayerNormColumnParallelLinear(ColumnParallelLinear):
    """Norm + column-parallel linear in one module with TransformerEngine's parameter names
    (``layer_norm_weight`` / ``layer_norm_bias`` / ``weight`` / ``bias``; reference
    ``extensions/transformer_engine.py:TELayerNormColumnParallelLinear``).  The norm runs in the sm_100a
    RMSNorm/LayerNorm kernel on the (sequence-parallel) input shard; the linear is the AG->GEMM pair op."""

    def __init__(self, input_size, output_size, *, config, init_method, gather_output=False, bias=True, skip_bias_add=False, is_expert=False,
                 skip_weight_param_allocation=False, tp_comm_buffer_name=None, tp_group=None, stride=1, **kw):
        import torch

        super().__init__(input_size, output_size, config=config, init_method=init_method, bias=bias, gather_output=gather_output, skip_bias_add=skip_bias_add,
                         is_expert=is_expert, skip_weight_param_allocation=skip_weight_param_allocation, tp_comm_buffer_name=tp_comm_buffer_name,
                         tp_group=tp_group, stride=stride)
        self.normalization = config.normalization
        self.eps = config.layernorm_epsilon
        self.zero_centered_gamma = config.layernorm_zero_centered_gamma
        dev = self.weight.device if self.weight is not None else "cpu"
        self.layer_norm_weight = torch.nn.Parameter(torch.full((input_size,), 0.0 if self.zero_centered_gamma else 1.0, dtype=config.params_dtype, device=dev))
        setattr(self.layer_norm_weight, "sequence_parallel", config.sequence_parallel)
        if self.normalization == "LayerNorm":
            self.layer_norm_bias = torch.nn.Parameter(torch.zeros(input_size, dtype=config.params_dtype, device=dev))
            setattr(self.layer_norm_bias, "sequence_parallel", config.sequence_parallel)
        else:
            self.register_parameter("layer_norm_bias", None)

    def forward(self, x, weight=None, runtime_gather_output=None):
        from ... import ops

        if self.normalization == "RMSNorm":
            x = ops.rms_norm(x, self.layer_norm_weight, self.eps, self.zero_centered_gamma)
        else:
            x = ops.layer_norm(x, self.layer_norm_weight, self.layer_norm_bias, self.eps, self.zero_centered_gamma)
        return super().forward(x, weight=weight, runtime_gather_output=runtime_gather_output)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..transformer.utils import make_sharded_tensors_for_checkpoint

        sd = self.state_dict(prefix="", keep_vars=True)
        return make_sharded_tensors_for_checkpoint(sd, prefix, {"weight": 0, "bias": 0}, sharded_offsets, tp_group=self.tp_group)


TELayerNormColumnParallelLinear = LayerNormColumnParallelLinear
TEDotProductAttention = DotProductAttention
TENorm = WrappedTorchNorm


def get_cpu_offload_context(*a, **k):
    from contextlib import nullcontext

    return nullcontext(), (lambda t: t)
