from ..mamba.mamba_layer_specs import mamba_stack_spec as hybrid_stack_spec  # noqa: F401
from ..mamba.mamba_layer_specs import mamba_stack_spec  # noqa: F401
