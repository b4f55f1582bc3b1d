"""Online weight resharding: checkpoint-free movement of weights between parallel layouts (reference ``megatron/core/resharding``)."""
from .planner import ShardDesc, TransferOp, build_reshard_plan, build_local_reshard_plan, build_centralized_reshard_plan  # noqa: F401
from .execute import execute_reshard_plan as execute_box_plan, reshard_state_dict  # noqa: F401
from .execute import execute_reshard_plan as _box_execute  # noqa: F401
from .execution import execute_reshard_plan as execute_refit_plan  # noqa: F401
from .refit import (  # noqa: F401
    clear_all_caches, clear_plan_cache, clear_service_cache, get_or_create_service, prepare_swap_model_weights, reshard_model_weights, swap_model_weights,
)
from .transforms import MXFP8ReshardTransform, ReshardTransform  # noqa: F401
from .utils import ParameterMetadata, ReshardPlan, ShardingDescriptor, extract_param_metadata  # noqa: F401

execute_reshard_plan = _box_execute      # the box-plan executor keeps its public name (state-dict resharding)
