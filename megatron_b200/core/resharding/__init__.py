from .planner import ShardDesc, TransferOp, build_reshard_plan  # noqa: F401
from .execute import execute_reshard_plan, reshard_state_dict  # noqa: F401
