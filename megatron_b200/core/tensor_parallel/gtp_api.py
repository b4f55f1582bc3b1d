"""Public GTP entry points (reference ``tensor_parallel/gtp_api.py``)."""
from .generalized_tensor_parallelism import GTPPrefetcher, apply_expert_gtp, apply_gtp, convert_linear_to_gtp, gather_gtp_state_dict  # noqa: F401
