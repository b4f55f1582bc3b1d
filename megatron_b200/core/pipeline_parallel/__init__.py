"""Pipeline-parallel schedules and point-to-point communication (reference ``megatron/core/pipeline_parallel/__init__.py`` exports the schedule selector)."""
from .schedules import get_forward_backward_func  # noqa: F401
