"""megatron.core-compatible API surface, implemented B200-first."""
from . import parallel_state, tensor_parallel, utils
from .model_parallel_config import ModelParallelConfig
from .package_info import __version__
from .timers import Timers
from .inference_params import InferenceParams

mpu = parallel_state


def __getattr__(name):
    if name == "DistributedDataParallel":
        from .distributed import DistributedDataParallel

        return DistributedDataParallel
    if name == "dist_checkpointing":
        import importlib

        return importlib.import_module(".dist_checkpointing", __name__)
    raise AttributeError(name)
from .safe_globals import register_safe_globals, safe_load_from_bytes  # noqa: E402,F401
