from ..engine import DynamicInferenceEngine  # noqa: F401
from ..zmq_coordinator import EngineWorker  # noqa: F401  (the coordinator-driven engine loop: reference ``run_engine_with_coordinator``)


class EngineSuspendedError(RuntimeError):
    """Raised when work is submitted to an engine that has been paused for a weight swap."""
