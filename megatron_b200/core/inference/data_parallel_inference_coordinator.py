from .data_parallel_coordinator import DataParallelInferenceCoordinator  # noqa: F401  (in-process router over engine replicas)
from .zmq_coordinator import ZMQCoordinator  # noqa: F401  (the socket-level coordinator process)
