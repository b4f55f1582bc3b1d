from .zmq_coordinator import Headers  # noqa: F401
