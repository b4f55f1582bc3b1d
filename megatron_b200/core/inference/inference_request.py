"""Request records (reference ``inference/inference_request.py``)."""
import enum

from .engine import InferenceRequest  # noqa: F401

DynamicInferenceRequest = InferenceRequest          # one record type serves both engines here


class Status(enum.Enum):
    """Life cycle of a request; ``InferenceRequest.status`` holds the lower-case name of the state it is in."""

    WAITING_IN_QUEUE = "waiting"
    ACTIVE_AND_GENERATING_TOKENS = "running"
    ACTIVE_BUT_NOT_GENERATING_TOKENS = "prefilling"
    COMPLETED = "finished"
    FAILED = "failed"

    @classmethod
    def of(cls, request: InferenceRequest) -> "Status":
        return cls(request.status)
