from .span_groups import SpanRecorder, get_recorder, set_recorder, span  # noqa: F401
from .training_metrics import TrainingMetrics  # noqa: F401
