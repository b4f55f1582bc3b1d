from ..engine import StaticInferenceEngine  # noqa: F401
