from .abstract_engine import AbstractEngine  # noqa: F401
from .dynamic_engine import DynamicInferenceEngine, EngineSuspendedError  # noqa: F401
from .static_engine import StaticInferenceEngine  # noqa: F401

AbstractEngine.register(DynamicInferenceEngine)
AbstractEngine.register(StaticInferenceEngine)
