from ..engine import StaticInferenceEngine as MCoreEngine  # noqa: F401  (deprecated reference name of the static engine)
