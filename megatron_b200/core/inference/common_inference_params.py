from .sampling import SamplingParams as CommonInferenceParams  # noqa: F401  (deprecated reference alias)
