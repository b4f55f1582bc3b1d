from .sampling import SamplingParams  # noqa: F401
