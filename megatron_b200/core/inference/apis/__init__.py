from ..inference_request import DynamicInferenceRequest, InferenceRequest, Status  # noqa: F401
from ..sampling_params import SamplingParams  # noqa: F401
from .async_llm import AsyncLLM, MegatronAsyncLLM  # noqa: F401
