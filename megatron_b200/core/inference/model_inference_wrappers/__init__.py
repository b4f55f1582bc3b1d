from .abstract_model_inference_wrapper import AbstractModelInferenceWrapper  # noqa: F401
from .gpt import GPTInferenceWrapper  # noqa: F401
from .inference_wrapper_config import InferenceWrapperConfig  # noqa: F401
