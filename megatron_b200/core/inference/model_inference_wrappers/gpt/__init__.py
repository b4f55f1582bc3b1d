from .gpt_inference_wrapper import GPTInferenceWrapper  # noqa: F401
