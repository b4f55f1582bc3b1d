from .text_generation_controller import TextGenerationController  # noqa: F401
