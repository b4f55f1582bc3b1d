from ..text_generation import IncrementalDetokenizer, TextGenerationController  # noqa: F401
