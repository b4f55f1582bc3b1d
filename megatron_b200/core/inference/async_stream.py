from .text_generation import AsyncStream  # noqa: F401
