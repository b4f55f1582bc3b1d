from ..text_generation import AsyncLLM  # noqa: F401

MegatronAsyncLLM = AsyncLLM
