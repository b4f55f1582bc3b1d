from .model import RetroConfig, RetroModel, chunked_cross_attention  # noqa: F401
