from .text_generation_server import MegatronServer, TextGenerationServer  # noqa: F401
