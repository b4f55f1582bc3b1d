from ..text_generation import TextGenerationServer  # noqa: F401

MegatronServer = TextGenerationServer
