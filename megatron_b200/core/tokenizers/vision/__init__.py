"""Multimodal tokenizers (reference ``megatron/core/tokenizers/vision``)."""
from .multimodal_tokenizer import MultimodalTokenizer  # noqa: F401
