"""Tokenizers (reference ``megatron/core/tokenizers``): a uniform ``MegatronTokenizer`` front
over null / sentencepiece / HuggingFace / byte-level back ends."""
from .tokenizer import MegatronTokenizer, NullTokenizer, build_tokenizer, build_tokenizer_from_args

__all__ = ["MegatronTokenizer", "NullTokenizer", "build_tokenizer", "build_tokenizer_from_args"]
