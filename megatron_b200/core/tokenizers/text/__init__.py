"""Text tokenizer libraries (reference ``megatron/core/tokenizers/text``)."""
from .chat_template import ChatTemplate  # noqa: F401
from .sft_tokenizer import SFTTokenizer  # noqa: F401
from .tiktoken_tokenizer import TikTokenTokenizer  # noqa: F401
