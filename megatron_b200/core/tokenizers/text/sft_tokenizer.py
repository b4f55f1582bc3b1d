"""Supervised-fine-tuning tokenizer (reference ``text/libraries/sft_tokenizer.py``): wraps a base tokenizer and a chat template; a conversation becomes
(tokens, targets) where every position that is not an assistant token carries ``IGNORE_INDEX`` so the loss skips prompts and system text."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from ..tokenizer import MegatronTokenizerBase
from .chat_template import ChatTemplate

IGNORE_INDEX = -100


class SFTTokenizer(MegatronTokenizerBase):
    def __init__(self, base: MegatronTokenizerBase, prompt_format: str = "chatml", max_length: Optional[int] = None):
        self.base = base
        self.template = ChatTemplate(prompt_format)
        self.max_length = max_length

    def tokenize(self, text: str) -> List[int]:
        return self.base.tokenize(text)

    def detokenize(self, ids: List[int]) -> str:
        return self.base.detokenize([i for i in ids if i != IGNORE_INDEX])

    @property
    def vocab_size(self) -> int:
        return self.base.vocab_size

    @property
    def eod(self) -> int:
        return self.base.eod

    @property
    def pad(self) -> int:
        return self.base.pad

    def tokenize_conversation(self, conversation: List[Dict[str, str]], return_target: bool = True, add_generation_prompt: bool = False):
        """``conversation``: ``[{"role": "system"|"user"|"assistant", "content": str}, …]`` → tokens, or (tokens, targets) with next-token targets masked
        to ``IGNORE_INDEX`` outside assistant turns (targets[i] is the label of position i, i.e. tokens[i + 1])."""
        ids, mask = self.template.tokenize_conversation(self.base, conversation, add_generation_prompt)
        if self.max_length is not None:
            ids, mask = ids[: self.max_length], mask[: self.max_length]
        if not return_target:
            return ids
        targets = [ids[i + 1] if i + 1 < len(ids) and mask[i + 1] else IGNORE_INDEX for i in range(len(ids))]
        return ids, targets
