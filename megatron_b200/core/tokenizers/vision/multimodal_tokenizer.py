"""Tokenizer for interleaved text + images (reference ``vision/libraries/multimodal_tokenizer.py``): a text tokenizer plus the image placeholder protocol —
``<image>`` in the text stands for ONE image and is replaced by ``image_token_id`` repeated ``num_image_tokens`` times (the number of vision-encoder
output embeddings per image, optionally with ``<img>`` / ``</img>`` delimiters), so the language model sees the right sequence length and the embedding
layer can scatter the vision features into those positions."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from ..tokenizer import MegatronTokenizerBase

IMAGE_TOKEN = "<image>"
IMAGE_TOKEN_INDEX = -200          # placeholder id used by LLaVA-style pipelines before expansion


class MultimodalTokenizer(MegatronTokenizerBase):
    def __init__(self, base: MegatronTokenizerBase, num_image_tokens: int = 576, image_token_id: Optional[int] = None, use_delimiters: bool = False,
                 prompt_format: Optional[str] = None):
        self.base = base
        self.num_image_tokens = num_image_tokens
        self.image_token_id = IMAGE_TOKEN_INDEX if image_token_id is None else image_token_id
        self.use_delimiters = use_delimiters
        self._img_start = base.tokenize("<img>") if use_delimiters else []
        self._img_end = base.tokenize("</img>") if use_delimiters else []
        self.template = None
        if prompt_format is not None:
            from ..text.chat_template import ChatTemplate

            self.template = ChatTemplate(prompt_format)

    def tokenize(self, text: str, expand_images: bool = True) -> List[int]:
        out: List[int] = []
        parts = text.split(IMAGE_TOKEN)
        for i, part in enumerate(parts):
            if part:
                out += self.base.tokenize(part)
            if i + 1 < len(parts):
                out += self._img_start + [self.image_token_id] * (self.num_image_tokens if expand_images else 1) + self._img_end
        return out

    def image_positions(self, ids: List[int]) -> List[Tuple[int, int]]:
        """[(start, end)) runs of image tokens — where the vision embeddings go."""
        runs, start = [], None
        for i, t in enumerate(list(ids) + [None]):
            if t == self.image_token_id and start is None:
                start = i
            elif t != self.image_token_id and start is not None:
                runs.append((start, i))
                start = None
        return runs

    def tokenize_conversation(self, conversation: List[Dict[str, str]], return_target: bool = True, add_generation_prompt: bool = False):
        assert self.template is not None, "prompt_format was not given"
        from ..text.sft_tokenizer import IGNORE_INDEX

        ids, mask = self.template.tokenize_conversation(self, conversation, add_generation_prompt)
        if not return_target:
            return ids
        targets = [ids[i + 1] if i + 1 < len(ids) and mask[i + 1] and ids[i + 1] != self.image_token_id else IGNORE_INDEX for i in range(len(ids))]
        return ids, targets

    def detokenize(self, ids: List[int]) -> str:
        return self.base.detokenize([i for i in ids if i != self.image_token_id and i >= 0])

    @property
    def vocab_size(self) -> int:
        return self.base.vocab_size

    @property
    def eod(self) -> int:
        return self.base.eod

    @property
    def pad(self) -> int:
        return self.base.pad
