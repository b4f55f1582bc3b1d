from __future__ import annotations

import json
import os
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Union


class MegatronTokenizerBase(ABC):
    @abstractmethod
    def tokenize(self, text: str) -> List[int]:
        ...

    def detokenize(self, ids: List[int]) -> str:
        raise NotImplementedError

    @property
    @abstractmethod
    def vocab_size(self) -> int:
        ...

    @property
    def eod(self) -> int:
        raise NotImplementedError

    @property
    def pad(self) -> int:
        return -1

    @property
    def bos(self) -> Optional[int]:
        return None

    @property
    def eos(self) -> Optional[int]:
        return self.eod

    def offsets(self, ids: List[int], text: str) -> List[int]:
        out, pos = [], 0
        for i in ids:
            out.append(pos)
            pos += len(self.detokenize([i]))
        return out


class NullTokenizer(MegatronTokenizerBase):
    """Whitespace-separated integers; the last id is EOD (``--tokenizer-type NullTokenizer``)."""

    def __init__(self, vocab_size: int, eod_id: Optional[int] = None, pad_id: Optional[int] = None):
        self._vocab_size_without_eod = int(vocab_size)
        self._eod_id = self._vocab_size_without_eod if eod_id is None else int(eod_id)      # --null-tokenizer-eod-id
        self._pad_id = -1 if pad_id is None else int(pad_id)                                  # --null-tokenizer-pad-id (-1: no pad token)

    def tokenize(self, text: str) -> List[int]:
        return [int(x) for x in text.split()]

    def detokenize(self, ids: List[int]) -> str:
        return " ".join(str(int(i)) for i in ids)

    @property
    def vocab_size(self) -> int:
        return self._vocab_size_without_eod + 1

    @property
    def eod(self) -> int:
        return self._eod_id

    @property
    def pad(self) -> int:
        return self._pad_id

    @property
    def unique_identifiers(self):
        return {"class": "NullTokenizer", "vocab_size": self.vocab_size}


class ByteLevelTokenizer(MegatronTokenizerBase):
    def tokenize(self, text: str) -> List[int]:
        return list(text.encode("utf-8"))

    def detokenize(self, ids: List[int]) -> str:
        return bytes(int(i) for i in ids if i < 256).decode("utf-8", errors="replace")

    @property
    def vocab_size(self) -> int:
        return 257

    @property
    def eod(self) -> int:
        return 256


class SentencePieceTokenizer(MegatronTokenizerBase):
    def __init__(self, model_file: str, chat_template: Optional[str] = None, **_ignored):
        import sentencepiece

        self.sp = sentencepiece.SentencePieceProcessor(model_file=model_file)
        self.chat_template = chat_template

    def tokenize(self, text: str) -> List[int]:
        return self.sp.encode(text)

    def detokenize(self, ids: List[int]) -> str:
        return self.sp.decode(list(map(int, ids)))

    @property
    def vocab_size(self) -> int:
        return self.sp.get_piece_size()

    @property
    def eod(self) -> int:
        return self.sp.eos_id()

    @property
    def bos(self):
        return self.sp.bos_id()

    @property
    def pad(self):
        return self.sp.pad_id()


class HuggingFaceTokenizer(MegatronTokenizerBase):
    def __init__(self, path: str, include_special_tokens: bool = False, chat_template: Optional[str] = None, **kw):
        import transformers

        self.tk = transformers.AutoTokenizer.from_pretrained(path, **kw)
        self.include_special_tokens = include_special_tokens       # reference ``--tokenizer-hf-no-include-special-tokens`` turns this off
        if chat_template is not None:
            self.tk.chat_template = chat_template

    def tokenize(self, text: str) -> List[int]:
        return self.tk.encode(text, add_special_tokens=self.include_special_tokens)

    def detokenize(self, ids: List[int]) -> str:
        return self.tk.decode(list(map(int, ids)))

    @property
    def vocab_size(self) -> int:
        return len(self.tk)

    @property
    def eod(self) -> int:
        return self.tk.eos_token_id

    @property
    def bos(self):
        return self.tk.bos_token_id

    @property
    def pad(self):
        return self.tk.pad_token_id if self.tk.pad_token_id is not None else -1


class MegatronTokenizer:
    """Factory mirroring ``MegatronTokenizer.from_pretrained(metadata_path={"library": ...}, ...)``."""

    @staticmethod
    def from_pretrained(tokenizer_path: Optional[str] = None, metadata_path: Optional[Union[str, Dict[str, Any]]] = None, **kwargs):
        lib = None
        if isinstance(metadata_path, dict):
            lib = metadata_path.get("library")
        if lib in ("null", "null-text"):
            return NullTokenizer(kwargs.get("vocab_size", 256))
        if lib == "byte-level":
            return ByteLevelTokenizer()
        if lib == "sentencepiece":
            return SentencePieceTokenizer(tokenizer_path)
        if isinstance(metadata_path, str) or (metadata_path is None and tokenizer_path and os.path.isdir(tokenizer_path) and os.path.exists(os.path.join(tokenizer_path, "tokenizer_metadata.json"))):
            # reference layout: a ``tokenizer_metadata.json`` next to the tokenizer files names the library (``write_metadata``)
            with open(metadata_path if isinstance(metadata_path, str) else os.path.join(tokenizer_path, "tokenizer_metadata.json")) as f:
                meta = json.load(f)
            return MegatronTokenizer.from_pretrained(tokenizer_path, {**meta, **({} if not isinstance(metadata_path, dict) else metadata_path)}, **{**meta.get("kwargs", {}), **kwargs})
        if lib == "tiktoken":
            from .text.tiktoken_tokenizer import TikTokenTokenizer

            return TikTokenTokenizer(tokenizer_path, **{k: v for k, v in kwargs.items() if k in ("pattern", "vocab_size", "num_special_tokens", "special_tokens")})
        if lib == "sft":
            from .text.sft_tokenizer import SFTTokenizer

            base = MegatronTokenizer.from_pretrained(tokenizer_path, {"library": kwargs.pop("base_library", "huggingface")}, **kwargs)
            return SFTTokenizer(base, kwargs.get("prompt_format", "chatml"))
        if lib in ("multimodal", "null-multimodal"):
            from .vision.multimodal_tokenizer import MultimodalTokenizer

            base = NullTokenizer(kwargs.get("vocab_size", 256)) if lib == "null-multimodal" else MegatronTokenizer.from_pretrained(
                tokenizer_path, {"library": kwargs.pop("base_library", "huggingface")})
            return MultimodalTokenizer(base, kwargs.get("num_image_tokens", 576), kwargs.get("image_token_id"))
        if lib in ("huggingface", None) and tokenizer_path:
            return HuggingFaceTokenizer(tokenizer_path, **{k: v for k, v in kwargs.items() if k != "vocab_size"})
        raise ValueError(f"cannot build a tokenizer from library={lib!r}, path={tokenizer_path!r}")

    @staticmethod
    def write_metadata(tokenizer_path: str, tokenizer_library: str, model_type: Optional[str] = None, chat_template: Optional[str] = None, overwrite: bool = False,
                       metadata_path: Optional[str] = None, **kwargs) -> str:
        """Describe a tokenizer directory / file so ``from_pretrained(path)`` can restore it without arguments (reference ``megatron_tokenizer.py:write_metadata``)."""
        target = metadata_path or os.path.join(tokenizer_path if os.path.isdir(tokenizer_path) else os.path.dirname(tokenizer_path), "tokenizer_metadata.json")
        if os.path.exists(target) and not overwrite:
            raise FileExistsError(f"{target} exists (overwrite=False)")
        with open(target, "w") as f:
            json.dump({"library": tokenizer_library, "model_type": model_type, "chat_template": chat_template, "kwargs": kwargs}, f, indent=1)
        return target


def build_tokenizer_from_args(args):
    """The tokenizer a training / serving command line describes (reference ``tokenizers/utils/build_tokenizer.py:build_tokenizer``): tokenizer type and model
    plus the per-library switches (``--null-tokenizer-eod-id``, ``--tiktoken-pattern``, ``--tokenizer-hf-no-use-fast``, ``--trust-remote-code``,
    ``--chat-template``, ``--tokenizer-special-tokens``, ``--tokenizer-metadata``)."""
    g = lambda n, d=None: getattr(args, n, d)  # noqa: E731
    if g("tokenizer_metadata"):
        return MegatronTokenizer.from_pretrained(g("tokenizer_model"), g("tokenizer_metadata"))
    t = (g("tokenizer_type") or "NullTokenizer").lower()
    special = g("tokenizer_special_tokens") or g("special_tokens")
    kw = {}
    if t == "nulltokenizer":
        kw = {"eod_id": g("null_tokenizer_eod_id"), "pad_id": g("null_tokenizer_pad_id")}
    elif t in ("tiktokentokenizer", "tiktoken", "tiktokenizer"):
        kw = {"pattern": g("tiktoken_pattern"), "num_special_tokens": g("tiktoken_num_special_tokens", 1000) or 1000, "special_tokens": special}
    elif t in ("huggingfacetokenizer", "hf"):
        kw = {"use_fast": not g("tokenizer_hf_no_use_fast", False), "trust_remote_code": bool(g("trust_remote_code", False)),
              "include_special_tokens": not g("tokenizer_hf_no_include_special_tokens", False), "chat_template": g("chat_template")}
        if special:
            kw["additional_special_tokens"] = list(special)
    elif t in ("sentencepiecetokenizer", "llama2tokenizer", "gptsentencepiecetokenizer"):
        kw = {"chat_template": g("chat_template"), "legacy": g("tokenizer_sentencepiece_legacy", False)}
    return build_tokenizer(g("tokenizer_type") or "NullTokenizer", vocab_size=g("vocab_size"), tokenizer_model=g("tokenizer_model"), **kw)


def build_tokenizer(tokenizer_type: str, vocab_size: Optional[int] = None, tokenizer_model: Optional[str] = None, **kw):
    t = tokenizer_type.lower()
    if t == "nulltokenizer":
        return NullTokenizer(vocab_size, **{k: v for k, v in kw.items() if k in ("eod_id", "pad_id")})
    if t in ("bytelevel", "byteleveltokenizer"):
        return ByteLevelTokenizer()
    if t in ("sentencepiecetokenizer", "llama2tokenizer", "gptsentencepiecetokenizer"):
        return SentencePieceTokenizer(tokenizer_model, **kw)
    if t in ("huggingfacetokenizer", "hf"):
        return HuggingFaceTokenizer(tokenizer_model, **kw)
    if t in ("tiktokentokenizer", "tiktoken", "tiktokenizer"):
        from .text.tiktoken_tokenizer import TikTokenTokenizer

        return TikTokenTokenizer(tokenizer_model, vocab_size=vocab_size, **kw)
    raise ValueError(f"unknown tokenizer type {tokenizer_type}")
