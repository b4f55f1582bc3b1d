"""Byte-pair tokenizer over a tiktoken-format vocabulary (reference ``text/libraries/tiktoken_tokenizer.py``).

The reference wraps the ``tiktoken`` package; that package is not a dependency here, so the algorithm is implemented directly: the vocabulary file is
either the ``.tiktoken`` text format (``base64(token bytes) rank`` per line) or the reference's JSON list (``[{"rank", "token_bytes", "token_str"}, …]``);
text is split by the pattern, every piece is encoded greedily by lowest merge rank (the byte-pair procedure tiktoken uses), special tokens take the ids
after the mergeable ranks.  Same layout as the reference: ``num_special_tokens`` slots (``<unk> <s> </s>`` first, then ``<SPECIAL_i>`` fillers) are placed
BEFORE the mergeable ranks, so ids are ``rank + num_special_tokens``."""
from __future__ import annotations

import base64
import json
import re
from typing import Dict, List, Optional

from ..tokenizer import MegatronTokenizerBase

# tiktoken's cl100k-style splitter without the unicode-property classes `re` lacks: contractions, letters, digit runs (<= 3), punctuation, whitespace
DEFAULT_PATTERN = r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\w]?[^\W\d_]+|\d{1,3}| ?[^\s\w]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"""
DEFAULT_SPECIAL = ["<unk>", "<s>", "</s>"]
# the two named splitters of the reference command line (``--tiktoken-pattern v1|v2``): the public cl100k-style pattern and the case-aware one used by
# tekken-style vocabularies.  Both need unicode property classes, i.e. the ``regex`` package; without it the ``re`` approximation above is used.
NAMED_PATTERNS = {
    "v1": r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+""",
    "v2": r"""[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+""",
}


def compile_pattern(pattern: Optional[str]):
    """``None`` → the built-in ``re`` splitter; ``"v1"`` / ``"v2"`` → the reference's named patterns; anything else is a regular expression (compiled with
    ``regex`` when it uses ``\\p{..}`` classes)."""
    if pattern is None:
        return re.compile(DEFAULT_PATTERN)
    text = NAMED_PATTERNS.get(pattern, pattern)
    if "\\p{" in text:
        try:
            import regex
        except ImportError:
            return re.compile(DEFAULT_PATTERN)
        return regex.compile(text)
    return re.compile(text)


def load_tiktoken_ranks(path: str, vocab_size: Optional[int] = None) -> Dict[bytes, int]:
    ranks: Dict[bytes, int] = {}
    with open(path, "rb") as f:
        head = f.read(1)
        f.seek(0)
        if head == b"[":
            for e in json.load(f):
                ranks[base64.b64decode(e["token_bytes"])] = int(e["rank"])
        else:
            for line in f:
                if line.strip():
                    tok, rank = line.split()
                    ranks[base64.b64decode(tok)] = int(rank)
    if vocab_size is not None:
        ranks = {k: v for k, v in ranks.items() if v < vocab_size}
    return ranks


def bpe_encode_piece(piece: bytes, ranks: Dict[bytes, int]) -> List[int]:
    """Greedy byte-pair merge by lowest rank (tiktoken ``_byte_pair_merge``)."""
    if piece in ranks:
        return [ranks[piece]]
    parts = [bytes([b]) for b in piece]
    while len(parts) > 1:
        best, best_i = None, -1
        for i in range(len(parts) - 1):
            r = ranks.get(parts[i] + parts[i + 1])
            if r is not None and (best is None or r < best):
                best, best_i = r, i
        if best is None:
            break
        parts[best_i : best_i + 2] = [parts[best_i] + parts[best_i + 1]]
    return [ranks[p] for p in parts]


class TikTokenTokenizer(MegatronTokenizerBase):
    def __init__(self, path: str, pattern: Optional[str] = None, vocab_size: Optional[int] = None, num_special_tokens: int = 1000,
                 special_tokens: Optional[List[str]] = None):
        special = list(special_tokens or DEFAULT_SPECIAL)
        assert len(special) == len(set(special)) and len(special) <= num_special_tokens
        special += [f"<SPECIAL_{i}>" for i in range(len(special), num_special_tokens)]
        self.special_tokens = special
        self.num_special_tokens = num_special_tokens
        self.ranks = load_tiktoken_ranks(path, None if vocab_size is None else vocab_size - num_special_tokens)
        self.decoder = {v: k for k, v in self.ranks.items()}
        self.pattern = compile_pattern(pattern)
        self._special_ids = {t: i for i, t in enumerate(special)}
        self._special_re = re.compile("|".join(re.escape(t) for t in sorted(special, key=len, reverse=True)))
        self._vocab_size = num_special_tokens + len(self.ranks)
        self._unk, self._bos, self._eos = (self._special_ids.get(t) for t in ("<unk>", "<s>", "</s>"))

    def _encode_ordinary(self, text: str) -> List[int]:
        out: List[int] = []
        for piece in self.pattern.findall(text):
            out += [r + self.num_special_tokens for r in bpe_encode_piece(piece.encode("utf-8"), self.ranks)]
        return out

    def tokenize(self, text: str, bos: bool = False, eos: bool = False, allowed_special: bool = True) -> List[int]:
        ids: List[int] = [self._bos] if bos and self._bos is not None else []
        pos = 0
        if allowed_special:
            for m in self._special_re.finditer(text):
                ids += self._encode_ordinary(text[pos : m.start()])
                ids.append(self._special_ids[m.group()])
                pos = m.end()
        ids += self._encode_ordinary(text[pos:])
        if eos and self._eos is not None:
            ids.append(self._eos)
        return ids

    def detokenize(self, ids: List[int], skip_special_tokens: bool = False) -> str:
        buf = bytearray()
        for i in ids:
            i = int(i)
            if i < self.num_special_tokens:
                if not skip_special_tokens:
                    buf += self.special_tokens[i].encode()
            else:
                buf += self.decoder[i - self.num_special_tokens]
        return buf.decode("utf-8", errors="replace")

    @property
    def vocab_size(self) -> int:
        return self._vocab_size

    @property
    def eod(self) -> int:
        return self._eos

    @property
    def bos(self):
        return self._bos

    @property
    def eos(self):
        return self._eos

    @property
    def unk(self):
        return self._unk

    @property
    def pad(self) -> int:
        return self._special_ids.get("<pad>", -1)

    @property
    def vocab(self) -> Dict[str, int]:
        v = dict(self._special_ids)
        v.update({k.decode("utf-8", errors="replace"): r + self.num_special_tokens for k, r in self.ranks.items()})
        return v
