"""Fill-in-the-middle transform for code pre-training (reference ``training/datasets/fim_dataset.py``): with probability ``fim_rate`` a
document is cut into (prefix, middle, suffix) at two random points and re-ordered as PSM (``<pre> p <suf> s <mid> m``) or, with probability
``fim_spm_rate``, SPM (``<pre> <suf> s <mid> p m``).  Applied per document inside a packed GPT sample (split at EOD), keeping the sample length."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch


@dataclass
class FIMConfig:
    fim_rate: float = 0.5
    fim_spm_rate: float = 0.5
    prefix_id: int = 0
    middle_id: int = 1
    suffix_id: int = 2
    pad_id: int = 3
    eod_id: int = 4
    min_doc_len: int = 4           # shorter fragments are left alone
    truncate_or_pad: bool = True


def _fim_doc(doc: np.ndarray, rng: np.random.RandomState, cfg: FIMConfig) -> np.ndarray:
    if len(doc) < cfg.min_doc_len or rng.binomial(1, cfg.fim_rate) == 0:
        return doc
    a, b = sorted(int(x) for x in rng.randint(0, len(doc) + 1, size=2))
    p, m, s = doc[:a], doc[a:b], doc[b:]
    if cfg.truncate_or_pad:        # keep the document length: the three sentinels replace tokens at the end of the suffix (or pad)
        over = 3
        if len(s) >= over:
            s = s[: len(s) - over]
        else:
            return doc
    pre, mid, suf = (np.array([x], dtype=doc.dtype) for x in (cfg.prefix_id, cfg.middle_id, cfg.suffix_id))
    if rng.binomial(1, cfg.fim_spm_rate):
        return np.concatenate([pre, suf, s, mid, p, m])
    return np.concatenate([pre, p, suf, s, mid, m])


def apply_fim(tokens: np.ndarray, rng: np.random.RandomState, cfg: FIMConfig) -> np.ndarray:
    """Transform every EOD-delimited document of one packed sample; output has the input's length."""
    tokens = np.asarray(tokens)
    cuts = np.nonzero(tokens == cfg.eod_id)[0]
    out: List[np.ndarray] = []
    start = 0
    for c in cuts:
        out.append(_fim_doc(tokens[start:c], rng, cfg))
        out.append(tokens[c : c + 1])
        start = c + 1
    if start < len(tokens):
        out.append(_fim_doc(tokens[start:], rng, cfg))
    res = np.concatenate(out) if out else tokens
    n = len(tokens)
    if len(res) < n:
        res = np.concatenate([res, np.full(n - len(res), cfg.pad_id, dtype=tokens.dtype)])
    return res[:n]


class GPTFIMDataset(torch.utils.data.Dataset):
    """Wrap a GPT dataset yielding ``{"tokens": [s+1] or tokens/labels}``; the transform is a pure function of (seed, index) → reproducible
    across restarts and independent of the number of data-loader workers."""

    def __init__(self, base, config: FIMConfig, seed: int = 1234):
        self.base, self.cfg, self.seed = base, config, seed

    def __len__(self) -> int:
        return len(self.base)

    def __getitem__(self, idx: int):
        item = dict(self.base[idx])
        rng = np.random.RandomState(seed=[self.seed, int(idx)])
        if "text" in item:
            item["text"] = torch.from_numpy(apply_fim(np.asarray(item["text"]), rng, self.cfg))
            return item
        toks = torch.cat([item["tokens"], item["labels"][-1:]]).numpy()
        new = apply_fim(toks, rng, self.cfg)
        item["tokens"], item["labels"] = torch.from_numpy(new[:-1].copy()), torch.from_numpy(new[1:].copy())
        if "loss_mask" in item:
            item["loss_mask"] = item["loss_mask"] * (item["labels"] != self.cfg.pad_id).to(item["loss_mask"].dtype)
        return item
