"""Masked-LM style datasets for BERT and T5 (reference ``datasets/masked_dataset.py``, ``bert_dataset.py``, ``t5_dataset.py``).

``MaskedTokenDataset`` draws fixed-length token windows from a token source (an ``IndexedDataset`` or a synthetic stream), then
* BERT: 15 % of positions are selected; 80 % → ``[MASK]``, 10 % → random token, 10 % kept; labels/loss-mask mark the selected
  positions; two segments A/B with a sentence-order label.
* T5: span corruption — contiguous spans (mean length 3) are replaced by sentinel tokens in the encoder input and spelled out in the
  decoder target (``<sentinel_i> span_i …``).
Deterministic per index (``numpy`` RNG seeded by ``seed + idx``) so resuming by ``consumed_samples`` reproduces the stream."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch


@dataclass
class MaskedDatasetConfig:
    sequence_length: int = 512
    vocab_size: int = 30522
    masking_probability: float = 0.15
    random_seed: int = 1234
    cls_id: int = 101
    sep_id: int = 102
    mask_id: int = 103
    pad_id: int = 0
    # T5
    sequence_length_decoder: int = 128
    mean_span_length: float = 3.0
    num_sentinels: int = 100
    bos_id: int = 1
    eos_id: int = 2


class _TokenSource:
    """Either real documents (``IndexedDataset``) or a synthetic i.i.d. stream (``--mock-data``)."""

    def __init__(self, cfg: MaskedDatasetConfig, indexed=None, num_samples: int = 1 << 20):
        self.cfg, self.indexed, self.num_samples = cfg, indexed, num_samples

    def __len__(self):
        return self.num_samples

    def window(self, idx: int, n: int, rng: np.random.Generator) -> np.ndarray:
        if self.indexed is None:
            lo = max(self.cfg.cls_id, self.cfg.sep_id, self.cfg.mask_id, self.cfg.eos_id) + 1
            return rng.integers(lo, self.cfg.vocab_size - self.cfg.num_sentinels, size=n, dtype=np.int64)
        out, d = [], idx % len(self.indexed)
        while sum(len(o) for o in out) < n:
            out.append(np.asarray(self.indexed[d], dtype=np.int64))
            d = (d + 1) % len(self.indexed)
        return np.concatenate(out)[:n]


class BERTMaskedDataset(torch.utils.data.Dataset):
    def __init__(self, cfg: MaskedDatasetConfig, indexed=None, num_samples: int = 1 << 20, binary_head: bool = True):
        self.cfg, self.src, self.binary_head = cfg, _TokenSource(cfg, indexed, num_samples), binary_head

    def __len__(self):
        return len(self.src)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        c = self.cfg
        rng = np.random.default_rng(c.random_seed + idx)
        n = c.sequence_length - 3
        toks = self.src.window(idx, n, rng)
        cut = int(rng.integers(1, n - 1))
        a, b = toks[:cut], toks[cut:]
        is_random = 0
        if self.binary_head and rng.random() < 0.5:
            a, b, is_random = b, a, 1  # sentence-order prediction: swapped segments
        ids = np.concatenate([[c.cls_id], a, [c.sep_id], b, [c.sep_id]])
        types = np.concatenate([np.zeros(len(a) + 2, dtype=np.int64), np.ones(len(b) + 1, dtype=np.int64)])
        special = (ids == c.cls_id) | (ids == c.sep_id)
        cand = np.flatnonzero(~special)
        k = max(1, int(round(len(cand) * c.masking_probability)))
        chosen = rng.choice(cand, size=k, replace=False)
        labels = np.full_like(ids, -1)
        labels[chosen] = ids[chosen]
        r = rng.random(k)
        masked = ids.copy()
        masked[chosen[r < 0.8]] = c.mask_id
        rnd = chosen[(r >= 0.8) & (r < 0.9)]
        masked[rnd] = rng.integers(c.mask_id + 1, c.vocab_size, size=len(rnd))
        loss_mask = (labels >= 0).astype(np.float32)
        labels[labels < 0] = 0
        return {"text": torch.from_numpy(masked), "types": torch.from_numpy(types), "labels": torch.from_numpy(labels), "is_random": torch.tensor(is_random),
                "loss_mask": torch.from_numpy(loss_mask), "padding_mask": torch.ones(len(ids), dtype=torch.int64), "truncated": torch.tensor(0)}


class T5MaskedDataset(torch.utils.data.Dataset):
    def __init__(self, cfg: MaskedDatasetConfig, indexed=None, num_samples: int = 1 << 20):
        self.cfg, self.src = cfg, _TokenSource(cfg, indexed, num_samples)

    def __len__(self):
        return len(self.src)

    def sentinel(self, i: int) -> int:
        return self.cfg.vocab_size - 1 - i

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        c = self.cfg
        rng = np.random.default_rng(c.random_seed + idx)
        toks = self.src.window(idx, c.sequence_length, rng)
        n = len(toks)
        budget = max(1, int(round(n * c.masking_probability)))
        spans, used, pos = [], 0, 0
        while used < budget and len(spans) < c.num_sentinels and pos < n - 1:
            gap = int(rng.geometric(min(1.0, budget / max(1, n)) if False else 1.0 / max(1.0, (n - budget) / max(1, budget / c.mean_span_length))))
            pos += gap
            ln = int(min(max(1, rng.poisson(c.mean_span_length)), budget - used, n - pos))
            if pos >= n or ln <= 0:
                break
            spans.append((pos, ln))
            used += ln
            pos += ln
        enc, dec_t = [], []
        cur = 0
        for i, (s, ln) in enumerate(spans):
            enc.extend(toks[cur:s].tolist())
            enc.append(self.sentinel(i))
            dec_t.append(self.sentinel(i))
            dec_t.extend(toks[s : s + ln].tolist())
            cur = s + ln
        enc.extend(toks[cur:].tolist())
        dec_in = [c.bos_id] + dec_t
        dec_out = dec_t + [c.eos_id]
        le, ld = c.sequence_length, c.sequence_length_decoder
        dec_in, dec_out = dec_in[:ld], dec_out[:ld]

        def pad(x, L):
            return np.asarray(x + [c.pad_id] * (L - len(x)), dtype=np.int64)

        e_len, d_len = min(len(enc), le), len(dec_in)
        enc_mask = np.zeros(le, dtype=np.int64)
        enc_mask[:e_len] = 1
        dec_mask = np.zeros(ld, dtype=np.int64)
        dec_mask[:d_len] = 1
        loss_mask = dec_mask.astype(np.float32)
        return {"text_enc": torch.from_numpy(pad(enc[:le], le)), "text_dec": torch.from_numpy(pad(dec_in, ld)), "labels": torch.from_numpy(pad(dec_out, ld)),
                "loss_mask": torch.from_numpy(loss_mask), "enc_mask": torch.from_numpy(enc_mask), "dec_mask": torch.from_numpy(dec_mask), "truncated": torch.tensor(0)}
