"""Supervised fine-tuning dataset (reference ``training/datasets/sft_dataset.py``): chat conversations → token ids with the loss restricted
to assistant turns, optionally several conversations packed into one fixed-length sample with per-document boundaries (``cu_seqlens``) for
THD attention.

Input format: JSONL, one conversation per line — ``{"messages": [{"role": "system"|"user"|"assistant", "content": str}, ...]}``.
The prompt template is explicit (no dependence on a HF chat template): ``<|role|>\\n{content}<|end|>\\n`` with role/end markers taken from the
config so any tokenizer's special tokens can be used."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

IGNORE_INDEX = -100


@dataclass
class SFTDatasetConfig:
    sequence_length: int
    pad_token_id: int = 0
    role_template: str = "<|{role}|>\n"
    end_of_turn: str = "<|end|>\n"
    train_on_roles: Tuple[str, ...] = ("assistant",)
    pack: bool = True
    truncate: bool = True
    seed: int = 1234


def tokenize_conversation(messages: Sequence[dict], tokenizer, cfg: SFTDatasetConfig) -> Tuple[List[int], List[int]]:
    """→ (token ids, per-token trainable flags).  The role header is never trained on; the end-of-turn marker of a trained turn is."""
    ids: List[int] = []
    train: List[int] = []
    for m in messages:
        head = tokenizer.tokenize(cfg.role_template.format(role=m["role"]))
        body = tokenizer.tokenize(m["content"] + cfg.end_of_turn)
        on = 1 if m["role"] in cfg.train_on_roles else 0
        ids += head + body
        train += [0] * len(head) + [on] * len(body)
    return ids, train


def pack_conversations(lengths: Sequence[int], capacity: int) -> List[List[int]]:
    """First-fit-decreasing bin packing of conversation indices into samples of ``capacity`` tokens (deterministic)."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    bins: List[List[int]] = []
    room: List[int] = []
    for i in order:
        n = min(lengths[i], capacity)
        for b in range(len(bins)):
            if room[b] >= n:
                bins[b].append(i)
                room[b] -= n
                break
        else:
            bins.append([i])
            room.append(capacity - n)
    return bins


class SFTDataset(torch.utils.data.Dataset):
    def __init__(self, conversations: Sequence[Sequence[dict]], tokenizer, config: SFTDatasetConfig):
        self.cfg = config
        self.items = [tokenize_conversation(c, tokenizer, config) for c in conversations]
        cap = config.sequence_length + 1                      # +1: inputs/labels are the shifted views
        if config.truncate:
            self.items = [(i[:cap], t[:cap]) for i, t in self.items]
        else:
            self.items = [(i, t) for i, t in self.items if len(i) <= cap]
        lens = [len(i) for i, _ in self.items]
        self.samples = pack_conversations(lens, cap) if config.pack else [[i] for i in range(len(lens))]

    @classmethod
    def from_jsonl(cls, path: str, tokenizer, config: SFTDatasetConfig):
        with open(path) as f:
            convs = [json.loads(l)["messages"] for l in f if l.strip()]
        return cls(convs, tokenizer, config)

    def __len__(self) -> int:
        return len(self.samples)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        S = self.cfg.sequence_length
        tokens = np.full(S, self.cfg.pad_token_id, dtype=np.int64)
        labels = np.full(S, IGNORE_INDEX, dtype=np.int64)
        loss_mask = np.zeros(S, dtype=np.float32)
        position_ids = np.zeros(S, dtype=np.int64)
        cu = [0]
        at = 0
        for ci in self.samples[idx]:
            ids, train = self.items[ci]
            n = min(len(ids) - 1, S - at)                     # a document of L tokens yields L-1 (input, label) pairs
            if n <= 0:
                break
            tokens[at : at + n] = ids[:n]
            labels[at : at + n] = ids[1 : n + 1]
            loss_mask[at : at + n] = train[1 : n + 1]         # predict token t+1 → its flag
            position_ids[at : at + n] = np.arange(n)
            at += n
            cu.append(at)
        labels[loss_mask == 0] = IGNORE_INDEX
        cu_seqlens = np.full(len(self.samples[idx]) + 2, at, dtype=np.int32)   # padded tail forms a last (ignored) document
        cu_seqlens[: len(cu)] = cu
        cu_seqlens[-1] = S
        return {"tokens": torch.from_numpy(tokens), "labels": torch.from_numpy(labels), "loss_mask": torch.from_numpy(loss_mask),
                "position_ids": torch.from_numpy(position_ids), "cu_seqlens": torch.from_numpy(cu_seqlens), "max_seqlen": torch.tensor(int(np.diff(cu_seqlens).max()))}
