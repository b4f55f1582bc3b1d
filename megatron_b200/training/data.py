"""Resumable data samplers and loaders (reference ``megatron/training/datasets/data_samplers.py:19,121``)."""
from __future__ import annotations

from typing import Iterator, List

import torch
import torch.distributed as dist

from ..core import parallel_state as ps


class MegatronPretrainingSampler:
    """Sequential, data-parallel-strided, resumable from ``consumed_samples``."""

    def __init__(self, total_samples: int, consumed_samples: int, micro_batch_size: int, data_parallel_rank: int, data_parallel_size: int, drop_last: bool = True):
        assert total_samples > 0 and micro_batch_size > 0 and data_parallel_size > 0
        assert data_parallel_rank < data_parallel_size
        self.total_samples, self.consumed_samples, self.micro_batch_size = total_samples, consumed_samples, micro_batch_size
        self.dp_rank, self.dp_size, self.drop_last = data_parallel_rank, data_parallel_size, drop_last
        self.per_step = micro_batch_size * data_parallel_size

    def __len__(self):
        return self.total_samples

    def __iter__(self) -> Iterator[List[int]]:
        batch = []
        for idx in range(self.consumed_samples % self.total_samples if self.consumed_samples >= self.total_samples else self.consumed_samples, self.total_samples):
            batch.append(idx)
            if len(batch) == self.per_step:
                lo = self.dp_rank * self.micro_batch_size
                yield batch[lo : lo + self.micro_batch_size]
                batch = []
        if batch and not self.drop_last:
            lo = self.dp_rank * self.micro_batch_size
            yield batch[lo : lo + self.micro_batch_size]


class MegatronPretrainingRandomSampler:
    """Epoch-wise shuffled variant; deterministic given (seed, epoch)."""

    def __init__(self, dataset, total_samples, consumed_samples, micro_batch_size, data_parallel_rank, data_parallel_size, data_sharding: bool = True, seed: int = 1234):
        self.total_samples, self.consumed_samples, self.micro_batch_size = total_samples, consumed_samples, micro_batch_size
        self.dp_rank, self.dp_size, self.seed = data_parallel_rank, data_parallel_size, seed
        self.per_step = micro_batch_size * data_parallel_size
        self.last_batch_size = total_samples % self.per_step

    def __len__(self):
        return self.total_samples

    def __iter__(self):
        active = self.total_samples - self.last_batch_size
        epoch = self.consumed_samples // active
        cur = self.consumed_samples % active
        bucket = (active // self.per_step) * self.micro_batch_size
        offset = cur // self.dp_size
        start = self.dp_rank * bucket
        g = torch.Generator()
        g.manual_seed(self.seed + epoch)
        perm = torch.randperm(bucket, generator=g).tolist()
        idxs = [start + x for x in perm[offset:]]
        batch = []
        for i in idxs:
            batch.append(i)
            if len(batch) == self.micro_batch_size:
                self.consumed_samples += self.per_step
                yield batch
                batch = []


def build_pretraining_data_loader(dataset, consumed_samples: int, args, dataloader_type: str = "single"):
    if dataset is None:
        return None
    dp_rank = ps.get_data_parallel_rank() if ps.is_initialized() else 0
    dp_size = ps.get_data_parallel_world_size() if ps.is_initialized() else 1
    if dataloader_type == "cyclic":
        sampler = MegatronPretrainingRandomSampler(dataset, len(dataset), consumed_samples, args.micro_batch_size, dp_rank, dp_size, seed=args.seed)
    else:
        sampler = MegatronPretrainingSampler(len(dataset), consumed_samples, args.micro_batch_size, dp_rank, dp_size)
    # a private generator: creating the iterator draws the worker base seed, which must not advance the global torch RNG — otherwise a resumed run
    # (RNG restored from the checkpoint, then the loader rebuilt) diverges from the uninterrupted one at the first dropout
    gen = torch.Generator()
    gen.manual_seed(int(getattr(args, "seed", 1234)) + 7919 * dp_rank + consumed_samples)
    return torch.utils.data.DataLoader(dataset, batch_sampler=sampler, num_workers=getattr(args, "num_workers", 0), pin_memory=torch.cuda.is_available(),
                                       persistent_workers=getattr(args, "num_workers", 0) > 0, generator=gen)


def get_batch_on_this_tp_rank(data_iterator, keys=("tokens", "labels", "loss_mask", "position_ids"), device=None):
    """TP rank 0 reads the batch and broadcasts it over the TP group (reference ``utils.py:2167``)."""
    from ..core.tensor_parallel.data import broadcast_data

    tp = ps.get_tensor_model_parallel_world_size()
    if tp == 1:
        b = next(data_iterator)
        dev = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and dist.get_backend() != "gloo" else torch.device("cpu"))
        return {k: b[k].to(dev, non_blocking=True) for k in keys if k in b}
    data = next(data_iterator) if ps.get_tensor_model_parallel_rank() == 0 else None
    ints = broadcast_data([k for k in keys if k != "loss_mask"], data, torch.int64)
    out = dict(ints)
    if "loss_mask" in keys:
        out.update(broadcast_data(["loss_mask"], data, torch.float32))
    return out
