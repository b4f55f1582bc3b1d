"""Global-batch → number of micro-batches, constant or ramp-up
(reference ``num_microbatches_calculator.py:17-593``)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional

_CALC: Optional["NumMicroBatchesCalculator"] = None


def get_num_microbatches() -> int:
    return _CALC.get()


def get_current_global_batch_size() -> int:
    return _CALC.get_current_global_batch_size()


def get_micro_batch_size() -> int:
    return _CALC.micro_batch_size


def get_current_running_global_batch_size() -> int:
    return _CALC.get_current_running_global_batch_size()


def update_num_microbatches(consumed_samples: int, consistency_check: bool = True, verbose: bool = False) -> None:
    _CALC.update(consumed_samples, consistency_check, verbose)


def init_num_microbatches_calculator(rank: int, rampup_batch_size: Optional[List[int]], global_batch_size: int, micro_batch_size: int,
                                     data_parallel_size: int, decrease_batch_size_if_needed: bool = False) -> None:
    global _CALC
    assert _CALC is None, "num microbatches calculator is already initialized"
    _CALC = _build(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed)


def reconfigure_num_microbatches_calculator(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed=False):
    global _CALC
    _CALC = _build(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed)


def destroy_num_microbatches_calculator():
    global _CALC
    _CALC = None


def unset_num_microbatches_calculator():
    destroy_num_microbatches_calculator()


def _build(rank, rampup, gbs, mbs, dp, decrease):
    if rampup is None:
        return ConstantNumMicroBatchesCalculator(gbs, mbs, dp, decrease, rank)
    assert len(rampup) == 3, "expected: start global batch size, batch size increment, ramp-up samples"
    return RampupBatchsizeNumMicroBatchesCalculator(gbs, mbs, dp, decrease, rank, int(rampup[0]), int(rampup[1]), int(rampup[2]))


def _round(batch_size: int, divisor: int) -> int:
    return (batch_size // divisor) * divisor


class NumMicroBatchesCalculator(ABC):
    def __init__(self):
        self.num_micro_batches = None
        self.current_global_batch_size = None
        self.micro_batch_size = None
        self.current_running_global_batch_size = None

    def get(self) -> int:
        return self.num_micro_batches

    def get_current_global_batch_size(self) -> int:
        return self.current_global_batch_size

    def get_micro_batch_size(self) -> int:
        return self.micro_batch_size

    def get_current_running_global_batch_size(self) -> int:
        return self.current_running_global_batch_size

    @abstractmethod
    def update(self, consumed_samples, consistency_check, verbose=False):
        ...


class ConstantNumMicroBatchesCalculator(NumMicroBatchesCalculator):
    def __init__(self, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed, rank):
        super().__init__()
        per = micro_batch_size * data_parallel_size
        if decrease_batch_size_if_needed:
            running = _round(global_batch_size, per)
            assert running % per == 0
            self.num_micro_batches = running // per
        else:
            assert global_batch_size % per == 0, f"global batch size ({global_batch_size}) is not divisible by micro batch size ({micro_batch_size}) times data parallel size ({data_parallel_size})"
            running = global_batch_size
            self.num_micro_batches = global_batch_size // per
        assert self.num_micro_batches >= 1
        self.current_global_batch_size = global_batch_size
        self.current_running_global_batch_size = running
        self.micro_batch_size = micro_batch_size

    def update(self, consumed_samples, consistency_check, verbose=False):
        pass


class RampupBatchsizeNumMicroBatchesCalculator(NumMicroBatchesCalculator):
    """Linear ramp: start at ``start_global_batch_size``, add ``batch_size_increment`` every
    ``ramup_samples / steps`` consumed samples until ``global_batch_size``."""

    def __init__(self, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed, rank, start_global_batch_size,
                 batch_size_increment, ramup_samples):
        super().__init__()
        assert global_batch_size > 0 and start_global_batch_size > 0 and batch_size_increment > 0 and ramup_samples >= 0
        self.global_batch_size, self.micro_batch_size, self.data_parallel_size = global_batch_size, micro_batch_size, data_parallel_size
        self.decrease_batch_size_if_needed, self.rank = decrease_batch_size_if_needed, rank
        self.start_global_batch_size, self.batch_size_increment, self.ramup_samples = start_global_batch_size, batch_size_increment, ramup_samples
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        diff = global_batch_size - start_global_batch_size
        assert diff >= 0 and diff % batch_size_increment == 0, "global batch size interval must be divisible by the increment"
        steps = diff // batch_size_increment
        self.rampup_samples_per_increment = ramup_samples / steps if steps else 0
        self.update(0, consistency_check=False)

    def update(self, consumed_samples, consistency_check, verbose=False):
        if consumed_samples > self.ramup_samples or self.rampup_samples_per_increment == 0:
            self.current_global_batch_size = self.global_batch_size
        else:
            steps = int(consumed_samples / self.rampup_samples_per_increment)
            self.current_global_batch_size = min(self.global_batch_size, self.start_global_batch_size + steps * self.batch_size_increment)
        per = self.micro_batch_times_data_parallel_size
        if consistency_check and not self.decrease_batch_size_if_needed:
            assert self.current_global_batch_size % per == 0, "current global batch size is not divisible by micro-batch-size times data parallel size"
        self.current_running_global_batch_size = _round(self.current_global_batch_size, per) if self.decrease_batch_size_if_needed else self.current_global_batch_size
        self.num_micro_batches = max(1, self.current_running_global_batch_size // per)
