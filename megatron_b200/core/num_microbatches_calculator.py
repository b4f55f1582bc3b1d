"""Global-batch → number of micro-batches: constant, step schedule (``"THRESHOLD:BS ..."``, reference ``num_microbatches_calculator.py:402-593``) or the
older linear ramp-up (the reference has deprecated ``rampup_batch_size`` into a no-op; it stays functional here and the step schedule takes precedence)."""
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


def init_num_microbatches_calculator(rank: int, rampup_batch_size: Optional[List[int]] = None, global_batch_size: Optional[int] = None, micro_batch_size: int = 1,
                                     data_parallel_size: int = 1, decrease_batch_size_if_needed: bool = False, step_batch_size_schedule: Optional[str] = None,
                                     seq_length: Optional[int] = None) -> None:
    global _CALC
    assert _CALC is None, "num microbatches calculator is already initialized"
    _CALC = _build(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed, step_batch_size_schedule, seq_length)


def reconfigure_num_microbatches_calculator(rank, rampup_batch_size=None, global_batch_size=None, micro_batch_size=1, data_parallel_size=1,
                                            decrease_batch_size_if_needed=False, step_batch_size_schedule=None, seq_length=None):
    global _CALC
    _CALC = _build(rank, rampup_batch_size, global_batch_size, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed, step_batch_size_schedule, seq_length)


def destroy_num_microbatches_calculator():
    global _CALC
    _CALC = None


def unset_num_microbatches_calculator():
    destroy_num_microbatches_calculator()


def _build(rank, rampup, gbs, mbs, dp, decrease, step_schedule=None, seq_length=None):
    if step_schedule is not None:
        if decrease:
            raise ValueError("Cannot specify both --step-batch-size-schedule and --decrease-batch-size-if-needed")
        return StepBatchsizeNumMicroBatchesCalculator(mbs, dp, decrease, rank, step_schedule, seq_length)
    assert gbs is not None, "--global-batch-size is required when not using --step-batch-size-schedule"
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


_SUFFIX = {"K": 10**3, "M": 10**6, "B": 10**9, "T": 10**12}


class StepBatchsizeNumMicroBatchesCalculator(NumMicroBatchesCalculator):
    """Piece-wise constant global batch size.  ``schedule`` is ``"THRESHOLD:BS THRESHOLD:BS ..."`` (space or comma separated, suffixes K / M / B / T on either
    number); thresholds are consumed TOKENS when ``seq_length`` is given (floor-divided into samples) and consumed samples otherwise; the first threshold
    must be 0.  Divisibility by ``micro_batch_size * data_parallel_size`` is only enforced for the batch size in force when ``update`` runs with
    ``consistency_check`` — earlier entries may predate the current GPU count."""

    def __init__(self, micro_batch_size, data_parallel_size, decrease_batch_size_if_needed, rank, schedule, seq_length=None):
        super().__init__()
        if decrease_batch_size_if_needed:
            raise ValueError("Step batch size schedules do not support decrease_batch_size_if_needed")
        self.micro_batch_size, self.data_parallel_size, self.rank, self.seq_length = micro_batch_size, data_parallel_size, rank, seq_length
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        assert self.micro_batch_times_data_parallel_size > 0
        self.schedule = self._parse_schedule(schedule, seq_length)
        assert self.schedule, "schedule must have at least one entry"
        assert self.schedule[0][0] == 0, f"first schedule entry must have threshold 0, got {self.schedule[0][0]}"
        for (a, _), (b, _) in zip(self.schedule, self.schedule[1:]):
            assert b > a, f"schedule thresholds must be strictly increasing, got {a} before {b}"
        assert all(bs > 0 for _, bs in self.schedule), "batch sizes must be positive"
        self.global_batch_size = self.schedule[-1][1]
        self.update(0, consistency_check=False, verbose=True)

    @staticmethod
    def _parse_numeric_value(text: str) -> int:
        text = text.strip().upper()
        mult = _SUFFIX.get(text[-1:], 1)
        return int(float(text[:-1] if mult != 1 else text) * mult)

    @classmethod
    def _parse_schedule(cls, text: str, seq_length):
        out = []
        for entry in text.strip().replace(",", " ").split():
            if ":" not in entry:
                raise ValueError(f'Invalid schedule entry "{entry}". Expected format: "THRESHOLD:BATCH_SIZE"')
            thr, bs = entry.split(":", 1)
            thr = cls._parse_numeric_value(thr)
            out.append((thr // seq_length if seq_length is not None else thr, cls._parse_numeric_value(bs)))
        return sorted(out, key=lambda e: e[0])

    def _get_batch_size_for_samples(self, consumed_samples: int) -> int:
        bs = self.schedule[0][1]
        for thr, b in self.schedule:
            if consumed_samples < thr:
                break
            bs = b
        return bs

    def update(self, consumed_samples, consistency_check, verbose=False):
        self.current_global_batch_size = self._get_batch_size_for_samples(consumed_samples)
        if consistency_check:
            assert self.current_global_batch_size % self.micro_batch_times_data_parallel_size == 0, (
                f"current global batch size ({self.current_global_batch_size}) is not divisible by micro_batch_size ({self.micro_batch_size}) * "
                f"data_parallel_size ({self.data_parallel_size})")
        self.current_running_global_batch_size = self.current_global_batch_size
        self.num_micro_batches = self.current_running_global_batch_size // self.micro_batch_times_data_parallel_size
