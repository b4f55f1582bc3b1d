"""Fault attribution by re-running a suspicious step (reference ``rerun_state_machine.py:129-1420``).

The training loop wraps forward-backward in ``while rsm.should_run_forward_backward(data_iterator)``.
When ``validate_result`` flags a value (NaN / Inf / spike) the step is re-run IN PLACE on the same data
(``RerunDataIterator`` replays the batches, RNG restored).  Same bad value again ⇒ deterministic ⇒ save a
checkpoint and exit with ``EXIT_CODE_RESUME_TO_DISAMBIGUATE`` so the scheduler can re-run on different
hardware (persistent fault vs. correct-but-unexpected result); a different value ⇒ transient fault.
"""
from __future__ import annotations

import logging
import math
import random
from collections import defaultdict
from enum import Enum
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)

EXIT_CODE_RESUME_TO_DISAMBIGUATE = 16
EXIT_CODE_FAILED_ON_RESULT_VALIDATION = 17


class RerunDiagnostic(str, Enum):
    CORRECT_RESULT = "correct_result"
    TRANSIENT_ERROR = "transient_error"
    PERSISTENT_ERROR = "persistent_error"


class RerunMode(str, Enum):
    DISABLED = "disabled"
    VALIDATE_RESULTS = "validate_results"
    REPORT_DETERMINISM_STATS = "report_stats"


class RerunState(Enum):
    NOT_RUNNING_YET = 0
    INITIAL_RUN = 1
    RERUNNING_IN_PLACE = 2
    WILL_RERUN_FROM_CHECKPOINT = 3
    RERUNNING_FROM_CHECKPOINT = 4
    RERUNNING_AGAIN_FROM_CHECKPOINT = 5


class RerunDataIterator:
    """Records the batches of the current step so the step can be replayed."""

    def __init__(self, iterable: Iterable[Any]):
        self.iterable = iterable
        self.saved: List[Any] = []
        self.replaying = False
        self.replay_pos = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.replaying:
            if self.replay_pos < len(self.saved):
                x = self.saved[self.replay_pos]
                self.replay_pos += 1
                return x
            self.replaying = False
        x = next(self.iterable)
        self.saved.append(x)
        return x

    def rewind(self):
        self.replaying, self.replay_pos = True, 0

    def advance(self):
        self.saved, self.replaying, self.replay_pos = [], False, 0

    def state_dict(self):
        return {"saved": self.saved, "replaying": self.replaying, "replay_pos": self.replay_pos}

    def load_state_dict(self, sd):
        self.saved, self.replaying, self.replay_pos = sd["saved"], sd["replaying"], sd["replay_pos"]


class RerunErrorInjector:
    """Deterministic fault injection for testing the machine (reference :1267-1364)."""

    def __init__(self, error_injection_rate: int = 0, error_injection_type: str = "transient_error"):
        self.rate, self.kind = error_injection_rate, error_injection_type
        self.should_inject = False
        self.injected = False
        self.calls = 0

    def maybe_inject(self) -> bool:
        if self.rate <= 0:
            return False
        self.calls += 1
        if self.kind == "persistent_error":
            return self.calls % self.rate == 0 or self.injected and self.should_inject
        if self.calls % self.rate == 0 and not self.injected:
            self.injected = True
            return True
        return False

    def maybe_miscompare(self, comparison_func, a, b, rerun_state):
        return comparison_func(a, b)


def _rng_snapshot():
    return {"py": random.getstate(), "np": np.random.get_state(), "torch": torch.get_rng_state(),
            "cuda": torch.cuda.get_rng_state() if torch.cuda.is_available() else None}


def _rng_restore(s):
    random.setstate(s["py"])
    np.random.set_state(s["np"])
    torch.set_rng_state(s["torch"])
    if s["cuda"] is not None:
        torch.cuda.set_rng_state(s["cuda"])


class RerunStateMachine:
    def __init__(self, mode: RerunMode = RerunMode.DISABLED, error_injector: Optional[RerunErrorInjector] = None,
                 state_save_func: Optional[Callable] = None, state_restore_func: Optional[Callable] = None):
        self.mode = RerunMode(mode)
        self.state = RerunState.NOT_RUNNING_YET
        self.error_injector = error_injector or RerunErrorInjector()
        self.state_save_func, self.state_restore_func = state_save_func, state_restore_func
        self.rerun_requested = False
        self.checkpoint_requested = False
        self.restart_again_requested = False
        self.continue_requested = False
        self.failed_validation_call = None
        self.initial_result = None
        self.suspicious_node = None
        self.saved_rng = None
        self.saved_user_state = None
        self.data_iterators: List[RerunDataIterator] = []
        self.validation_counts: Dict[str, int] = defaultdict(int)
        self.stats: Dict[str, List[float]] = defaultdict(list)
        self.current_iteration = -1
        self.last_diagnostic: Optional[RerunDiagnostic] = None

    # ---- loop control -----------------------------------------------------------------------------------
    def set_mode(self, mode):
        self.mode = RerunMode(mode)

    def get_mode(self):
        return self.mode

    def _collect(self, data_iterator):
        its = data_iterator if isinstance(data_iterator, list) else [data_iterator]
        return [i for i in its if isinstance(i, RerunDataIterator)]

    def _any_rank(self, flag: bool) -> bool:
        if dist.is_available() and dist.is_initialized():
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([1 if flag else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return bool(t.item())
        return flag

    def should_run_forward_backward(self, data_iterator=None) -> bool:
        """Call in a ``while`` loop around forward-backward: True the first time each step, True again
        when a re-run in place was requested, False when the step is accepted."""
        self.validation_counts.clear()
        if self.mode == RerunMode.DISABLED:
            if self.state == RerunState.NOT_RUNNING_YET:
                self.state = RerunState.INITIAL_RUN
                return True
            self.state = RerunState.NOT_RUNNING_YET
            return False
        if self.state == RerunState.NOT_RUNNING_YET:
            self.state = RerunState.INITIAL_RUN
            self.current_iteration += 1
            self.data_iterators = self._collect(data_iterator)
            for it in self.data_iterators:
                it.advance()
            self.saved_rng = _rng_snapshot()
            if self.state_save_func is not None:
                self.saved_user_state = self.state_save_func()
            self.rerun_requested = self.checkpoint_requested = self.continue_requested = False
            return True
        if self.state == RerunState.INITIAL_RUN:
            if self._any_rank(self.rerun_requested):
                self.state = RerunState.RERUNNING_IN_PLACE
                for it in self.data_iterators:
                    it.rewind()
                _rng_restore(self.saved_rng)
                if self.state_restore_func is not None:
                    self.state_restore_func(self.saved_user_state)
                self.rerun_requested = False
                logger.warning("rerun state machine: re-running iteration %d in place", self.current_iteration)
                return True
            self.state = RerunState.NOT_RUNNING_YET
            return False
        if self.state == RerunState.RERUNNING_IN_PLACE:
            self.state = RerunState.NOT_RUNNING_YET if not self.checkpoint_requested else RerunState.WILL_RERUN_FROM_CHECKPOINT
            return False
        if self.state in (RerunState.RERUNNING_FROM_CHECKPOINT, RerunState.RERUNNING_AGAIN_FROM_CHECKPOINT):
            self.state = RerunState.NOT_RUNNING_YET
            return False
        self.state = RerunState.NOT_RUNNING_YET
        return False

    def should_checkpoint_and_exit(self) -> Tuple[bool, bool, int]:
        """(save a checkpoint?, exit?, exit code) — to be queried after the while loop."""
        if self.mode == RerunMode.DISABLED:
            return False, False, 0
        if self._any_rank(self.checkpoint_requested):
            self.checkpoint_requested = False
            return True, True, EXIT_CODE_RESUME_TO_DISAMBIGUATE
        if self._any_rank(self.restart_again_requested):
            self.restart_again_requested = False
            return False, True, EXIT_CODE_FAILED_ON_RESULT_VALIDATION
        return False, False, 0

    # ---- validation ---------------------------------------------------------------------------------------
    def validate_result(self, result: Any, rejection_func: Callable[[Any], bool], message: str = "unexpected result", comparison_func: Optional[Callable] = None,
                        tolerance: float = 0.0, fatal: bool = True) -> None:
        """Flag ``result`` if ``rejection_func(result)``; drive the rerun protocol accordingly."""
        if self.mode == RerunMode.DISABLED:
            if rejection_func(result) and fatal:
                raise RuntimeError(message)
            return
        key = message
        self.validation_counts[key] += 1
        val = float(result) if isinstance(result, (int, float)) or (torch.is_tensor(result) and result.numel() == 1) else result
        cmp = comparison_func or (lambda a, b: abs(a - b) / max(abs(a), abs(b), 1e-12) if (math.isfinite(a) and math.isfinite(b)) else (0.0 if (str(a) == str(b)) else float("inf")))
        if self.mode == RerunMode.REPORT_DETERMINISM_STATS:
            if self.state == RerunState.INITIAL_RUN:
                self.initial_result = val
                self.rerun_requested = True
            elif self.state == RerunState.RERUNNING_IN_PLACE:
                self.stats[key].append(cmp(self.initial_result, val))
            return
        injected = self.error_injector.maybe_inject() if self.state == RerunState.INITIAL_RUN else False
        if self.state == RerunState.INITIAL_RUN:
            if rejection_func(result) or injected:
                self.failed_validation_call = (key, self.validation_counts[key])
                self.initial_result = float("nan") if injected and not rejection_func(result) else val
                self.rerun_requested = True
                logger.error("rerun state machine: %s (%s) at iteration %d — scheduling an in-place re-run", message, val, self.current_iteration)
        elif self.state == RerunState.RERUNNING_IN_PLACE:
            if self.failed_validation_call != (key, self.validation_counts[key]):
                return
            same = cmp(self.initial_result, val) <= tolerance
            if same:
                # reproducible: could be a persistent HW fault or a genuinely bad (but correct) value → disambiguate elsewhere
                self.last_diagnostic = None
                self.checkpoint_requested = True
                logger.error("rerun: result reproduced (%s); checkpointing to re-run on different hardware", val)
            else:
                self.last_diagnostic = RerunDiagnostic.TRANSIENT_ERROR
                logger.error("rerun: result NOT reproduced (%s vs %s) ⇒ transient error on this rank", self.initial_result, val)
                if fatal and rejection_func(result):
                    self.restart_again_requested = True
        elif self.state == RerunState.RERUNNING_FROM_CHECKPOINT:
            same = cmp(self.initial_result, val) <= tolerance
            self.last_diagnostic = RerunDiagnostic.CORRECT_RESULT if same else RerunDiagnostic.PERSISTENT_ERROR
            logger.error("rerun from checkpoint on other hardware: %s", self.last_diagnostic.value)

    def is_unexpectedly_large(self, result: float, threshold: float, context: str, num_samples: int = 100, resample: bool = False) -> bool:
        """Spike detector against a running max of the first ``num_samples`` observations."""
        hist = self.stats["_spike_" + context]
        v = float(result)
        if len(hist) < num_samples or resample:
            hist.append(v)
            return False
        return v > threshold * max(hist)

    def get_determinism_stats(self) -> Dict[str, Dict[str, float]]:
        return {k: {"max_rel_diff": max(v), "mean_rel_diff": sum(v) / len(v), "n": len(v)} for k, v in self.stats.items() if v and not k.startswith("_spike_")}

    # ---- checkpoint ------------------------------------------------------------------------------------------
    def state_dict(self, data_iterator=None, ckpt_format: Any = None) -> Optional[dict]:
        if self.mode == RerunMode.DISABLED:
            return None
        return {"mode": self.mode.value, "state": RerunState.WILL_RERUN_FROM_CHECKPOINT.value if self.checkpoint_requested else self.state.value,
                "initial_result": self.initial_result, "failed_validation_call": self.failed_validation_call, "current_iteration": self.current_iteration,
                "rng": self.saved_rng, "data": [it.state_dict() for it in self.data_iterators]}

    def load_state_dict(self, sd: Optional[dict]) -> None:
        if not sd:
            return
        if sd["state"] == RerunState.WILL_RERUN_FROM_CHECKPOINT.value:
            self.state = RerunState.RERUNNING_FROM_CHECKPOINT
            self.initial_result, self.failed_validation_call = sd["initial_result"], sd["failed_validation_call"]
            self.current_iteration = sd["current_iteration"]
            if sd.get("rng"):
                _rng_restore(sd["rng"])


_MACHINE: Optional[RerunStateMachine] = None


def initialize_rerun_state_machine(mode="disabled", error_injection_rate: int = 0, error_injection_type: str = "transient_error", **kw) -> RerunStateMachine:
    global _MACHINE
    _MACHINE = RerunStateMachine(RerunMode(mode), RerunErrorInjector(error_injection_rate, error_injection_type), **kw)
    return _MACHINE


def get_rerun_state_machine() -> RerunStateMachine:
    global _MACHINE
    if _MACHINE is None:
        _MACHINE = RerunStateMachine()
    return _MACHINE


def destroy_rerun_state_machine():
    global _MACHINE
    _MACHINE = None
