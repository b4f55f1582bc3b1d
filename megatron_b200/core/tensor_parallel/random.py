"""RNG stream tracker and activation checkpointing.

Parity: reference ``tensor_parallel/random.py`` — ``CudaRNGStatesTracker`` :229,
``model_parallel_cuda_manual_seed`` :446 (seed offsets 2718 + tp_rank for the
model-parallel stream, 1024 + 100*ep + etp for the expert stream),
``checkpoint`` :675, ``CheckpointWithoutOutput`` :848.

Design difference: streams are device-agnostic ``torch.Generator`` *states*
(works on CPU for the gloo tests and on CUDA), and the output-discarding
checkpoint re-materialises by resizing the output's storage in place
(``untyped_storage().resize_``) rather than through a C++ storage-aliasing
extension (reference N4).
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, Optional

import torch
from torch.utils.checkpoint import detach_variable

from .. import parallel_state as ps

_MODEL_PARALLEL_RNG_TRACKER_NAME = "model-parallel-rng"
_EXPERT_PARALLEL_RNG_TRACKER_NAME = "expert-parallel-rng"
_DATA_PARALLEL_RNG_TRACKER_NAME = "data-parallel-rng"


def _dev_is_cuda() -> bool:
    return torch.cuda.is_available()


def _get_state():
    return torch.cuda.get_rng_state() if _dev_is_cuda() else torch.get_rng_state()


def _set_state(state):
    if _dev_is_cuda():
        torch.cuda.set_rng_state(state)
    else:
        torch.set_rng_state(state)


def _seed(seed: int):
    if _dev_is_cuda():
        torch.cuda.manual_seed(seed)
    else:
        torch.manual_seed(seed)


class CudaRNGStatesTracker:
    """Named RNG states; ``fork(name)`` temporarily swaps the default generator."""

    def __init__(self, use_cudagraphable_rng: bool = False, is_inference_rng_tracker: bool = False):
        self.use_cudagraphable_rng = use_cudagraphable_rng
        self.is_inference_rng_tracker = is_inference_rng_tracker
        self.reset()

    def is_initialized(self) -> bool:
        return self._is_initialized

    def reset(self):
        self._is_initialized = False
        self.states_: Dict[str, torch.Tensor] = {}
        self.seeds_ = set()

    def get_states(self):
        return dict(self.states_)

    def set_states(self, states):
        self._is_initialized = True
        self.states_ = states

    def add(self, name: str, seed: int):
        self._is_initialized = True
        if seed in self.seeds_:
            raise Exception(f"seed {seed} already exists")
        if name in self.states_:
            raise Exception(f"rng state {name} already exists")
        self.seeds_.add(seed)
        saved = _get_state()
        _seed(seed)
        self.states_[name] = _get_state()
        _set_state(saved)

    @contextlib.contextmanager
    def fork(self, name: str = _MODEL_PARALLEL_RNG_TRACKER_NAME):
        if name not in self.states_:
            raise Exception(f"rng state {name} is not added")
        saved = _get_state()
        _set_state(self.states_[name])
        try:
            yield
        finally:
            self.states_[name] = _get_state()
            _set_state(saved)


_TRACKER: Optional[CudaRNGStatesTracker] = None


def initialize_rng_tracker(use_te_rng_tracker=False, inference_rng_tracker=False, use_cudagraphable_rng=False, force_reset=False):
    global _TRACKER
    if force_reset or _TRACKER is None:
        _TRACKER = CudaRNGStatesTracker(use_cudagraphable_rng, inference_rng_tracker)


def get_cuda_rng_tracker(use_te_rng_tracker=False, inference_rng_tracker=False, use_cudagraphable_rng=False):
    initialize_rng_tracker(use_te_rng_tracker, inference_rng_tracker, use_cudagraphable_rng)
    return _TRACKER


def get_all_rng_states():
    return get_cuda_rng_tracker().get_states()


def get_data_parallel_rng_tracker_name():
    return _DATA_PARALLEL_RNG_TRACKER_NAME


def get_expert_parallel_rng_tracker_name():
    return _EXPERT_PARALLEL_RNG_TRACKER_NAME


def model_parallel_cuda_manual_seed(
    seed: int,
    te_rng_tracker: bool = False,
    inference_rng_tracker: bool = False,
    use_cudagraphable_rng: bool = False,
    tp_rank: Optional[int] = None,
    ep_rank: Optional[int] = None,
    etp_rank: Optional[int] = None,
    force_reset_rng: bool = False,
):
    """Default stream = same on all TP ranks (data-parallel RNG);
    model-parallel stream differs per TP rank; expert stream per (ep, etp)."""
    tp_rank = ps.get_tensor_model_parallel_rank() if tp_rank is None else tp_rank
    ep_rank = ps.get_expert_model_parallel_rank() if ep_rank is None else ep_rank
    etp_rank = ps.get_expert_tensor_parallel_rank() if etp_rank is None else etp_rank
    initialize_rng_tracker(te_rng_tracker, inference_rng_tracker, use_cudagraphable_rng, force_reset=force_reset_rng)
    tracker = get_cuda_rng_tracker()
    tracker.reset()
    _seed(seed)
    tracker.add(_DATA_PARALLEL_RNG_TRACKER_NAME, seed)
    tracker.add(_MODEL_PARALLEL_RNG_TRACKER_NAME, seed + 2718 + tp_rank)
    tracker.add(_EXPERT_PARALLEL_RNG_TRACKER_NAME, seed + 1024 + 100 * ep_rank + etp_rank)


def model_parallel_reconfigure_tp_seed(seed: int):
    model_parallel_cuda_manual_seed(seed, force_reset_rng=True)


# -----------------------------------------------------------------------------
# activation checkpointing
# -----------------------------------------------------------------------------


def _snapshot_rng():
    return torch.get_rng_state(), (_get_state() if _dev_is_cuda() else None), get_cuda_rng_tracker().get_states()


def _restore_rng(snap):
    cpu, dev, tracker = snap
    torch.set_rng_state(cpu)
    if dev is not None:
        _set_state(dev)
    get_cuda_rng_tracker().set_states(tracker)


class CheckpointFunction(torch.autograd.Function):
    """Full recompute of ``run_function`` in backward with RNG replay.

    ``distribute_saved_activations`` shards the saved input over the TP group
    (1/tp of the memory) and all-gathers it back before recompute.
    """

    @staticmethod
    def forward(ctx, run_function, distribute_saved_activations, *args):
        ctx.run_function = run_function
        ctx.distribute = distribute_saved_activations
        ctx.rng = _snapshot_rng()
        with torch.no_grad():
            outputs = run_function(*args)
        if distribute_saved_activations:
            from .utils import split_tensor_into_1d_equal_chunks

            ctx.input_0_shape = args[0].shape
            args = (split_tensor_into_1d_equal_chunks(args[0].detach(), new_buffer=True),) + tuple(args[1:])
        ctx.save_for_backward(*[a if torch.is_tensor(a) else None for a in args])
        ctx.non_tensors = [None if torch.is_tensor(a) else a for a in args]
        return outputs

    @staticmethod
    def backward(ctx, *grads):
        if not torch.autograd._is_checkpoint_valid():
            raise RuntimeError("checkpointing is not compatible with .grad(); use .backward()")
        saved = list(ctx.saved_tensors)
        inputs = [s if s is not None else nt for s, nt in zip(saved, ctx.non_tensors)]
        if ctx.distribute:
            from .utils import gather_split_1d_tensor

            inputs[0] = gather_split_1d_tensor(inputs[0]).view(ctx.input_0_shape)
        now = _snapshot_rng()
        _restore_rng(ctx.rng)
        detached = detach_variable(tuple(inputs))
        with torch.enable_grad():
            outputs = ctx.run_function(*detached)
        _restore_rng(now)
        if isinstance(outputs, torch.Tensor):
            outputs = (outputs,)
        pairs = [(o, g) for o, g in zip(outputs, grads) if torch.is_tensor(o) and o.requires_grad and g is not None]
        if pairs:
            torch.autograd.backward([p[0] for p in pairs], [p[1] for p in pairs])
        return (None, None) + tuple(x.grad if isinstance(x, torch.Tensor) else None for x in detached)


def checkpoint(function: Callable, distribute_saved_activations: bool, *args):
    return CheckpointFunction.apply(function, distribute_saved_activations, *args)


class CheckpointWithoutOutput:
    """Checkpoint a cheap op (norm, activation) *and* free its output.

    ``checkpoint(fn, *args)`` runs ``fn`` without graph; after the consumer has
    saved the output for its own backward, ``discard_output_and_register_recompute
    (hook_tensor)`` frees the output storage and arranges for it to be recomputed
    right before ``hook_tensor``'s gradient is produced.
    """

    def __init__(self, fp8: bool = False):
        self.run_function = None
        self.rng = None
        self.ctx_inputs = None
        self.outputs = None

    def checkpoint(self, run_function: Callable, *args):
        self.run_function = run_function
        self.rng = _snapshot_rng()
        self.ctx_inputs = args
        outputs = _CheckpointWithoutOutputFn.apply(run_function, self, *args)
        self.outputs = outputs if isinstance(outputs, tuple) else (outputs,)
        return outputs

    def _recompute(self, _grad):
        if self.ctx_inputs is None:
            return
        now = _snapshot_rng()
        _restore_rng(self.rng)
        with torch.enable_grad():
            detached = detach_variable(tuple(self.ctx_inputs))
            outs = self.run_function(*detached)
        _restore_rng(now)
        outs = outs if isinstance(outs, tuple) else (outs,)
        for kept, fresh in zip(self.outputs, outs):
            # refill the freed storage through `.data` (separate version counter) so autograd's
            # saved-tensor version check still sees the tensor as unmodified
            kept.untyped_storage().resize_(kept.numel() * kept.element_size())
            kept.data.copy_(fresh.detach().reshape(kept.shape))
        self._recomputed = (detached, outs)

    def discard_output_and_register_recompute(self, hook_tensor: torch.Tensor):
        for o in self.outputs:
            o.untyped_storage().resize_(0)
        if hook_tensor.requires_grad:
            hook_tensor.register_hook(self._recompute)


class _CheckpointWithoutOutputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, run_function, owner, *args):
        ctx.owner = owner
        with torch.no_grad():
            out = run_function(*args)
        ctx.n_args = len(args)
        return out

    @staticmethod
    def backward(ctx, *grads):
        owner = ctx.owner
        if getattr(owner, "_recomputed", None) is None:
            owner._recompute(None)
        detached, outs = owner._recomputed
        pairs = [(o, g) for o, g in zip(outs, grads) if torch.is_tensor(o) and o.requires_grad and g is not None]
        if pairs:
            torch.autograd.backward([p[0] for p in pairs], [p[1] for p in pairs])
        grads_in = tuple(x.grad if isinstance(x, torch.Tensor) else None for x in detached)
        # Break the cycle owner → outputs → tensor → grad_fn(ctx) → owner: it runs through a C++
        # autograd node, so Python's GC cannot collect it and the recomputed activations would
        # otherwise stay alive forever (measured: 352 MiB / layer / micro-batch on Llama-3 8B).
        for o in owner.outputs or ():
            o.untyped_storage().resize_(0)
        owner.ctx_inputs = owner._recomputed = owner.outputs = owner.run_function = None
        ctx.owner = None
        return (None, None) + grads_in
