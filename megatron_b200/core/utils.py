"""Small shared helpers (reference: ``megatron/core/utils.py``)."""
from __future__ import annotations

import functools
import logging
import math
import operator
import time
import warnings
from contextlib import nullcontext
from functools import reduce
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)

try:
    from packaging.version import Version as PkgVersion
except Exception:  # pragma: no cover
    PkgVersion = None


def is_torch_min_version(version: str, check_equality: bool = True) -> bool:
    cur = PkgVersion(torch.__version__.split("+")[0])
    ref = PkgVersion(version)
    return cur >= ref if check_equality else cur > ref


def is_te_min_version(version, check_equality=True) -> bool:
    """TransformerEngine is never used by this framework."""
    return False


def ensure_divisibility(numerator: int, denominator: int) -> None:
    assert numerator % denominator == 0, f"{numerator} is not divisible by {denominator}"


def divide(numerator: int, denominator: int) -> int:
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


def get_attr_wrapped_model(model, attr, allow_none=True, return_model_obj=False):
    """Walk ``.module`` wrappers (DDP, Float16Module) until ``attr`` is found."""
    if isinstance(model, list):
        raise RuntimeError("get_attr_wrapped_model expects a single model, not a list of chunks")
    while not hasattr(model, attr):
        if not hasattr(model, "module"):
            if allow_none:
                return None
            raise RuntimeError(f"couldn't find attribute {attr}")
        model = model.module
    return model if return_model_obj else getattr(model, attr)


def get_model_config(model):
    return get_attr_wrapped_model(model, "config", allow_none=False)


def get_model_type(model):
    return get_attr_wrapped_model(model, "model_type", allow_none=True)


def get_model_xattn(model):
    try:
        return get_attr_wrapped_model(model, "xattn_needed", allow_none=False)
    except RuntimeError:
        return False


def unwrap_model(model, module_instances=None):
    """Strip wrapper modules (anything exposing ``.module``)."""
    return_list = isinstance(model, list)
    models = model if return_list else [model]
    out = []
    for m in models:
        if module_instances is None:
            while hasattr(m, "module"):
                m = m.module
        else:
            while isinstance(m, module_instances):
                m = m.module
        out.append(m)
    return out if return_list else out[0]


class GlobalMemoryBuffer:
    """Named, reusable scratch tensors (reference ``utils.py:727``).

    Used for the sequence-parallel all-gather destination so that the gathered
    activation is never a fresh allocation.  On GPU ranks the symmetric-heap
    variant (``megatron_b200.parallel.symm.SymmetricHeap``) supersedes this for
    buffers that peers must address.
    """

    def __init__(self):
        self.buffer: Dict[Tuple[str, torch.dtype], torch.Tensor] = {}

    def get_tensor(self, tensor_shape, dtype, name, mem_alloc_context: Optional[Callable] = None, device=None):
        required_len = reduce(operator.mul, tensor_shape, 1)
        key = (name, dtype)
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else "cpu"
        cur = self.buffer.get(key)
        if cur is None or cur.numel() < required_len or cur.device != torch.empty(0, device=device).device:
            ctx = mem_alloc_context() if mem_alloc_context else nullcontext()
            with ctx:
                cur = torch.empty(required_len, dtype=dtype, device=device, requires_grad=False)
            self.buffer[key] = cur
        return cur[0:required_len].view(*tensor_shape)


def _kernel_make_viewless_tensor(inp, requires_grad):
    out = torch.empty((1,), dtype=inp.dtype, device=inp.device, requires_grad=requires_grad)
    out.data = inp.data
    return out


class MakeViewlessTensor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, requires_grad):
        return _kernel_make_viewless_tensor(inp, requires_grad)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


def make_viewless_tensor(inp, requires_grad, keep_graph):
    """Detach the ``._base`` reference so a pipeline output can be freed early."""
    if inp._base is None:
        return inp
    if keep_graph:
        return MakeViewlessTensor.apply(inp, requires_grad)
    return _kernel_make_viewless_tensor(inp, requires_grad)


def assert_viewless_tensor(tensor, extra_msg=None):
    if isinstance(tensor, list):
        [assert_viewless_tensor(t) for t in tensor]
        return tensor
    if not isinstance(tensor, torch.Tensor):
        return tensor
    assert tensor._base is None, f"expected a viewless tensor. {extra_msg or ''}"
    return tensor


def safely_set_viewless_tensor_data(tensor, new_data_tensor):
    assert_viewless_tensor(tensor)
    tensor.data = new_data_tensor


def init_method_normal(sigma: float):
    def init_(tensor):
        return torch.nn.init.normal_(tensor, mean=0.0, std=sigma)

    return init_


def scaled_init_method_normal(sigma: float, num_layers: int, multiplier: float = 2.0):
    std = sigma / math.sqrt(multiplier * num_layers)

    def init_(tensor):
        return torch.nn.init.normal_(tensor, mean=0.0, std=std)

    return init_


def init_method_constant(val: float):
    def init_(tensor):
        return torch.nn.init.constant_(tensor, val)

    return init_


def log_single_rank(lg: logging.Logger, *args, rank: int = 0, **kwargs):
    if dist.is_available() and dist.is_initialized():
        if dist.get_rank() == rank:
            lg.log(*args, **kwargs)
    else:
        lg.log(*args, **kwargs)


def log_on_each_pipeline_stage(lg: logging.Logger, *args, **kwargs):
    from . import parallel_state as ps

    if ps.get_data_parallel_rank(with_context_parallel=True) == 0 and ps.get_tensor_model_parallel_rank() == 0:
        lg.log(*args, **kwargs)


def get_pg_size(group=None) -> int:
    if not dist.is_available() or not dist.is_initialized():
        return 1
    if group is None:
        return dist.get_world_size()
    return dist.get_world_size(group=group)


def get_pg_rank(group=None) -> int:
    if not dist.is_available() or not dist.is_initialized():
        return 0
    if group is None:
        return dist.get_rank()
    return dist.get_rank(group=group)


def get_pg_src_rank(group=None) -> int:
    if group is None or not dist.is_initialized():
        return 0
    return dist.get_process_group_ranks(group)[0]


def get_tensor_model_parallel_group_if_none(tp_group, is_expert=False, check_initialized=True):
    from . import parallel_state as ps

    if tp_group is not None:
        return tp_group
    if not dist.is_available() or not dist.is_initialized():
        return None
    if is_expert:
        return ps.get_expert_tensor_parallel_group(check_initialized=check_initialized)
    return ps.get_tensor_model_parallel_group(check_initialized=check_initialized)


def local_multi_tensor_l2_norm(tensor_lists, per_tensor=False):
    """Pure-torch l2 norm over a list of tensors (CPU fallback of ops.multi_tensor_l2norm)."""
    ts = tensor_lists[0] if tensor_lists and isinstance(tensor_lists[0], (list, tuple)) else tensor_lists
    if not ts:
        return torch.zeros(1), None
    norms = torch.stack([torch.linalg.vector_norm(t.float()) for t in ts])
    return torch.linalg.vector_norm(norms).reshape(1), (norms if per_tensor else None)


def local_multi_tensor_scale(tensor_lists, scale):
    src, dst = tensor_lists
    for s, d in zip(src, dst):
        d.copy_(s * scale)


def check_param_hashes_across_dp_replicas(model: List[torch.nn.Module], cross_check: bool = False) -> bool:
    """Replica-consistency check (reference ``utils.py:936``): hash params, compare over DP."""
    from . import parallel_state as ps

    group = ps.get_data_parallel_group()
    ws = dist.get_world_size(group=group)
    ok = True
    for chunk in model:
        for name, p in chunk.named_parameters():
            data = p.detach().float().cpu().contiguous()
            h = torch.tensor(
                [float(data.double().sum()), float(data.double().abs().sum()), float(data.numel())],
                dtype=torch.float64,
            )
            dev = p.device if dist.get_backend(group) == "nccl" else "cpu"
            hs = [torch.empty_like(h, device=dev) for _ in range(ws)]
            dist.all_gather(hs, h.to(dev), group=group)
            for other in hs:
                if not torch.equal(other.cpu(), h):
                    ok = False
    return ok


def get_batch_on_this_cp_rank(batch: Dict[str, Any], cp_size: Optional[int] = None, cp_rank: Optional[int] = None):
    """Load-balanced (zig-zag) sequence split for causal context parallelism.

    Sequence is cut into ``2*cp`` chunks; rank ``r`` keeps chunks ``r`` and
    ``2cp-1-r`` (reference ``utils.py:2504-2563``).
    """
    from . import parallel_state as ps

    cp_size = ps.get_context_parallel_world_size() if cp_size is None else cp_size
    if cp_size <= 1:
        return batch
    cp_rank = ps.get_context_parallel_rank() if cp_rank is None else cp_rank
    out = {}
    for key, val in batch.items():
        if val is None:
            out[key] = None
            continue
        seq_dim = 1 if key != "attention_mask" else 2
        n = val.shape[seq_dim]
        assert n % (2 * cp_size) == 0
        chunks = val.view(*val.shape[:seq_dim], 2 * cp_size, n // (2 * cp_size), *val.shape[seq_dim + 1 :])
        idx = torch.tensor([cp_rank, 2 * cp_size - 1 - cp_rank], device=val.device)
        sel = chunks.index_select(seq_dim, idx)
        out[key] = sel.reshape(*val.shape[:seq_dim], -1, *val.shape[seq_dim + 1 :])
    return out


# ---- NVTX ---------------------------------------------------------------------

_nvtx_enabled = False


def configure_nvtx_profiling(enabled: bool) -> None:
    global _nvtx_enabled
    _nvtx_enabled = enabled


def nvtx_range_push(msg=None, suffix=None):
    if _nvtx_enabled and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(msg if suffix is None else f"{msg or ''}.{suffix}")


def nvtx_range_pop(msg=None, suffix=None):
    if _nvtx_enabled and torch.cuda.is_available():
        torch.cuda.nvtx.range_pop()


def nvtx_decorator(message: Optional[str] = None, color: Optional[str] = None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            nvtx_range_push(message or fn.__qualname__)
            try:
                return fn(*a, **k)
            finally:
                nvtx_range_pop()

        return wrapped

    return deco


# ---- experimental / deprecation gates ------------------------------------------


def experimental_fn(introduced_with_version: str):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            from . import config as _cfg

            if not _cfg.is_experimental_enabled():
                raise RuntimeError(f"{fn.__name__} is experimental; call config.set_experimental_flag(True)")
            return fn(*a, **k)

        return wrapped

    return deco


def deprecated(version=None, removal_version=None, alternative=None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            warnings.warn(f"{fn.__name__} is deprecated" + (f"; use {alternative}" if alternative else ""), DeprecationWarning)
            return fn(*a, **k)

        return wrapped

    return deco


def internal_api(fn):
    return fn


def prepare_input_tensors_for_wgrad_compute(grad_output, all_gathered_input):
    """Flatten [s, b, *] → [s*b, *] so wgrad is a single 2-D GEMM."""
    if grad_output.dim() == 3:
        grad_output = grad_output.reshape(-1, grad_output.shape[-1])
        all_gathered_input = all_gathered_input.reshape(-1, all_gathered_input.shape[-1])
    return grad_output, all_gathered_input


def get_te_version():
    return None


def is_float8tensor(t) -> bool:
    return False


def make_tp_sharded_tensor_for_checkpoint(tensor, key, tp_axis=0, replica_id=None, prepend_offsets=(), tp_group=None, dp_cp_group=None, **kwargs):
    from .dist_checkpointing.mapping import ShardedTensor
    from . import parallel_state as ps

    prepend_axis_num = len(prepend_offsets)
    tp_rank = get_pg_rank(tp_group) if tp_group is not None else ps.get_tensor_model_parallel_rank()
    tp_size = get_pg_size(tp_group) if tp_group is not None else ps.get_tensor_model_parallel_world_size()
    if replica_id is None:
        dp_rank = get_pg_rank(dp_cp_group) if dp_cp_group is not None else ps.get_data_parallel_rank(with_context_parallel=True)
        replica_id = (0, 0, dp_rank)
    return ShardedTensor.from_rank_offsets(
        key, tensor, *prepend_offsets, (tp_axis + prepend_axis_num, tp_rank, tp_size),
        replica_id=replica_id, prepend_axis_num=prepend_axis_num, **kwargs,
    )


def make_sharded_tensor_for_checkpoint(tensor, key, prepend_offsets=(), replica_id=None, tp_group=None, dp_cp_group=None, **kwargs):
    from .dist_checkpointing.mapping import ShardedTensor
    from . import parallel_state as ps

    prepend_axis_num = len(prepend_offsets)
    if replica_id is None:
        tp_rank = get_pg_rank(tp_group) if tp_group is not None else ps.get_tensor_model_parallel_rank()
        dp_rank = get_pg_rank(dp_cp_group) if dp_cp_group is not None else ps.get_data_parallel_rank(with_context_parallel=True)
        replica_id = (0, tp_rank, dp_rank)
    return ShardedTensor.from_rank_offsets(
        key, tensor, *prepend_offsets, replica_id=replica_id, prepend_axis_num=prepend_axis_num, **kwargs
    )


class StragglerDetector:
    """CUDA-event section timing + min/max over ranks (reference ``utils.py:1493-2147``).

    Usage::
        sd = StragglerDetector(); sd.configure(world, rank, enabled=True)
        with sd(bdata=True): batch = next(it)
        with sd(): loss = fwd_bwd()
        sd.report(total_flops, log_interval)
    """

    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance._init()
        return cls._instance

    def _init(self):
        self.enabled = False
        self.world = 1
        self.rank = 0
        self._ev: List[Tuple[Any, Any]] = []
        self._cpu: List[float] = []
        self._bdata: List[float] = []
        self._mode_bdata = False

    def configure(self, world, rank, mmcnt=1, amp=3.0, port=65535, prefill=1024, enabled=False):
        self.world, self.rank, self.enabled = world, rank, enabled
        self.amp = amp

    def __call__(self, bdata: bool = False):
        self._mode_bdata = bdata
        return self

    def __enter__(self):
        if not self.enabled:
            return self
        self._t0 = time.perf_counter()
        if torch.cuda.is_available() and not self._mode_bdata:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self._ev.append((s, e))
        return self

    def __exit__(self, *exc):
        if not self.enabled:
            return False
        dt = time.perf_counter() - self._t0
        if self._mode_bdata:
            self._bdata.append(dt)
        else:
            self._cpu.append(dt)
            if torch.cuda.is_available() and self._ev:
                self._ev[-1][1].record()
        return False

    def elapsed(self):
        gpu = 0.0
        if torch.cuda.is_available() and self._ev:
            torch.cuda.synchronize()
            gpu = sum(s.elapsed_time(e) for s, e in self._ev) / 1e3
        return sum(self._cpu), gpu, sum(self._bdata)

    def report(self, total_flops: float = 0.0, log_interval: int = 0) -> bool:
        if not self.enabled:
            return False
        cpu, gpu, bd = self.elapsed()
        t = torch.tensor([cpu, gpu, bd], dtype=torch.float64)
        rec = {"rank": self.rank, "cpu_s": cpu, "gpu_s": gpu, "batch_s": bd}
        if dist.is_initialized() and self.world > 1:
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            lo, hi = t.clone().to(dev), t.clone().to(dev)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            rec.update(min=lo.tolist(), max=hi.tolist())
        if total_flops and (gpu or cpu):
            rec["tflops"] = total_flops / max(gpu or cpu, 1e-9) / 1e12
        if self.rank == 0:
            logger.info("straggler report: %s", rec)
        self._ev.clear(), self._cpu.clear(), self._bdata.clear()
        self.last_report = rec
        return True
