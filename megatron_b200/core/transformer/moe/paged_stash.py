"""Paged stash for dynamically sized expert activations (reference ``transformer/moe/paged_stash.py:29-1326``).

With dropless routing the number of tokens an EP rank receives changes every micro-batch, but a
CUDA-graphed step needs static buffers, i.e. worst-case ones.  Keeping a worst-case buffer
alive per micro-batch in flight (1F1B holds ``pp`` of them) wastes most of the memory: the
*sum* over micro-batches is close to the average load, only each single one can be large.

So every activation saved for backward inside the stash context is moved out of its
worst-case buffer into a pool of fixed-size pages, sized for the expected total, and moved
back right before its backward.  Allocation is entirely on the device — a ring of free page
ids, a head / count pair and an overflow flag — so there is no host synchronisation and the
whole thing can be captured; the host polls ``check_paged_stash_overflow()`` once per step and
``PagedStashRunner`` re-runs a step that overflowed with stashing off.  An optional pinned-host
pool takes the overflow instead when ``num_tokens_host > 0``.

Shapes are static throughout: a ``[T_max, H]`` tensor with ``n ≤ T_max`` valid rows always
issues ``T_max`` row moves; rows ``≥ n`` and rows of an overflowing request land in a scratch
page.  (The row move itself is ``index_copy_`` / ``index_select``; on a B200 those are the
row-gather kernels of ``ops/csrc/moe_kernels.cu`` when the tensors are bf16.)
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import Any, Dict, Optional, Tuple

import torch

from .... import ops


def _stash_kernel_ok(t: torch.Tensor, buf) -> bool:
    return (t.is_cuda and ops.has_ext() and hasattr(ops.ext(), "paged_stash") and t.dtype == buf.dtype and (buf.hidden_size * t.element_size()) % 16 == 0
            and t.shape[-1] == buf.hidden_size)


class PagedStashBuffer:
    """``pages [P + 1, page_size, H]`` (the last page is scratch) with a device-resident free ring."""

    def __init__(self, num_tokens: int, hidden_size: int, page_size: int, device, dtype, overflow: Optional[torch.Tensor] = None,
                 host_spill: Optional[torch.Tensor] = None, num_tokens_host: int = 0):
        self.page_size, self.hidden_size, self.dtype, self.device = page_size, hidden_size, dtype, torch.device(device)
        self.num_pages = -(-num_tokens // page_size)
        self.num_host_pages = -(-num_tokens_host // page_size) if num_tokens_host > 0 else 0
        self.pages = torch.zeros(self.num_pages + 1, page_size, hidden_size, device=device, dtype=dtype)
        self.host_pages = None
        if self.num_host_pages:
            self.host_pages = torch.zeros(self.num_host_pages + 1, page_size, hidden_size, dtype=dtype, pin_memory=torch.cuda.is_available())
        self.overflow = overflow if overflow is not None else torch.zeros(1, dtype=torch.int32, device=device)
        self.host_spill = host_spill if host_spill is not None else torch.zeros(1, dtype=torch.int32, device=device)
        # state[:, 0] = device pool, state[:, 1] = host pool
        self.free_ring = [torch.arange(self.num_pages, device=device, dtype=torch.int64),
                          torch.arange(max(self.num_host_pages, 1), device=device, dtype=torch.int64)]
        self.head = torch.zeros(2, dtype=torch.int64, device=device)     # next id to hand out
        self.tail = torch.zeros(2, dtype=torch.int64, device=device)     # next slot to return an id to
        self.count = torch.tensor([self.num_pages, self.num_host_pages], dtype=torch.int64, device=device)

    def reset(self):
        self.free_ring[0].copy_(torch.arange(self.num_pages, device=self.device))
        self.head.zero_()
        self.tail.zero_()
        self.count.copy_(torch.tensor([self.num_pages, self.num_host_pages], device=self.device))
        self.overflow.zero_()
        self.host_spill.zero_()

    def free_pages(self) -> int:
        return int(self.count[0])

    # ---- device-side allocator (static shapes, no .item()) -------------------------------
    def _alloc(self, need: torch.Tensor, max_pages: int, pool: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """→ (page ids ``[max_pages]`` with the scratch id in unused slots, ok flag)."""
        cap = self.num_pages if pool == 0 else self.num_host_pages
        ok = (need <= self.count[pool]) & (cap > 0)
        take = torch.where(ok, need, torch.zeros_like(need))
        slot = torch.arange(max_pages, device=self.device)
        ring = self.free_ring[pool]
        ids = ring[(self.head[pool] + slot) % max(cap, 1)]
        ids = torch.where(slot < take, ids, torch.full_like(ids, cap))   # cap == scratch page index
        self.head[pool] = (self.head[pool] + take) % max(cap, 1)
        self.count[pool] -= take
        return ids, ok

    def _release(self, ids: torch.Tensor, pool: int) -> None:
        cap = self.num_pages if pool == 0 else self.num_host_pages
        if cap == 0:
            return
        valid = ids < cap
        n = valid.sum()
        # compact the valid ids to the front (stable) and write them behind the ring's tail
        order = torch.argsort((~valid).to(torch.int8), stable=True)
        packed = ids[order]
        slot = torch.arange(ids.numel(), device=self.device)
        pos = (self.tail[pool] + slot) % cap
        ring = self.free_ring[pool]
        keep = ring[pos]
        ring[pos] = torch.where(slot < n, packed, keep)
        self.tail[pool] = (self.tail[pool] + n) % cap
        self.count[pool] += n

    def __repr__(self):
        return (f"PagedStashBuffer(pages={self.num_pages}x{self.page_size}x{self.hidden_size} {self.dtype}, host_pages={self.num_host_pages}, "
                f"free={self.free_pages()})")


class PagedTensor:
    """Handle of one stashed ``[T_max, H]`` activation: its page table and valid-row count."""

    def __init__(self, tensor: torch.Tensor, num_tokens: Optional[torch.Tensor] = None):
        self.shape, self.dtype, self.device = tensor.shape, tensor.dtype, tensor.device
        self._tensor: Optional[torch.Tensor] = tensor
        t_max = tensor.shape[0]
        self.num_tokens = num_tokens if num_tokens is not None else torch.tensor(t_max, device=tensor.device)
        self.page_ids: Optional[torch.Tensor] = None
        self.pool = 0                    # tensor: 0 device / 1 host / 2 nowhere (overflow), decided on the device
        self.where: Optional[torch.Tensor] = None

    def _row_index(self, buf: PagedStashBuffer, ids: torch.Tensor, cap: int) -> torch.Tensor:
        rows = torch.arange(self.shape[0], device=self.device)
        page = ids[rows // buf.page_size]
        idx = page * buf.page_size + rows % buf.page_size
        scratch = cap * buf.page_size + rows % buf.page_size
        return torch.where(rows < self.num_tokens, idx, scratch)

    def offload_to_stash(self, buf: PagedStashBuffer) -> None:
        t = self._tensor.reshape(self.shape[0], -1)
        max_pages = -(-self.shape[0] // buf.page_size)
        need = (self.num_tokens.to(torch.int64) + buf.page_size - 1) // buf.page_size
        ids, ok = buf._alloc(need, max_pages, 0)
        if _stash_kernel_ok(t, buf):
            # one kernel: rows below the device-resident token count go to their page, nothing else is touched (reference paged_stash_copy_kernel)
            ops.ext().paged_stash(t.contiguous(), buf.pages, ids, self.num_tokens.to(torch.int64).reshape(1), buf.page_size, False)
            ops._count()
        else:
            buf.pages.view(-1, buf.hidden_size).index_copy_(0, self._row_index(buf, ids, buf.num_pages), t)
        self.where = torch.where(ok, 0, 2)
        self.page_ids = ids
        if buf.host_pages is not None:
            need_h = torch.where(ok, torch.zeros_like(need), need)
            ids_h, ok_h = buf._alloc(need_h, max_pages, 1)
            idx_h = self._row_index(buf, ids_h, buf.num_host_pages)
            buf.host_pages.view(-1, buf.hidden_size).index_copy_(0, idx_h.cpu(), t.to("cpu", non_blocking=True))
            self.host_ids = ids_h
            spilled = (~ok) & ok_h
            buf.host_spill.copy_(torch.maximum(buf.host_spill, spilled.to(torch.int32).reshape(1)))
            self.where = torch.where(spilled, 1, self.where)
        buf.overflow.copy_(torch.maximum(buf.overflow, (self.where == 2).to(torch.int32).reshape(1)))
        self._tensor = None              # the worst-case buffer can be reused now

    def reload_from_stash(self, buf: PagedStashBuffer) -> torch.Tensor:
        if buf.host_pages is None and buf.pages.is_cuda and _stash_kernel_ok(buf.pages.view(-1, buf.hidden_size), buf):
            # gather + zero-fill of the rows beyond the valid count in one kernel (reference paged_stash_pop_kernel)
            out = torch.empty(self.shape[0], buf.hidden_size, device=self.device, dtype=buf.dtype)
            ops.ext().paged_stash(out, buf.pages, self.page_ids, self.num_tokens.to(torch.int64).reshape(1), buf.page_size, True)
            ops._count()
            buf._release(self.page_ids, 0)
            self.page_ids = None
            return out.view(self.shape)
        out = buf.pages.view(-1, buf.hidden_size).index_select(0, self._row_index(buf, self.page_ids, buf.num_pages))
        if buf.host_pages is not None:
            idx_h = self._row_index(buf, self.host_ids, buf.num_host_pages)
            from_host = buf.host_pages.view(-1, buf.hidden_size).index_select(0, idx_h.cpu()).to(self.device, non_blocking=True)
            out = torch.where(self.where == 1, from_host, out)
            buf._release(self.host_ids, 1)
        buf._release(self.page_ids, 0)
        rows = torch.arange(self.shape[0], device=self.device).unsqueeze(-1)
        out = torch.where(rows < self.num_tokens, out, torch.zeros_like(out))
        self.page_ids = None
        return out.view(self.shape)


class PagedStashManager:
    """Owns one pool per ``(hidden, dtype)`` and the ``saved_tensors_hooks`` that route through it."""

    _instance: Optional["PagedStashManager"] = None

    @classmethod
    def get_instance(cls) -> "PagedStashManager":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self):
        self.enabled = False
        self.buffers: Dict[Tuple[int, torch.dtype], PagedStashBuffer] = {}
        self.overflow: Optional[torch.Tensor] = None
        self.host_spill: Optional[torch.Tensor] = None
        self.page_size = 64
        self.min_rows = 1
        self.current_num_tokens: Optional[torch.Tensor] = None
        self.stashed = 0

    def allocate_stash_buffers(self, sizes: Dict[Tuple[int, torch.dtype], int], page_size: int, device, num_tokens_host: int = 0) -> None:
        """``sizes``: tokens to provision per ``(hidden, dtype)`` — the caller's expected total load."""
        self.page_size = page_size
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
        self.host_spill = torch.zeros(1, dtype=torch.int32, device=device)
        self.buffers = {k: PagedStashBuffer(n, k[0], page_size, device, k[1], self.overflow, self.host_spill, num_tokens_host) for k, n in sizes.items()}

    def release_stash_buffers(self) -> None:
        self.buffers.clear()

    def set_num_tokens(self, n: Optional[torch.Tensor]) -> None:
        """Valid rows of the activations about to be saved (device scalar from the dispatcher)."""
        self.current_num_tokens = n

    def on_save_for_backward(self, tensor: torch.Tensor) -> Any:
        if not self.enabled or tensor.dim() != 2 or tensor.shape[0] < self.min_rows or isinstance(tensor, torch.nn.Parameter) or not tensor.is_floating_point():
            return tensor
        buf = self.buffers.get((tensor.shape[1], tensor.dtype))
        if buf is None or buf.device != tensor.device:
            return tensor
        pt = PagedTensor(tensor.detach(), self.current_num_tokens)
        pt.offload_to_stash(buf)
        self.stashed += 1
        return (pt, buf)

    def on_get_saved_tensor(self, saved: Any) -> torch.Tensor:
        if torch.is_tensor(saved):
            return saved
        pt, buf = saved
        return pt.reload_from_stash(buf)


class PagedStashContext:
    def __init__(self, manager: PagedStashManager):
        self.m = manager
        self.hooks = torch.autograd.graph.saved_tensors_hooks(manager.on_save_for_backward, manager.on_get_saved_tensor)

    def __enter__(self):
        self.m.enabled = True
        self.hooks.__enter__()
        return self

    def __exit__(self, *a):
        self.hooks.__exit__(*a)
        self.m.enabled = False


def get_paged_stash_context(enabled: bool = True, num_tokens: Optional[torch.Tensor] = None):
    """Wrap the expert MLP forward: ``with get_paged_stash_context(cfg.moe_paged_stash, tokens_on_rank): ...``"""
    m = PagedStashManager.get_instance()
    if not enabled or not m.buffers:
        return nullcontext()
    m.set_num_tokens(num_tokens)
    return PagedStashContext(m)


def paged_stash_reset(enabled: bool = True) -> None:
    m = PagedStashManager.get_instance()
    for b in m.buffers.values():
        b.reset()
    m.stashed = 0
    if not enabled:
        m.release_stash_buffers()


def check_paged_stash_overflow() -> bool:
    m = PagedStashManager.get_instance()
    return bool(m.overflow is not None and int(m.overflow.item()) != 0)


def check_paged_stash_host_spill() -> bool:
    m = PagedStashManager.get_instance()
    return bool(m.host_spill is not None and int(m.host_spill.item()) != 0)


class PagedStashRunner:
    """Wraps ``forward_backward_func``: run with stashing; if the pool overflowed, the saved
    activations of that step are garbage, so discard the gradients and run the step again with
    stashing off (worst-case buffers) — one retry always suffices."""

    def __init__(self, forward_backward_func, zero_grad_func=None):
        self.fb = forward_backward_func
        self.zero_grad = zero_grad_func
        self.reruns = 0

    def __call__(self, *args, data_iterator=None, **kwargs):
        from ...rerun_state_machine import RerunDataIterator

        it = data_iterator
        if it is not None and not isinstance(it, (RerunDataIterator, list)):
            it = RerunDataIterator(it)
        paged_stash_reset(True)
        out = self.fb(*args, data_iterator=it, **kwargs)
        if not check_paged_stash_overflow():
            if isinstance(it, RerunDataIterator):
                it.advance()
            return out
        self.reruns += 1
        if self.zero_grad is not None:
            self.zero_grad()
        if isinstance(it, RerunDataIterator):
            it.rewind()
        m = PagedStashManager.get_instance()
        saved, m.buffers = m.buffers, {}
        try:
            out = self.fb(*args, data_iterator=it, **kwargs)
        finally:
            m.buffers = saved
            paged_stash_reset(True)
        if isinstance(it, RerunDataIterator):
            it.advance()
        return out
