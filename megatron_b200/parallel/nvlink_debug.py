"""Debug harness for the hand-rolled cross-GPU protocols (what a framework with its own flag/barrier protocols inside kernels must add).

``CheckedNVLinkBackend`` wraps an ``NVLinkBackend`` (enable with ``MEGATRON_B200_NVL_DEBUG=1``; ``collectives.enable_for_group`` then returns the wrapper) and
after EVERY NVLink collective or fused GEMM⇄collective op recomputes the result through NCCL (+ a plain GEMM) and compares:

* all-gather results must be BITWISE identical (pure data movement — any difference is a protocol bug: a missed flag, a stale epoch, a torn 16-byte store);
* reductions / fused GEMM results must agree to a tolerance derived from the dtype (the switch accumulates in fp32 in a different order);
* a per-rank operation sequence number and a rolling hash of (op name, shapes) is all-reduced every ``sync_every`` ops: ranks that diverged in WHICH collective
  they issued (the classic cause of flag-protocol deadlocks) are reported by name instead of hanging;
* optional poison: inputs are copied and the original workspace view is filled with NaN after the comparison, so a later kernel that reads a buffer the protocol
  declared dead produces NaNs immediately instead of plausible stale numbers.

The wrapper is slow by design (2x communication + synchronisation); it is a correctness tool for bring-up, CI on small shapes and post-mortems."""
from __future__ import annotations

import hashlib
import os
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


class NVLinkProtocolError(RuntimeError):
    pass


class CheckedNVLinkBackend:
    def __init__(self, inner, group=None, sync_every: int = 16, rtol: Optional[float] = None, poison: bool = False):
        self.inner, self.group = inner, group if group is not None else getattr(inner, "group", None)
        self.sync_every, self.rtol, self.poison = sync_every, rtol, poison
        self.seq = 0
        self._hash = hashlib.sha256()
        self.checked = 0

    def __getattr__(self, name):           # everything not intercepted goes straight through
        return getattr(self.inner, name)

    # ---- bookkeeping ----
    def _record(self, op: str, *tensors) -> None:
        self.seq += 1
        self._hash.update(f"{op}:{[tuple(t.shape) for t in tensors if isinstance(t, torch.Tensor)]}".encode())
        if self.seq % self.sync_every == 0 and dist.is_initialized():
            digest = int.from_bytes(self._hash.digest()[:7], "little")
            t = torch.tensor([digest, -digest, self.seq, -self.seq], dtype=torch.int64, device=self._dev(tensors))
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            if t[0] != -t[1] or t[2] != -t[3]:
                raise NVLinkProtocolError(f"rank {dist.get_rank()}: collective sequence diverged at op #{self.seq} ('{op}'): ranks issued different operations")

    @staticmethod
    def _dev(tensors):
        for t in tensors:
            if isinstance(t, torch.Tensor):
                return t.device
        return "cpu"

    def _tol(self, ref: torch.Tensor) -> float:
        if self.rtol is not None:
            return self.rtol
        return {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 5e-3}.get(ref.dtype, 1e-3)

    def _compare(self, op: str, got: torch.Tensor, ref: torch.Tensor, exact: bool) -> None:
        self.checked += 1
        if got.shape != ref.shape:
            raise NVLinkProtocolError(f"{op}: shape {tuple(got.shape)} vs reference {tuple(ref.shape)}")
        if exact:
            if not torch.equal(got, ref):
                bad = (got != ref).reshape(-1).nonzero()
                raise NVLinkProtocolError(f"rank {dist.get_rank()}: {op} is not bitwise identical to NCCL: {bad.numel()} elements differ, first at flat index {int(bad[0])}")
            return
        g, r = got.float(), ref.float()
        if not torch.isfinite(g).all():
            raise NVLinkProtocolError(f"rank {dist.get_rank()}: {op} produced non-finite values (poisoned or unwritten buffer was read)")
        err = (g - r).abs().max().item()
        lim = self._tol(ref) * (r.abs().max().item() + 1e-6)
        if err > lim:
            raise NVLinkProtocolError(f"rank {dist.get_rank()}: {op} max error {err:.4g} exceeds {lim:.4g} against the NCCL reference")

    # ---- reference implementations over NCCL / gloo ----
    def _ref_all_gather(self, x: torch.Tensor) -> torch.Tensor:
        ws = dist.get_world_size(self.group)
        out = x.new_empty((x.shape[0] * ws,) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return out

    def _ref_reduce_scatter(self, x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        ws = dist.get_world_size(self.group)
        acc = x.float().contiguous()
        out = acc.new_empty((x.shape[0] // ws,) + tuple(x.shape[1:]))
        dist.reduce_scatter_tensor(out, acc, group=self.group)
        return (out * scale).to(x.dtype)

    # ---- intercepted operations ----
    def all_gather(self, x: torch.Tensor, *a, **k):
        self._record("all_gather", x)
        keep = x.clone()
        out = self.inner.all_gather(x, *a, **k)
        self._compare("all_gather", out, self._ref_all_gather(keep), exact=True)
        return out

    def reduce_scatter(self, x: torch.Tensor, *a, scale: float = 1.0, **k):
        self._record("reduce_scatter", x)
        keep = x.clone()
        out = self.inner.reduce_scatter(x, *a, scale=scale, **k)
        self._compare("reduce_scatter", out, self._ref_reduce_scatter(keep, scale), exact=False)
        if self.poison and x.is_floating_point():
            x.fill_(float("nan"))          # the protocol consumed this buffer: nobody may read it again
        return out

    def all_reduce(self, x: torch.Tensor, *a, scale: float = 1.0, **k):
        self._record("all_reduce", x)
        ref = x.float().clone()
        dist.all_reduce(ref, group=self.group)
        out = self.inner.all_reduce(x, *a, scale=scale, **k)
        self._compare("all_reduce", out, (ref * scale).to(x.dtype), exact=False)
        return out

    def all_gather_gemm(self, x: torch.Tensor, w: torch.Tensor):
        self._record("all_gather_gemm", x, w)
        ref = torch.matmul(self._ref_all_gather(x.clone()).float(), w.float().t())
        out = self.inner.all_gather_gemm(x, w)
        self._compare("all_gather_gemm", out, ref.to(out.dtype).view_as(out), exact=False)
        return out

    def gemm_reduce_scatter(self, x: torch.Tensor, w: torch.Tensor):
        self._record("gemm_reduce_scatter", x, w)
        part = torch.matmul(x.float(), w.float().t())
        ref = self._ref_reduce_scatter(part.view(x.shape[0], -1, w.shape[0]) if x.dim() == 3 else part)
        out = self.inner.gemm_reduce_scatter(x, w)
        self._compare("gemm_reduce_scatter", out, ref.to(out.dtype).reshape(out.shape), exact=False)
        return out


def maybe_wrap(backend, group=None):
    """Used by ``collectives.enable_for_group``: returns the checked wrapper when ``MEGATRON_B200_NVL_DEBUG`` is set."""
    if os.environ.get("MEGATRON_B200_NVL_DEBUG", "0") not in ("", "0"):
        return CheckedNVLinkBackend(backend, group, poison=os.environ.get("MEGATRON_B200_NVL_DEBUG") == "poison")
    return backend
