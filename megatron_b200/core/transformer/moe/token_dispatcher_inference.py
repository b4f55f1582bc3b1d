"""Token dispatchers for MoE *inference* (reference ``transformer/moe/token_dispatcher_inference.py:51-675``).

Serving wants what training does not care about: every shape static (decode steps are replayed
from a CUDA graph), and no device→host read in the layer.  Both dispatchers below therefore
avoid the permutation whose output length depends on the routing:

  every local expert sees the full gathered token block ``[T, H]`` and its output is weighted
  by the (mostly zero) routing probability column of that expert,

i.e. the expert input is the block repeated ``num_local_experts`` times, ``tokens_per_expert`` is the
host constant ``[T] * L`` and nothing is read back.  For decode (``T`` = a few hundred rows at most,
weights dominate the traffic) the extra rows are free — the grouped GEMM is bound by streaming the
expert weights, which happens once either way.  Rows beyond the valid-token count (padding of a
graph-sized batch, the padded part of a shorter rank) have their routing probabilities zeroed so
they contribute nothing and are dropped after the combine.

* ``NCCLAllGatherDispatcher`` (``inference_moe_token_dispatcher_type="nccl"``): equal token count on
  every EP rank, plain all-gather / reduce-scatter.
* ``NVLSAllGatherVDispatcher`` (``"nvls"``): ranks may hold different token counts; blocks are padded
  to ``max_tokens_per_rank`` (a static shape) and exchanged through the variable-count collectives
  of ``parallel/collectives.py`` — multimem NVLink kernels when the EP group has a symmetric heap.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from ...tensor_parallel.mappings import gather_from_sequence_parallel_region, reduce_scatter_to_sequence_parallel_region
from .token_dispatcher import MoEAllGatherTokenDispatcher


class InferenceAllGatherDispatcherBase(MoEAllGatherTokenDispatcher):
    # [1] int32 on the device: valid tokens over all EP ranks this step (class-level so the experts can gate on it)
    _valid_tokens_tensor: Optional[torch.Tensor] = None
    _host_valid_tokens_estimate: Optional[int] = None

    def __init__(self, *args, runs_metadata_sync: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self._runs_metadata_sync = runs_metadata_sync

    @classmethod
    def _valid_tokens(cls) -> Optional[torch.Tensor]:
        return cls._valid_tokens_tensor

    @classmethod
    def _get_host_valid_tokens_estimate(cls) -> Optional[int]:
        return cls._host_valid_tokens_estimate

    @classmethod
    def allocate_valid_tokens_tensor(cls, device="cpu") -> None:
        if cls._valid_tokens_tensor is None or cls._valid_tokens_tensor.device != torch.device(device):
            InferenceAllGatherDispatcherBase._valid_tokens_tensor = torch.zeros(1, dtype=torch.int32, device=device)

    def update_metadata(self, local_tokens: int) -> None:
        raise NotImplementedError

    # ---- shared dense-masked expert layout --------------------------------------------------
    def _to_expert_layout(self, hidden_states, probs, row_valid: Optional[torch.Tensor]):
        T = hidden_states.shape[0]
        lo, hi = self.local_expert_indices[0], self.local_expert_indices[-1]
        local_probs = probs[:, lo : hi + 1]
        if row_valid is not None:
            local_probs = local_probs * row_valid.unsqueeze(-1).to(local_probs.dtype)
        L = self.num_local_experts
        self._T = T
        x = hidden_states.unsqueeze(0).expand(L, T, hidden_states.shape[-1]).reshape(L * T, -1)
        p = local_probs.t().reshape(L * T)
        tokens_per_expert = torch.full((L,), T, dtype=torch.long)          # host constant: no sync
        return x, tokens_per_expert, p

    def combine_preprocess(self, expert_output):
        return expert_output.view(self.num_local_experts, self._T, -1).sum(dim=0)

    def combine_postprocess(self, hidden_states):
        return hidden_states.view(self.hidden_shape)


class NCCLAllGatherDispatcher(InferenceAllGatherDispatcherBase):
    def update_metadata(self, local_tokens: int) -> None:
        cls = InferenceAllGatherDispatcherBase
        cls._host_valid_tokens_estimate = local_tokens * self.ep_size
        if cls._valid_tokens_tensor is not None:
            cls._valid_tokens_tensor.fill_(local_tokens * self.ep_size)

    def token_dispatch(self, hidden_states, probs):
        if self._runs_metadata_sync:
            self.update_metadata(hidden_states.shape[0])
        if self.ep_size > 1 or self.tp_size > 1:
            probs = gather_from_sequence_parallel_region(probs, group=self.tp_ep_group)
            hidden_states = gather_from_sequence_parallel_region(hidden_states, group=self.tp_ep_group, use_global_buffer=False)
        return hidden_states, probs

    def dispatch_postprocess(self, hidden_states, probs):
        return self._to_expert_layout(hidden_states, probs, None)

    def token_combine(self, hidden_states):
        if self.ep_size > 1 or self.tp_size > 1:
            hidden_states = reduce_scatter_to_sequence_parallel_region(hidden_states, group=self.tp_ep_group)
        return hidden_states


class NVLSAllGatherVDispatcher(InferenceAllGatherDispatcherBase):
    """Variable token count per EP rank.  ``set_real_token_count_tensor`` hands in the device
    scalar with THIS rank's valid rows (the engine's batch bookkeeping already has it); blocks
    travel padded to ``max_tokens_per_rank``."""

    _real_token_count: Optional[torch.Tensor] = None
    _step_metadata: Optional[torch.Tensor] = None     # [1 + ep]: total valid, then per-rank counts
    _max_tokens_per_rank: Optional[int] = None

    @classmethod
    def set_real_token_count_tensor(cls, tensor: torch.Tensor) -> None:
        NVLSAllGatherVDispatcher._real_token_count = tensor

    @classmethod
    def modify_real_token_count_for_mtp(cls, mtp_token_count: int) -> None:
        if cls._real_token_count is not None:
            cls._real_token_count.fill_(mtp_token_count)

    @classmethod
    def allocate_buffers(cls, ep_size: int, max_tokens_per_rank: int, device="cpu") -> None:
        NVLSAllGatherVDispatcher._step_metadata = torch.zeros(1 + ep_size, dtype=torch.int32, device=device)
        NVLSAllGatherVDispatcher._max_tokens_per_rank = max_tokens_per_rank
        InferenceAllGatherDispatcherBase._valid_tokens_tensor = NVLSAllGatherVDispatcher._step_metadata[0:1]

    @classmethod
    def _delete_buffers(cls):
        NVLSAllGatherVDispatcher._step_metadata = None
        NVLSAllGatherVDispatcher._real_token_count = None
        InferenceAllGatherDispatcherBase._valid_tokens_tensor = None

    @classmethod
    def _ep_max_tokens(cls) -> Optional[int]:
        return cls._max_tokens_per_rank

    def update_metadata(self, local_tokens: int) -> None:
        """One tiny all-gather of the per-rank counts; stays on the device."""
        cls = NVLSAllGatherVDispatcher
        if cls._step_metadata is None or cls._step_metadata.numel() != 1 + self.ep_size:
            cls.allocate_buffers(self.ep_size, cls._max_tokens_per_rank or local_tokens, device=cls._real_token_count.device if cls._real_token_count is not None else "cpu")
        mine = cls._real_token_count if cls._real_token_count is not None else torch.tensor(local_tokens, dtype=torch.int32, device=cls._step_metadata.device)
        mine = mine.to(torch.int32).reshape(1)
        counts = cls._step_metadata[1:]
        if self.ep_size > 1:
            dist.all_gather_into_tensor(counts, mine, group=self.ep_group)
        else:
            counts.copy_(mine)
        cls._step_metadata[0:1] = counts.sum().reshape(1)
        InferenceAllGatherDispatcherBase._host_valid_tokens_estimate = local_tokens * self.ep_size

    def dispatch_preprocess(self, hidden_states, routing_map, probs):
        hs, probs = super().dispatch_preprocess(hidden_states, routing_map, probs)
        m = NVLSAllGatherVDispatcher._max_tokens_per_rank or hs.shape[0]
        if hs.shape[0] > m:
            raise ValueError(f"{hs.shape[0]} tokens exceed max_tokens_per_rank={m}")
        self._local_rows = hs.shape[0]
        if hs.shape[0] < m:                                # pad to the static exchange shape
            hs = torch.cat([hs, hs.new_zeros(m - hs.shape[0], hs.shape[1])])
            probs = torch.cat([probs, probs.new_zeros(m - probs.shape[0], probs.shape[1])])
        return hs, probs

    def token_dispatch(self, hidden_states, probs):
        if self._runs_metadata_sync:
            self.update_metadata(self._local_rows)
        m = hidden_states.shape[0]
        if self.ep_size > 1:
            from ....parallel.collectives import all_gather_v

            sizes = [m] * self.ep_size
            hidden_states = all_gather_v(hidden_states, sizes, group=self.ep_group)
            probs = all_gather_v(probs, sizes, group=self.ep_group)
        counts = NVLSAllGatherVDispatcher._step_metadata[1:].to(hidden_states.device)
        row = torch.arange(m, device=hidden_states.device)
        self._row_valid = (row.unsqueeze(0) < counts.unsqueeze(1)).reshape(-1)      # [ep * m]
        return hidden_states, probs

    def dispatch_postprocess(self, hidden_states, probs):
        return self._to_expert_layout(hidden_states, probs, self._row_valid)

    def token_combine(self, hidden_states):
        if self.ep_size > 1:
            from ....parallel.collectives import reduce_scatter_v

            m = hidden_states.shape[0] // self.ep_size
            hidden_states = reduce_scatter_v(hidden_states, [m] * self.ep_size, group=self.ep_group)
        return hidden_states

    def combine_postprocess(self, hidden_states):
        return hidden_states[: self._local_rows].view(self.hidden_shape)


def get_inference_token_dispatcher(kind: str):
    return {"nccl": NCCLAllGatherDispatcher, "nvls": NVLSAllGatherVDispatcher}[kind]
