"""Record / replay of top-k routing decisions (reference ``transformer/moe/router_replay.py:8-208``).

RL post-training evaluates the same tokens under several policies (rollout engine, reference
policy, trainee) and wants every pass to send a token to the SAME experts; activation
recompute wants the recomputed forward to route like the original one.  Each router that
opts in (``config.moe_enable_routing_replay``) owns one ``RouterReplay``; the class keeps
the list of live instances so a trainer can switch all of them with one call.

* ``RECORD``           — run the normal top-k, remember the indices.
* ``REPLAY_FORWARD``   — ignore the scores' ordering: use the target indices, gather their
  scores as probabilities (gradients still flow to the gathered scores) and queue the
  indices for the matching recompute.
* ``REPLAY_BACKWARD``  — pop the queued indices in FIFO order (recompute inside backward).
"""
from __future__ import annotations

from collections import deque
from enum import Enum
from typing import Callable, Deque, List, Optional, Tuple

import torch


class RouterReplayAction(Enum):
    RECORD = "record"
    REPLAY_FORWARD = "replay_forward"
    REPLAY_BACKWARD = "replay_backward"


class RouterReplay:
    global_router_replay_instances: List["RouterReplay"] = []

    def __init__(self):
        self.target_topk_idx: Optional[torch.Tensor] = None
        self.recorded_topk_idx: Optional[torch.Tensor] = None
        self.router_replay_action: Optional[RouterReplayAction] = None
        self.replay_backward_list: Deque[torch.Tensor] = deque()
        self.static_buffer: Optional[torch.Tensor] = None  # CUDA-graph-stable record slot
        RouterReplay.global_router_replay_instances.append(self)

    # ---- whole-model switches -------------------------------------------------------------
    @staticmethod
    def set_replay_data(all_layers_topk_indices: List[torch.Tensor]):
        inst = RouterReplay.global_router_replay_instances
        if len(all_layers_topk_indices) != len(inst):
            raise ValueError(f"replay data for {len(all_layers_topk_indices)} routers, model has {len(inst)}")
        for r, idx in zip(inst, all_layers_topk_indices):
            r.set_target_indices(idx)

    @staticmethod
    def get_recorded_data() -> List[Optional[torch.Tensor]]:
        return [r.get_recorded_indices() for r in RouterReplay.global_router_replay_instances]

    @staticmethod
    def clear_global_indices():
        for r in RouterReplay.global_router_replay_instances:
            r.clear_indices()

    @staticmethod
    def set_global_router_replay_action(action: Optional[RouterReplayAction]):
        for r in RouterReplay.global_router_replay_instances:
            r.set_router_replay_action(action)

    @staticmethod
    def clear_global_router_replay_action():
        RouterReplay.set_global_router_replay_action(None)

    @staticmethod
    def clear_global_router_replay_instances():
        RouterReplay.global_router_replay_instances.clear()

    @staticmethod
    def set_global_static_buffers(static_buffer: torch.Tensor):
        """``static_buffer [num_routers, max_tokens, topk]``: router ``i`` records into row ``i``
        in place so a captured graph keeps writing to the same address."""
        inst = RouterReplay.global_router_replay_instances
        if static_buffer.shape[0] != len(inst):
            raise ValueError(f"static buffer has {static_buffer.shape[0]} rows for {len(inst)} routers")
        for i, r in enumerate(inst):
            r.set_static_buffer(static_buffer[i])

    @staticmethod
    def clear_global_static_buffers():
        for r in RouterReplay.global_router_replay_instances:
            r.clear_static_buffer()

    # ---- per-router state -----------------------------------------------------------------
    def set_target_indices(self, topk_indices: torch.Tensor):
        self.target_topk_idx = topk_indices
        self.replay_backward_list.clear()

    def get_recorded_indices(self) -> Optional[torch.Tensor]:
        return self.recorded_topk_idx

    def clear_indices(self):
        self.target_topk_idx = None
        self.recorded_topk_idx = None
        self.replay_backward_list.clear()

    def set_router_replay_action(self, action: Optional[RouterReplayAction]):
        self.router_replay_action = action

    def clear_router_replay_action(self):
        self.router_replay_action = None

    def set_static_buffer(self, buffer: torch.Tensor):
        self.static_buffer = buffer

    def clear_static_buffer(self):
        self.static_buffer = None

    def record_indices(self, topk_indices: torch.Tensor):
        if self.static_buffer is not None:
            n = topk_indices.shape[0]
            if n > self.static_buffer.shape[0]:
                raise ValueError(f"{n} tokens do not fit the static replay buffer ({self.static_buffer.shape[0]})")
            self.static_buffer[:n].copy_(topk_indices)
            self.recorded_topk_idx = self.static_buffer[:n]
        else:
            self.recorded_topk_idx = topk_indices

    def get_replay_topk(self, scores: torch.Tensor, topk: int, default_compute_topk: Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]]):
        """Drop-in for the router's top-k: → ``(values [T, k], indices [T, k])``."""
        act = self.router_replay_action
        if act is None:
            return default_compute_topk(scores, topk)
        if act is RouterReplayAction.RECORD:
            vals, idx = default_compute_topk(scores, topk)
            self.record_indices(idx)
            return vals, idx
        if act is RouterReplayAction.REPLAY_FORWARD:
            if self.target_topk_idx is None:
                raise RuntimeError("REPLAY_FORWARD without target indices (call set_replay_data first)")
            idx = self.target_topk_idx.to(scores.device)
            self.replay_backward_list.append(idx)
        else:
            if not self.replay_backward_list:
                raise RuntimeError("REPLAY_BACKWARD with an empty replay queue")
            idx = self.replay_backward_list.popleft().to(scores.device)
        if idx.shape != (scores.shape[0], topk):
            raise ValueError(f"replayed indices {tuple(idx.shape)} do not match [{scores.shape[0]}, {topk}]")
        return scores.gather(1, idx), idx
