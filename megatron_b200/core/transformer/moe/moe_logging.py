"""Per-layer MoE metric accounting (reference ``transformer/moe/moe_logging.py:34-390``).

Routers push scalars (aux loss, z-loss, tokens-per-expert imbalance, ...) while the forward
runs; once per logging interval ``report`` reduces them over the ranks that each hold a part
of the picture and hands means / per-layer curves to TensorBoard, W&B and the console line.
A metric row has one slot per layer so pipeline stages can be summed into place, and the
tracker is a process-wide singleton because routers have no handle on the trainer.

``moe_utils.save_to_aux_losses_tracker`` (the reference's older entry point) writes into the
same tracker."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import torch
import torch.distributed as dist


@dataclass
class MetricEntry:
    values: torch.Tensor                     # [num_layers] running sum since the last clear
    reduce_group: Optional[object] = None    # partial sums live on these ranks (SUM)
    avg_group: Optional[object] = None       # replicas disagree by data (MEAN)
    needs_dp_avg: bool = True                # averaged over DP at report time
    percentiles: Optional[List[float]] = None


_TRACKER: Optional["MoEMetricsTracker"] = None


def get_moe_metrics_tracker() -> "MoEMetricsTracker":
    global _TRACKER
    if _TRACKER is None:
        _TRACKER = MoEMetricsTracker()
    return _TRACKER


def set_moe_metrics_tracker(tracker: "MoEMetricsTracker") -> None:
    global _TRACKER
    _TRACKER = tracker


def destroy_moe_metrics_tracker() -> None:
    global _TRACKER
    _TRACKER = None


class MoEMetricsTracker:
    def __init__(self):
        self._metrics: Dict[str, MetricEntry] = {}

    @property
    def metrics(self) -> Dict[str, MetricEntry]:
        return self._metrics

    def record(self, name: str, value: torch.Tensor, layer_number: Optional[int], num_layers: int, reduce_group=None, avg_group=None,
               needs_dp_avg: bool = True, percentiles: Optional[List[float]] = None) -> None:
        """Accumulate ``value`` (0-d) into layer ``layer_number`` (1-based) of metric ``name``."""
        if layer_number is None:
            return
        e = self._metrics.get(name)
        if e is None or e.values.numel() != num_layers:
            e = MetricEntry(torch.zeros(num_layers, device=value.device, dtype=torch.float32))
            self._metrics[name] = e
        e.values[layer_number - 1] += value.detach().float().reshape(())
        e.reduce_group, e.avg_group, e.needs_dp_avg, e.percentiles = reduce_group, avg_group, needs_dp_avg, percentiles

    def ensure_initialized(self, names: Union[str, List[str]], num_layers: int, device="cpu") -> None:
        """Ranks without a MoE layer (other PP stages) still take part in the reductions."""
        for n in self._resolve_names(names):
            if n not in self._metrics:
                self._metrics[n] = MetricEntry(torch.zeros(num_layers, device=device, dtype=torch.float32))

    def clear(self) -> None:
        for e in self._metrics.values():
            e.values.zero_()

    def _resolve_names(self, track_names) -> List[str]:
        if track_names is None:
            return list(self._metrics)
        return [track_names] if isinstance(track_names, str) else list(track_names)

    def _sync_metrics(self, names: List[str], pp_group=None, dp_group=None) -> None:
        for n in names:
            e = self._metrics.get(n)
            if e is None or not dist.is_initialized():
                continue
            v = e.values
            if pp_group is not None and dist.get_world_size(pp_group) > 1:
                dist.all_reduce(v, group=pp_group)             # each stage filled its own layers
            if e.reduce_group is not None and dist.get_world_size(e.reduce_group) > 1:
                dist.all_reduce(v, group=e.reduce_group)
            if e.avg_group is not None and dist.get_world_size(e.avg_group) > 1:
                dist.all_reduce(v, group=e.avg_group)
                v.div_(dist.get_world_size(e.avg_group))
            if e.needs_dp_avg and dp_group is not None and dist.get_world_size(dp_group) > 1:
                dist.all_reduce(v, group=dp_group)
                v.div_(dist.get_world_size(dp_group))

    @staticmethod
    def _count_moe_layers(num_layers: int, moe_layer_freq=None, mtp_num_layers: Optional[int] = None) -> int:
        if moe_layer_freq is None:
            n = num_layers
        elif isinstance(moe_layer_freq, int):
            n = sum(1 for i in range(num_layers) if i % moe_layer_freq == 0)
        else:
            n = int(sum(moe_layer_freq))
        return n + (mtp_num_layers or 0)

    def _aggregate(self, names: List[str], loss_scale: float, num_moe_layers: int) -> Dict[str, torch.Tensor]:
        out = {}
        for n in names:
            e = self._metrics.get(n)
            if e is not None:
                out[n] = e.values.float().sum() * loss_scale / max(num_moe_layers, 1)
        return out

    def report(self, loss_scale: float, iteration: int, writer=None, wandb_writer=None, total_loss_dict: Optional[dict] = None,
               per_layer_logging: bool = False, force_initialize: bool = False, track_names=None, num_layers: Optional[int] = None,
               moe_layer_freq=None, mtp_num_layers: Optional[int] = None, pp_group=None, dp_group=None) -> str:
        """Reduce → aggregate → log → clear.  Returns the console fragment (``name: value |``)."""
        if force_initialize and num_layers is not None and track_names is not None:
            self.ensure_initialized(track_names, num_layers + (mtp_num_layers or 0))
        names = [n for n in self._resolve_names(track_names) if n in self._metrics]
        self._sync_metrics(names, pp_group, dp_group)
        any_entry = next((self._metrics[n] for n in names), None)
        slots = any_entry.values.numel() if any_entry is not None else 0
        n_moe = self._count_moe_layers(num_layers if num_layers is not None else slots, moe_layer_freq, mtp_num_layers)
        scalars = self._aggregate(names, loss_scale, n_moe)
        if total_loss_dict is not None:
            for k, v in scalars.items():
                total_loss_dict[k] = total_loss_dict.get(k, 0.0) + v
        self._log_scalars(scalars, iteration, writer, wandb_writer)
        if per_layer_logging:
            self._log_per_layer(names, loss_scale, iteration, writer, wandb_writer)
        text = self._format(scalars)
        self.clear()
        return text

    @staticmethod
    def _log_scalars(scalars, iteration, writer, wandb_writer) -> None:
        for k, v in scalars.items():
            if writer is not None:
                writer.add_scalar(k, float(v), iteration)
            if wandb_writer is not None:
                wandb_writer.log({k: float(v)}, iteration)

    def _log_per_layer(self, names, loss_scale, iteration, writer, wandb_writer) -> None:
        for n in names:
            e = self._metrics[n]
            for i, v in enumerate((e.values.float() * loss_scale).tolist()):
                if v == 0.0:
                    continue                                   # dense layer: no router
                if writer is not None:
                    writer.add_scalar(f"moe/{n}_layer_{i}", v, iteration)
                if wandb_writer is not None:
                    wandb_writer.log({f"moe/{n}_layer_{i}": v}, iteration)
            if e.percentiles:
                nz = e.values[e.values != 0].float() * loss_scale
                if nz.numel():
                    qs = torch.quantile(nz, torch.tensor(e.percentiles, device=nz.device, dtype=nz.dtype))
                    for p, q in zip(e.percentiles, qs.tolist()):
                        if writer is not None:
                            writer.add_scalar(f"moe/{n}_p{int(p * 100)}", q, iteration)

    @staticmethod
    def _format(scalars: Dict[str, Union[float, torch.Tensor]]) -> str:
        return "".join(f" {k}: {float(v):.6E} |" for k, v in scalars.items())
