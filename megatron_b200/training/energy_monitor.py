"""GPU energy accounting through NVML (reference ``training/energy_monitor.py:22-95``): joules per interval and per token."""
from __future__ import annotations

import time
from typing import Optional

import torch


class EnergyMonitor:
    """``monitor.lap()`` returns the energy (J) used by this rank's GPU since the previous lap; ``total()`` since ``setup()``.
    Uses ``nvmlDeviceGetTotalEnergyConsumption`` (mJ counter); falls back to integrating sampled power when unsupported."""

    def __init__(self):
        self._nvml = None
        self._h = None
        self._start_mj = self._last_mj = 0
        self._t_last = None
        self._p_last = None
        self._acc_j = 0.0
        self.enabled = False

    def setup(self, device_index: Optional[int] = None) -> bool:
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = device_index if device_index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
            self._nvml, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._start_mj = self._last_mj = self._read_mj()
            self.enabled = True
        except Exception:
            self.enabled = False
        return self.enabled

    def _read_mj(self) -> float:
        try:
            return float(self._nvml.nvmlDeviceGetTotalEnergyConsumption(self._h))
        except Exception:
            # integrate instantaneous power (mW) between calls
            now, p = time.time(), float(self._nvml.nvmlDeviceGetPowerUsage(self._h))
            if self._t_last is not None:
                self._acc_j += 0.5 * (p + self._p_last) / 1000.0 * (now - self._t_last)
            self._t_last, self._p_last = now, p
            return self._acc_j * 1000.0

    def lap(self) -> float:
        if not self.enabled:
            return 0.0
        mj = self._read_mj()
        out = (mj - self._last_mj) / 1000.0
        self._last_mj = mj
        return out

    def total(self) -> float:
        return (self._read_mj() - self._start_mj) / 1000.0 if self.enabled else 0.0

    def shutdown(self):
        if self.enabled:
            try:
                self._nvml.nvmlShutdown()
            except Exception:
                pass
            self.enabled = False
