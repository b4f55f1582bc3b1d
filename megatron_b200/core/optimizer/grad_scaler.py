"""Loss scalers for fp16 training (reference ``optimizer/grad_scaler.py``)."""
from abc import ABC, abstractmethod

import torch


def _dev():
    return torch.cuda.current_device() if torch.cuda.is_available() else "cpu"


class MegatronGradScaler(ABC):
    def __init__(self, initial_scale: float):
        assert initial_scale > 0.0
        self._scale = torch.tensor([initial_scale], dtype=torch.float, device=_dev())

    @property
    def scale(self):
        return self._scale

    @property
    def inv_scale(self):
        return self._scale.double().reciprocal().float()

    @abstractmethod
    def update(self, found_inf: bool):
        ...

    @abstractmethod
    def state_dict(self):
        ...

    @abstractmethod
    def load_state_dict(self, state_dict):
        ...


class ConstantGradScaler(MegatronGradScaler):
    def update(self, found_inf: bool):
        pass

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict):
        pass


class DynamicGradScaler(MegatronGradScaler):
    """Back off by ``backoff_factor`` after ``hysteresis`` overflows; grow by
    ``growth_factor`` after ``growth_interval`` clean steps."""

    def __init__(self, initial_scale, min_scale, growth_factor, backoff_factor, growth_interval, hysteresis):
        super().__init__(initial_scale)
        assert 0.0 < min_scale <= initial_scale and growth_factor > 1.0 and 0.0 < backoff_factor < 1.0
        assert growth_interval > 0 and hysteresis > 0
        self.min_scale = torch.tensor([min_scale], dtype=torch.float, device=_dev())
        self.growth_factor = torch.tensor([growth_factor], dtype=torch.float, device=_dev())
        self.backoff_factor = torch.tensor([backoff_factor], dtype=torch.float, device=_dev())
        self.growth_interval, self.hysteresis = growth_interval, hysteresis
        self._growth_tracker, self._hysteresis_tracker = 0, hysteresis

    def update(self, found_inf: bool):
        if found_inf:
            self._growth_tracker = 0
            self._hysteresis_tracker -= 1
            if self._hysteresis_tracker <= 0:
                self._scale = torch.max(self._scale * self.backoff_factor, self.min_scale)
        else:
            self._growth_tracker += 1
            if self._growth_tracker == self.growth_interval:
                self._growth_tracker = 0
                self._hysteresis_tracker = self.hysteresis
                self._scale = self._scale * self.growth_factor

    def state_dict(self):
        return {"scale": self._scale, "growth_tracker": self._growth_tracker, "hysteresis_tracker": self._hysteresis_tracker}

    def load_state_dict(self, state_dict):
        self._scale = state_dict["scale"].to(_dev())
        self._growth_tracker = state_dict["growth_tracker"]
        self._hysteresis_tracker = state_dict["hysteresis_tracker"]
