"""Engine interface (reference ``inference/engines/abstract_engine.py``)."""
import abc


class AbstractEngine(abc.ABC):
    @abc.abstractmethod
    def generate(self, *args, **kwargs):
        """Run the requests to completion and return them."""
