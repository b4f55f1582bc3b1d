from ..training.energy_monitor import EnergyMonitor  # noqa: F401  (reference ``core/energy_monitor.py``)
