"""Spec provider behind ``--transformer-impl transformer_engine`` (reference ``extensions/transformer_engine_spec_provider.py``):
returns the native modules, see ``extensions/transformer_engine.py``."""
from ..models.backends import LocalSpecProvider


class TESpecProvider(LocalSpecProvider):
    """Same building blocks as the local provider — the B200 kernels are not optional add-ons of a second backend."""
