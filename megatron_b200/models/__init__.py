"""Model zoo entry points (re-exports of ``megatron_b200.core.models`` + presets)."""
from .presets import PRESETS, build_gpt_model, make_transformer_config


def __getattr__(name):
    if name == "GPTModel":
        from ..core.models.gpt.gpt_model import GPTModel

        return GPTModel
    raise AttributeError(name)
