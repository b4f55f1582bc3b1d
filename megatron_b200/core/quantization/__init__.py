"""Per-layer quantisation recipes selected by glob patterns from a YAML file (reference ``core/quantization/`` 271 LoC).

```yaml
configs:
  fp8_hybrid:  {recipe: tensorwise, fp8_format: hybrid}
  bf16:        {recipe: none}
matchers:
  - {pattern: "decoder.layers.0.*", config: bf16}       # first layer stays bf16
  - {pattern: "*.linear_fc1",       config: fp8_hybrid}
  - {pattern: "*",                  config: bf16}
```
The first matching pattern wins.  ``apply_quantization_recipe(model, recipe)`` attaches the matching ``QuantizationConfig`` to every tensor-parallel linear
layer (``module.quant_config``); the layers consult it before the model-wide ``config.fp8`` / autocast state (``tensor_parallel/layers.py::_fp8_choice``), so a
recipe can quantise single projections, keep the first / last layers or the output head in bf16, or mix MXFP8 and NVFP4.  ``TransformerConfig.quant_recipe``
(a ``RecipeConfig`` or a YAML path; reference ``--te-precision-config-file`` / ``--kitchen-config-file``) is applied by ``GPTModel`` at construction."""
from __future__ import annotations

import fnmatch
from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass
class QuantizationConfig:
    name: str = "bf16"
    recipe: str = "none"          # none | tensorwise | delayed | mxfp8 (blockwise) | nvfp4
    fp8_format: str = "hybrid"    # hybrid (e4m3 fwd / e5m2 grads) | e4m3
    extra: Dict = field(default_factory=dict)

    @property
    def enabled(self) -> bool:
        return self.recipe != "none"


@dataclass
class Matcher:
    pattern: str
    config: str


class RecipeConfig:
    def __init__(self, configs: Dict[str, QuantizationConfig], matchers: List[Matcher]):
        self.configs, self.matchers = configs, matchers

    @classmethod
    def from_dict(cls, d: dict) -> "RecipeConfig":
        cfgs = {}
        for name, body in (d.get("configs") or {}).items():
            body = dict(body or {})
            cfgs[name] = QuantizationConfig(name=name, recipe=body.pop("recipe", "none"), fp8_format=body.pop("fp8_format", "hybrid"), extra=body)
        ms = [Matcher(m["pattern"], m["config"]) for m in d.get("matchers") or []]
        for m in ms:
            if m.config not in cfgs:
                raise KeyError(f"matcher '{m.pattern}' refers to unknown config '{m.config}'")
        return cls(cfgs, ms)

    @classmethod
    def from_yaml_file(cls, path: str) -> "RecipeConfig":
        import yaml

        with open(path) as f:
            return cls.from_dict(yaml.safe_load(f))

    def match(self, module_path: str) -> Optional[QuantizationConfig]:
        for m in self.matchers:
            if fnmatch.fnmatchcase(module_path, m.pattern):
                return self.configs[m.config]
        return None


def load_quantization_recipe(path: str) -> RecipeConfig:
    return RecipeConfig.from_yaml_file(path)


def get_quant_config_or_none(module_path: str, recipe: Optional[RecipeConfig]) -> Optional[QuantizationConfig]:
    return recipe.match(module_path) if recipe is not None else None


def apply_quantization_recipe(model, recipe) -> Dict[str, str]:
    """Attach per-layer configs; returns {module path: config name} for the layers that matched (handy for logging what runs in which precision)."""
    from ..tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear

    if isinstance(recipe, str):
        recipe = load_quantization_recipe(recipe)
    chosen: Dict[str, str] = {}
    for name, mod in model.named_modules():
        if isinstance(mod, (ColumnParallelLinear, RowParallelLinear)):
            qc = recipe.match(name)
            if qc is not None:
                mod.quant_config = qc
                chosen[name] = qc.name
    return chosen
