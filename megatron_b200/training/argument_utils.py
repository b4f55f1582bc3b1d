"""Dataclass <-> argparse bridge and the YAML configuration container (reference ``megatron/training/argument_utils.py`` ``ArgumentGroupFactory``,
``training/yaml_arguments.py``, ``training/config/``).

* ``add_dataclass_arguments(parser, ConfigClass, …)`` generates one ``--kebab-case`` flag per dataclass field from its type annotation and default
  (``bool`` → ``--flag / --no-flag``, ``Optional[T]`` → T, ``List[T]`` → ``nargs="*"``, ``Literal`` / ``Enum`` → choices; callables, tensors and dtype
  fields are skipped), so every knob of ``TransformerConfig`` / ``ModelParallelConfig`` / ``OptimizerConfig`` / ``DistributedDataParallelConfig`` is a
  command-line option without being written twice.  Flags the hand-written parser already defines are left alone.
* ``dataclass_from_args(ConfigClass, args, **overrides)`` builds the config back from a namespace (fields absent from the namespace keep their defaults).
* ``load_yaml_config(path)`` / ``apply_yaml(args, cfg)``: a YAML file with (optionally nested) sections whose leaves are argument names —
  ``--yaml-cfg run.yaml`` then sets exactly what the equivalent flags would; explicit command-line flags win over the file.
* ``args_to_yaml(args)``: the effective configuration, grouped by the config class that owns each field (what the reference logs at start-up)."""
from __future__ import annotations

import argparse
import dataclasses
import enum
import typing
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

_SKIP_TYPES = ("Callable", "torch.dtype", "Tensor", "ProcessGroup", "ModuleSpec", "Timers")


def _unwrap_optional(tp):
    origin = typing.get_origin(tp)
    if origin is typing.Union:
        args = [a for a in typing.get_args(tp) if a is not type(None)]
        if len(args) == 1:
            return args[0], True
    return tp, False


def _resolve_hints(cls) -> Dict[str, Any]:
    try:
        return typing.get_type_hints(cls)
    except Exception:            # forward references to types that are not importable here
        return {f.name: f.type for f in dataclasses.fields(cls)}


def field_to_argparse_kwargs(name: str, tp, default) -> Optional[Dict[str, Any]]:
    """argparse kwargs for one field, or None when the field cannot be a command-line option."""
    if isinstance(tp, str):
        if any(s in tp for s in _SKIP_TYPES):
            return None
        tp = {"int": int, "float": float, "str": str, "bool": bool}.get(tp.replace("Optional[", "").rstrip("]"), None)
        if tp is None:
            return None
    tp, _ = _unwrap_optional(tp)
    origin = typing.get_origin(tp)
    if tp is bool:
        return dict(action=argparse.BooleanOptionalAction, default=default)
    if tp in (int, float, str):
        return dict(type=tp, default=default)
    if isinstance(tp, type) and issubclass(tp, enum.Enum):
        return dict(type=lambda s, _t=tp: _t[s] if s in _t.__members__ else _t(s), choices=list(tp), default=default)
    if origin is typing.Literal:
        choices = list(typing.get_args(tp))
        return dict(type=type(choices[0]), choices=choices, default=default)
    if origin in (list, List, tuple, Tuple, Sequence) or tp in (list, tuple):
        inner = [a for a in typing.get_args(tp) if a is not Ellipsis]
        elem, _ = _unwrap_optional(inner[0]) if inner else (str, False)
        if elem in (int, float, str):
            return dict(type=elem, nargs="*", default=default)
        return None
    if origin is typing.Union:                       # e.g. Union[int, List[int]]
        for a in typing.get_args(tp):
            if a in (int, float, str):
                return dict(type=a, default=default)
    return None


def add_dataclass_arguments(parser: argparse.ArgumentParser, cls, title: Optional[str] = None, exclude: Iterable[str] = (), prefix: str = "") -> List[str]:
    """One flag per (eligible) field of ``cls`` that the parser does not define yet.  Returns the destination names that were added."""
    existing = {a.dest for a in parser._actions}
    taken_flags = {s for a in parser._actions for s in a.option_strings}
    group = parser.add_argument_group(title or cls.__name__)
    hints = _resolve_hints(cls)
    added = []
    for f in dataclasses.fields(cls):
        dest = prefix + f.name
        flag = "--" + dest.replace("_", "-")
        if f.name in exclude or dest in existing or flag in taken_flags or not f.init:
            continue
        default = None if f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING else (
            f.default if f.default is not dataclasses.MISSING else f.default_factory())
        kw = field_to_argparse_kwargs(f.name, hints.get(f.name, f.type), default)
        if kw is None:
            continue
        help_ = (f.metadata or {}).get("help") or f"{cls.__name__}.{f.name}"
        group.add_argument(flag, dest=dest, help=help_, **kw)
        added.append(dest)
    return added


def dataclass_from_args(cls, args, prefix: str = "", **overrides):
    """Build ``cls`` from the fields present (and not None-for-required) in ``args``; ``overrides`` win."""
    kw = {}
    for f in dataclasses.fields(cls):
        if not f.init:
            continue
        dest = prefix + f.name
        if f.name in overrides:
            kw[f.name] = overrides[f.name]
        elif hasattr(args, dest):
            v = getattr(args, dest)
            required = f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING
            if v is None and not required and f.default is not None:
                continue          # "not given": keep the dataclass default
            kw[f.name] = v
    return cls(**kw)


# ---- YAML ---------------------------------------------------------------------------------------------------------------------------------------
def _flatten(d: Dict[str, Any], out: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    out = {} if out is None else out
    for k, v in d.items():
        if isinstance(v, dict):
            _flatten(v, out)
        else:
            key = str(k).replace("-", "_")
            if key in out and out[key] != v:
                raise ValueError(f"YAML config sets {key!r} twice with different values ({out[key]!r} and {v!r})")
            out[key] = v
    return out


def load_yaml_config(path: str) -> Dict[str, Any]:
    import yaml

    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    if not isinstance(raw, dict):
        raise ValueError(f"{path}: top level of a config file must be a mapping")
    return _flatten(raw)


def apply_yaml(args: argparse.Namespace, cfg: Dict[str, Any], parser: Optional[argparse.ArgumentParser] = None, explicit: Iterable[str] = ()) -> argparse.Namespace:
    """Set ``args.<key> = value`` for every YAML leaf.  Unknown keys are an error (a typo must not silently train the wrong model); keys given explicitly on the
    command line (``explicit`` dest names) keep their command-line value; values go through the parser's type conversion when the parser is given."""
    explicit = set(explicit)
    actions = {a.dest: a for a in parser._actions} if parser is not None else {}
    for k, v in cfg.items():
        if not hasattr(args, k):
            raise ValueError(f"unknown configuration key {k!r} in the YAML file")
        if k in explicit:
            continue
        a = actions.get(k)
        if a is not None and a.type is not None and v is not None:
            v = [a.type(x) for x in v] if isinstance(v, list) else a.type(v)
        if a is not None and a.choices is not None and v is not None and v not in a.choices:
            raise ValueError(f"{k}: {v!r} is not one of {list(a.choices)}")
        setattr(args, k, v)
    return args


def explicit_dests(parser: argparse.ArgumentParser, argv: Sequence[str]) -> List[str]:
    """Destination names of the options that literally appear in ``argv``."""
    by_flag = {s: a.dest for a in parser._actions for s in a.option_strings}
    return [by_flag[t.split("=")[0]] for t in argv if t.startswith("--") and t.split("=")[0] in by_flag]


def args_to_yaml(args: argparse.Namespace, classes: Sequence[type] = ()) -> str:
    """Effective configuration grouped by owning config class (everything else under ``run``)."""
    import yaml

    left = {k: v for k, v in vars(args).items() if not k.startswith("_")}
    doc: Dict[str, Dict[str, Any]] = {}
    for cls in classes:
        names = [f.name for f in dataclasses.fields(cls) if f.name in left]
        doc[cls.__name__] = {n: _plain(left.pop(n)) for n in names}
    doc["run"] = {k: _plain(v) for k, v in sorted(left.items())}
    return yaml.safe_dump(doc, sort_keys=False, default_flow_style=False)


def _plain(v):
    if isinstance(v, enum.Enum):
        return v.name
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (int, float, str, bool)) or v is None:
        return v
    return str(v)
