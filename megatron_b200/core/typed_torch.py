"""Typing helpers for module calls (reference ``typed_torch.py:16-222``).

``nn.Module.__call__`` is typed ``(*args: Any) -> Any``, so a type checker loses the signature of ``forward`` at every call site.
``apply_module(m)`` returns ``m.__call__`` typed as ``m.forward``; ``copy_signature(f)`` makes a wrapper advertise ``f``'s
parameters; ``not_none`` narrows optionals.  All of them are free at run time."""
from __future__ import annotations

import functools
from typing import Any, Callable, Generic, Optional, ParamSpec, Protocol, TypeVar

import torch

P = ParamSpec("P")
T = TypeVar("T")
R_co = TypeVar("R_co", covariant=True)


class _Module(Generic[P, R_co], Protocol):
    def forward(self, *args: P.args, **kwargs: P.kwargs) -> R_co:
        ...


def apply_module(m: "_Module[P, R_co]", *, check_subclass: bool = True) -> Callable[P, R_co]:
    """``apply_module(layer)(x, mask)`` ≡ ``layer(x, mask)`` (hooks included) with ``forward``'s static signature."""
    if check_subclass and not isinstance(m, torch.nn.Module):
        raise TypeError(f"{type(m).__name__} is not a torch.nn.Module")
    return m.__call__  # type: ignore[return-value, operator]


def not_none(value: Optional[T]) -> T:
    if value is None:
        raise ValueError("expected a value, got None")
    return value


def copy_signature(source: Callable[P, Any], *, drop_first: bool = False, return_type: Any = None) -> Callable[[Callable[..., T]], Callable[P, T]]:
    """Decorator: the decorated callable keeps its body but is seen (``inspect.signature``, IDEs, type checkers) with ``source``'s
    parameters — used by thin wrappers that forward ``*args, **kwargs``.  ``drop_first`` removes ``self``/``cls``."""
    import inspect

    def decorator(decorated: Callable[..., T], /) -> Callable[P, T]:
        @functools.wraps(decorated)
        def wrapper(*args, **kwargs):
            return decorated(*args, **kwargs)

        try:
            sig = inspect.signature(source)
            params = list(sig.parameters.values())
            if drop_first and params:
                params = params[1:]
            ret = sig.return_annotation if return_type is None else return_type
            wrapper.__signature__ = sig.replace(parameters=params, return_annotation=ret)  # type: ignore[attr-defined]
        except (TypeError, ValueError):
            pass
        return wrapper  # type: ignore[return-value]

    return decorator
