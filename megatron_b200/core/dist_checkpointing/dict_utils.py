"""Nested dict/list helpers (reference ``dist_checkpointing/dict_utils.py``)."""
from __future__ import annotations

from typing import Callable, Iterable, Tuple

import torch


def nested_values(x):
    it = x.values() if isinstance(x, dict) else x
    for v in it:
        if isinstance(v, (dict, list)):
            yield from nested_values(v)
        else:
            yield v


def nested_items_iter(x):
    it = x.items() if isinstance(x, dict) else enumerate(x)
    for k, v in it:
        if isinstance(v, (dict, list)):
            yield from nested_items_iter(v)
        else:
            yield x, k, v


def dict_list_map_inplace(f: Callable, x):
    if isinstance(x, dict):
        for k, v in x.items():
            x[k] = dict_list_map_inplace(f, v)
    elif isinstance(x, list):
        x[:] = [dict_list_map_inplace(f, v) for v in x]
    else:
        return f(x)
    return x


def dict_list_map_outplace(f: Callable, x):
    if isinstance(x, dict):
        return {k: dict_list_map_outplace(f, v) for k, v in x.items()}
    if isinstance(x, list):
        return [dict_list_map_outplace(f, v) for v in x]
    return f(x)


def extract_matching_values(x, predicate: Callable, return_lists_as_dicts: bool = False):
    """Split a nested structure into (matching, non-matching) with the same nesting."""
    if isinstance(x, dict):
        m, n = {}, {}
        for k, v in x.items():
            if isinstance(v, (dict, list)):
                a, b = extract_matching_values(v, predicate, return_lists_as_dicts)
                if a:
                    m[k] = a
                if b or not a:
                    n[k] = b
            elif predicate(v):
                m[k] = v
            else:
                n[k] = v
        return m, n
    if isinstance(x, list):
        m = {} if return_lists_as_dicts else []
        n = {} if return_lists_as_dicts else []
        for i, v in enumerate(x):
            if isinstance(v, (dict, list)) and v:
                a, b = extract_matching_values(v, predicate, return_lists_as_dicts)
                if a:
                    if return_lists_as_dicts:
                        m[i] = a
                    else:
                        m.append(a)
                if b or not a:
                    if return_lists_as_dicts:
                        n[i] = b
                    else:
                        n.append(b)
            else:
                tgt = m if predicate(v) else n
                if return_lists_as_dicts:
                    tgt[i] = v
                else:
                    tgt.append(v)
        return m, n
    raise ValueError(f"unexpected top-level object type {type(x)}")


def merge(x1, x2, key: Tuple = ()):
    """Merge ``x2`` into ``x1`` recursively (dicts by key, lists element-wise)."""
    if isinstance(x1, dict) and isinstance(x2, dict):
        for k, v2 in x2.items():
            if k not in x1:
                x1[k] = v2
            else:
                x1[k] = merge(x1[k], v2, key + (k,))
    elif isinstance(x1, list) and isinstance(x2, list):
        if len(x1) != len(x2):
            raise ValueError(f"cannot merge lists of different length at {key}")
        for i, v2 in enumerate(x2):
            x1[i] = merge(x1[i], v2, key + (i,))
    elif isinstance(x1, list) and isinstance(x2, dict):
        for k, v2 in x2.items():
            x1[k] = merge(x1[k], v2, key + (k,)) if k < len(x1) and isinstance(x1[k], (dict, list)) else v2
    else:
        raise ValueError(f"duplicate non-dict and non-list values at {key}: {type(x1)} vs {type(x2)}")
    return x1


def map_reduce(xs: Iterable, key_fn: Callable, value_fn: Callable = lambda x: x, reduce_fn: Callable = lambda x: x) -> dict:
    out = {}
    for x in xs:
        out.setdefault(key_fn(x), []).append(value_fn(x))
    return {k: reduce_fn(v) for k, v in out.items()}


def diff(x1, x2, prefix: Tuple = ()):
    """(only_left, only_right, mismatch) key paths."""
    mismatch, only_left, only_right = [], [], []
    if isinstance(x1, dict) and isinstance(x2, dict):
        only_left = [prefix + (k,) for k in x1.keys() - x2.keys()]
        only_right = [prefix + (k,) for k in x2.keys() - x1.keys()]
        for k in x2.keys() & x1.keys():
            a, b, c = diff(x1[k], x2[k], prefix + (k,))
            only_left += a
            only_right += b
            mismatch += c
    elif isinstance(x1, (list, tuple)) and isinstance(x2, (list, tuple)):
        if len(x1) != len(x2):
            mismatch.append((prefix, len(x1), len(x2)))
        for i, (a1, a2) in enumerate(zip(x1, x2)):
            a, b, c = diff(a1, a2, prefix + (i,))
            only_left += a
            only_right += b
            mismatch += c
    else:
        if isinstance(x1, torch.Tensor) and isinstance(x2, torch.Tensor):
            ne = x1.shape != x2.shape or bool(torch.any(x1.cpu() != x2.cpu()))
        else:
            try:
                ne = bool(x1 != x2)
            except Exception:
                ne = True
        if ne:
            mismatch.append((prefix, type(x1), type(x2)))
    return only_left, only_right, mismatch
