"""Dataset utilities (reference ``datasets/utils.py``)."""
from __future__ import annotations

import logging
import os
import subprocess
import sysconfig
from enum import Enum
from typing import List, Optional, Tuple

import numpy

logger = logging.getLogger(__name__)


class Split(Enum):
    train = 0
    valid = 1
    test = 2


def compile_helpers():
    """Build ``helpers_cpp`` next to ``helpers.cpp`` (g++ + pybind11, in-tree)."""
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "helpers_cpp" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    src = os.path.join(here, "helpers.cpp")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    import pybind11

    cmd = ["g++", "-O3", "-Wall", "-shared", "-std=c++17", "-fPIC", "-fdiagnostics-color", f"-I{pybind11.get_include()}",
           f"-I{sysconfig.get_paths()['include']}", src, "-o", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"failed to compile the dataset helpers:\n{r.stdout}")
    return out


def normalize(weights: List[float]) -> List[float]:
    w = numpy.array(weights, dtype=numpy.float64)
    return (w / numpy.sum(w)).tolist()


def get_blend_from_list(blend: Optional[List[str]]) -> Optional[Tuple[List[str], Optional[List[float]]]]:
    """``["30", "a", "70", "b"]`` → (["a","b"], [30,70]);  ``["a","b"]`` → (["a","b"], None)."""
    if blend is None:
        return None
    if len(blend) % 2 == 1:
        return blend, None
    try:
        w = [float(x) for x in blend[0::2]]
    except ValueError:
        return blend, None
    return [p.strip() for p in blend[1::2]], w
