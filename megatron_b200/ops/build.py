"""In-tree build of the native extension ``megatron_b200/ops/_C*.so`` for sm_100a.

Kernels are compiled with nvcc directly (``-gencode arch=compute_100a,code=sm_100a
-lineinfo``) — no torch arch list involved, so the build cross-compiles on a box without a
GPU.  Kernel translation units do not include torch headers (seconds each); only
``bindings.cpp`` does.  ``python -m megatron_b200.ops.build [--force] [--verbose]``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "build"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

KERNEL_SOURCES = ["norm.cu", "elementwise.cu", "cross_entropy.cu", "multi_tensor.cu", "gemm_sm100.cu"]
OPTIONAL_SOURCES = ["flash_attn_sm100.cu", "nvlink_collectives.cu", "fused_tp_gemm.cu", "grouped_gemm_sm100.cu", "moe_kernels.cu", "runtime_native.cu", "gemm_fp8_sm100.cu", "flash_attn_bwd_sm100.cu", "extra_kernels.cu", "gemm_mxfp8_sm100.cu", "gemm_nvfp4_sm100.cu", "paged_attention.cu", "routing_kernels.cu", "misc_kernels.cu"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def target_path() -> Path:
    return HERE / f"_C{_ext_suffix()}"


def _hash_sources(files) -> str:
    h = hashlib.sha256()
    for f in sorted(files):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, verbose, log: Path):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.write_text(" ".join(map(str, cmd)) + "\n" + r.stdout)
    if verbose:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd))}\n{r.stdout}")
    return r.stdout


def build_all(force: bool = False, verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension

    sources = [CSRC / s for s in KERNEL_SOURCES] + [CSRC / s for s in OPTIONAL_SOURCES if (CSRC / s).exists()]
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    binding = CSRC / "bindings.cpp"
    digest = _hash_sources(sources + headers + [binding])
    stamp = BUILD / "stamp.json"
    tgt = target_path()
    if not force and tgt.exists() and stamp.exists() and json.loads(stamp.read_text()).get("digest") == digest:
        return tgt
    BUILD.mkdir(exist_ok=True)
    nvcc = _nvcc()
    cuda_home = Path(nvcc).resolve().parent.parent
    optional_defs = [f"-DMB200_HAVE_{Path(s).stem.upper()}" for s in OPTIONAL_SOURCES if (CSRC / s).exists()]

    def compile_cu(src: Path):
        obj = BUILD / (src.stem + ".o")
        _run([nvcc, *ARCH_FLAGS, *NVCC_FLAGS, *optional_defs, "-I", str(CSRC), "-c", str(src), "-o", str(obj)], verbose, BUILD / (src.stem + ".log"))
        return obj

    def compile_binding():
        obj = BUILD / "bindings.o"
        inc = []
        for p in cpp_extension.include_paths():
            inc += ["-isystem", p]
        inc += ["-isystem", sysconfig.get_paths()["include"], "-isystem", str(cuda_home / "include"), "-I", str(CSRC)]
        abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
               *optional_defs, *inc, "-c", str(binding), "-o", str(obj)]
        _run(cmd, verbose, BUILD / "bindings.log")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources) + 1)) as ex:
        futs = [ex.submit(compile_cu, s) for s in sources] + [ex.submit(compile_binding)]
        objs = [f.result() for f in futs]

    torch_lib = Path(torch.__file__).parent / "lib"
    link = [nvcc, "-shared", *ARCH_FLAGS, "-o", str(tgt), *map(str, objs), f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            "-ltorch_python", "-Xlinker", f"-rpath={torch_lib}"]
    _run(link, verbose, BUILD / "link.log")
    stamp.write_text(json.dumps({"digest": digest}))
    return tgt


def build_datasets_helpers() -> str:
    """The C++ dataset index builders (``core/datasets/helpers.cpp``, reference N1) — g++ + pybind11, in-tree."""
    from ..core.datasets.utils import compile_helpers

    return compile_helpers()


def ptxas_report() -> str:
    """Registers / spills / smem per kernel, scraped from the last build logs."""
    out = []
    for log in sorted(BUILD.glob("*.log")):
        txt = log.read_text()
        if "ptxas info" in txt:
            out.append(f"== {log.stem} ==")
            out += [ln for ln in txt.splitlines() if "Used" in ln or "Compiling entry" in ln or "spill" in ln]
    return "\n".join(out)


if __name__ == "__main__":
    p = build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
    if "--report" in sys.argv:
        print(ptxas_report())
