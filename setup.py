"""Packaging: ``pip install -e .`` builds the sm_100a extension IN-TREE (``megatron_b200/ops/_C*.so``, same artefact as ``python -m megatron_b200.ops.build``)
and the pybind11 dataset helpers, then installs the package."""
import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class BuildNative(build_py):
    def run(self):
        sys.path.insert(0, HERE)
        try:
            from megatron_b200.ops import build as b

            b.build_all()
            b.build_datasets_helpers()
        except Exception as e:      # no nvcc: the pure-PyTorch CPU path still installs
            print(f"[setup] native build skipped: {type(e).__name__}: {e}")
        super().run()


setup(
    name="megatron_b200",
    version="0.1.0",
    description="Blackwell-native Megatron-Core: tcgen05 / TMEM / TMA kernels, NVLink-fused tensor parallelism",
    packages=find_packages(include=["megatron_b200", "megatron_b200.*"]),
    package_data={"megatron_b200.ops": ["_C*.so", "csrc/*"], "megatron_b200.core.datasets": ["*.so", "*.cpp"]},
    python_requires=">=3.10",
    install_requires=["torch>=2.6", "numpy"],
    cmdclass={"build_py": BuildNative},
)
