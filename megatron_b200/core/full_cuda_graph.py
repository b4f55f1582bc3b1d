"""Whole-step CUDA graph (reference ``core/full_cuda_graph.py:138-267``); implementation lives with the layer graphs."""
from .transformer.cuda_graphs import FullCudaGraphWrapper  # noqa: F401
