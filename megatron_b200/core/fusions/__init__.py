"""Reference-compatible entry points of ``megatron/core/fusions`` — each resolves to a hand-written sm_100a kernel in
``megatron_b200.ops`` (CUDA) or its PyTorch reference (CPU)."""
