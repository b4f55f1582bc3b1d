"""Third-party kernel-library adapters of the reference (``megatron/core/extensions/``).

This framework ships its own sm_100a kernels, so the adapters reduce to availability flags and a spec provider that returns the native
modules: ``transformer_engine`` / ``kitchen`` specs requested by a config resolve to the same local layers."""
