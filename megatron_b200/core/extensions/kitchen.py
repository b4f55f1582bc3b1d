"""Kitchen (NVIDIA-internal quantisation playground) adapter — reference ``extensions/kitchen.py`` is itself a stub because the
package is not public.  ``HAVE_KITCHEN`` is False; recipes that would come from it map to ``megatron_b200.core.quantization`` /
``post_training.quantize`` (per-layer YAML-matched fp8 / mxfp8 / nvfp4 linears on the in-tree block-scaled GEMMs)."""
HAVE_KITCHEN = False


def _unavailable(name):
    def f(*a, **k):
        raise RuntimeError(f"kitchen.{name} is not available (the package is not public); use megatron_b200.core.quantization recipes instead")
    f.__name__ = name
    return f


KitchenSpecProvider = _unavailable("KitchenSpecProvider")
QuantizeRecipe = _unavailable("QuantizeRecipe")
get_qlinear_params_from_predefined = _unavailable("get_qlinear_params_from_predefined")
