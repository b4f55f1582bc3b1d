from .module import AutoHuggingFaceModel, HuggingFaceModule, build_hf_model  # noqa: F401
