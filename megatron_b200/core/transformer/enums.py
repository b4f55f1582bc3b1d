from ..enums import AttnBackend, AttnMaskType, AttnType, LayerType, ModelType  # noqa: F401
