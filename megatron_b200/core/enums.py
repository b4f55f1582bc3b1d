"""Enumerations (reference ``megatron/core/enums.py``, ``transformer/enums.py``)."""
import enum


class ModelType(enum.Enum):
    encoder_or_decoder = 1
    encoder_and_decoder = 2  # kept for checkpoint compatibility
    retro_encoder = 3
    retro_decoder = 4


class Fp8Recipe(str, enum.Enum):
    delayed = "delayed"
    tensorwise = "tensorwise"
    mxfp8 = "mxfp8"
    blockwise = "blockwise"
    custom = "custom"


class Fp4Recipe(str, enum.Enum):
    nvfp4 = "nvfp4"


class AttnType(enum.Enum):
    self_attn = 1
    cross_attn = 2


class AttnMaskType(enum.Enum):
    padding = 1
    causal = 2
    no_mask = 3
    padding_causal = 4
    arbitrary = 5
    causal_bottom_right = 6


class AttnBackend(enum.Enum):
    flash = 1
    fused = 2
    unfused = 3
    local = 4
    auto = 5
    b200 = 6  # our sm_100a tcgen05 attention kernel


class LayerType(enum.Enum):
    encoder = 1
    decoder = 2
    embedding = 3
    loss = 4
    mtp = 5


class CudaGraphScope(enum.Enum):
    full_iteration = 1
    attn = 2
    mlp = 3
    moe = 4
    moe_router = 5
    moe_preprocess = 6
    mamba = 7
    full = 8
