from .fim_dataset import FIMConfig, apply_fim, GPTFIMDataset  # noqa: F401
from .sft_dataset import SFTDataset, SFTDatasetConfig, pack_conversations  # noqa: F401
