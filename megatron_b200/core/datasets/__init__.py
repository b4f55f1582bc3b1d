from .blended_megatron_dataset_builder import BlendedMegatronDatasetBuilder
from .gpt_dataset import GPTDataset, GPTDatasetConfig, MockGPTDataset
from .indexed_dataset import IndexedDataset, IndexedDatasetBuilder
from .utils import Split, compile_helpers
