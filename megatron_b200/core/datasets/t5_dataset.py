"""Import-path parity (reference ``datasets/t5_dataset.py``): the T5 span-corruption dataset lives in ``masked_dataset.py``."""
from .masked_dataset import MaskedDatasetConfig as T5MaskedDatasetConfig, T5MaskedDataset  # noqa: F401
