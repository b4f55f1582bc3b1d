"""Import-path parity (reference ``datasets/bert_dataset.py``): the BERT masked-LM dataset lives in ``masked_dataset.py``."""
from .masked_dataset import BERTMaskedDataset, MaskedDatasetConfig as BERTMaskedDatasetConfig  # noqa: F401
