"""Import-path parity: the config lives in ``megatron_dataset.py`` (reference ``datasets/blended_megatron_dataset_config.py``)."""
from .megatron_dataset import BlendedMegatronDatasetConfig, convert_split_vector_to_split_matrix, parse_and_normalize_split  # noqa: F401
