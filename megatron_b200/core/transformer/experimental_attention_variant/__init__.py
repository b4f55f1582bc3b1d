from .absorbed_mla import AbsorbedMLASelfAttention, DSAMLASelfAttention  # noqa: F401
from .dsa import (  # noqa: F401
    DSAIndexer,
    DSAIndexerLossAutoScaler,
    DSAttention,
    compute_dsa_indexer_loss,
    compute_index_scores,
    is_dsa_skip_topk_layer,
    source_dsa_compute_layer,
    sparse_attention_topk,
    topk_causal_indices,
)
