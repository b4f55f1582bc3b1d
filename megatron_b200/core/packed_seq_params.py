"""THD (packed variable-length) attention parameters (reference ``packed_seq_params.py``)."""
from dataclasses import dataclass
from typing import Optional

from torch import Tensor


@dataclass
class PackedSeqParams:
    qkv_format: str = None
    cu_seqlens_q: Tensor = None
    cu_seqlens_kv: Tensor = None
    cu_seqlens_q_padded: Tensor = None
    cu_seqlens_kv_padded: Tensor = None
    max_seqlen_q: int = None
    max_seqlen_kv: int = None
    local_cp_size: Optional[int] = None
    cp_group: object = None
