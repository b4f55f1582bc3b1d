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


def packed_seq_params_from_documents(tokens, eod_token: int) -> "PackedSeqParams":
    """``tokens [b, s]`` holding several documents per row (separated by ``eod_token``) → the THD description of the flattened ``[1, b·s]`` row: one packed
    sequence per document (a document ends WITH its eod token; rows never share a sequence).  This is how ``--reset-attention-mask --reset-position-ids`` reaches
    the attention kernels: as ``cu_seqlens`` (a band mask inside the kernel) instead of a dense ``[b, 1, s, s]`` mask."""
    import torch

    b, s = tokens.shape
    flat = tokens.reshape(-1)
    ends = (flat == eod_token).nonzero().flatten() + 1
    rows = torch.arange(0, b * s + 1, s, device=tokens.device)
    cu = torch.unique(torch.cat([rows, ends.to(rows.dtype)]))            # sorted, duplicates (an eod at the end of a row) removed
    lens = cu[1:] - cu[:-1]
    m = int(lens.max()) if lens.numel() else s
    cu = cu.to(torch.int32)
    return PackedSeqParams(qkv_format="thd", cu_seqlens_q=cu, cu_seqlens_kv=cu, max_seqlen_q=m, max_seqlen_kv=m)
