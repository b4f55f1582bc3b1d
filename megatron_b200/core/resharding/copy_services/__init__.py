from .base import CopyService, RecvOp, SendOp, match_local_ops_by_task_id  # noqa: F401
from .gloo_copy_service import GlooCopyService  # noqa: F401
from .nccl_copy_service import NCCLCopyService  # noqa: F401
from .nvlink_copy_service import NVLinkCopyService  # noqa: F401

# the reference's names for its one-sided services resolve to the NVLink peer-memory service
NVSHMEMCopyService = NVLinkCopyService
