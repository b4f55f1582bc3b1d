"""Managed-memory allocator (N5, ``core/inference/unified_memory.py`` + ``mb200_managed_malloc`` in runtime_native.cu) on a real GPU: tensors allocated inside
the pool are cudaMallocManaged memory, usable by kernels AND readable from the host without a copy."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_unified_memory_pool_allocates_managed_memory():
    from megatron_b200.core.inference import unified_memory as um

    assert um.has_unified_memory(), "managed-memory allocator failed to load from the in-tree extension"
    with um.unified_memory_pool():
        kv = torch.zeros(4, 1024, 1024, device="cuda", dtype=torch.float32)      # 16 MiB "KV cache"
    normal = torch.zeros(16, device="cuda")
    kv += 3.0
    kv[1, 2, 3] = 42.0
    y = (kv @ torch.ones(1024, 8, device="cuda")).sum()
    torch.cuda.synchronize()
    assert abs(y.item() - (3.0 * 4 * 1024 * 1024 * 8 + 39.0 * 8)) < 1.0
    cudart = pytest.importorskip("cuda.bindings.runtime")
    err_m, attr_m = cudart.cudaPointerGetAttributes(kv.data_ptr())
    err_n, attr_n = cudart.cudaPointerGetAttributes(normal.data_ptr())
    assert int(err_m) == 0 and int(err_n) == 0
    assert int(attr_m.type) == 3 and int(attr_n.type) == 2, (attr_m.type, attr_n.type)    # cudaMemoryTypeManaged / cudaMemoryTypeDevice
    # host reads the same pages directly (they migrate on demand)
    host_view = (ctypes.c_float * 8).from_address(kv.data_ptr() + 4 * (1 * 1024 * 1024 + 2 * 1024))
    assert host_view[3] == 42.0 and host_view[0] == 3.0
    kv.mul_(2.0)                                                                        # and the GPU gets them back
    torch.cuda.synchronize()
    assert kv[1, 2, 3].item() == 84.0
