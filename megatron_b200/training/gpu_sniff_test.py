"""Quick GPU / fabric health probe run before a job (reference ``training/gpu_sniff_test.py``).

Per rank: HBM copy bandwidth, bf16 GEMM throughput (our tcgen05 kernel AND cuBLAS), and — with more than one rank — an NVLink
all-reduce bus bandwidth.  Every value is compared with the min/median over ranks so a slow GPU or link stands out.
``python -m megatron_b200.training.gpu_sniff_test`` (works under torchrun)."""
from __future__ import annotations

import json
import os

import torch
import torch.distributed as dist


def _time(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def sniff(thresholds=None) -> dict:
    thresholds = thresholds or {"copy_TBps": 4.0, "gemm_TFLOPs": 900.0, "allreduce_GBps": 200.0}
    out = {"rank": dist.get_rank() if dist.is_initialized() else 0, "device": torch.cuda.get_device_name()}
    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    out["copy_TBps"] = round(2 * a.numel() / _time(lambda: b.copy_(a)) / 1e12, 2)
    m = 8192
    x, w = torch.randn(m, m, device="cuda").bfloat16(), torch.randn(m, m, device="cuda").bfloat16()
    out["cublas_TFLOPs"] = round(2 * m**3 / _time(lambda: torch.matmul(x, w.t())) / 1e12, 1)
    try:
        from .. import ops

        out["gemm_TFLOPs"] = round(2 * m**3 / _time(lambda: ops.gemm_nt(x, w)) / 1e12, 1)
    except Exception as e:  # extension missing
        out["gemm_TFLOPs"] = 0.0
        out["gemm_error"] = str(e)[:100]
    if dist.is_initialized() and dist.get_world_size() > 1:
        n = dist.get_world_size()
        t = torch.empty(256 << 20, dtype=torch.bfloat16, device="cuda")
        sec = _time(lambda: dist.all_reduce(t), iters=5)
        out["allreduce_GBps"] = round(2 * (n - 1) / n * t.numel() * 2 / sec / 1e9, 1)
    out["ok"] = all(out.get(k, v) >= v for k, v in thresholds.items() if k in out)
    return out


def main():
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    res = sniff()
    if dist.is_initialized():
        allr = [None] * dist.get_world_size()
        dist.all_gather_object(allr, res)
        if dist.get_rank() == 0:
            print(json.dumps(allr, indent=1))
    else:
        print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
