#!/usr/bin/env python
"""Per-GPU memory estimate for a preset under a parallel layout (reference ``tools/report_theoretical_memory.py``)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatron_b200.models.presets import PRESETS  # noqa: E402
from megatron_b200.training.theoretical_memory import report  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3_8b", choices=sorted(PRESETS))
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--dp", type=int, default=1)
    ap.add_argument("--ep", type=int, default=1)
    ap.add_argument("--micro-batch", type=int, default=1)
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--fp32-grads", action="store_true")
    ap.add_argument("--no-dist-opt", action="store_true")
    ap.add_argument("--recompute", default="selective", choices=["none", "selective", "full"])
    a = ap.parse_args()
    print(report(a.model, a.tp, a.pp, a.dp, a.ep, a.micro_batch, a.seq, a.fp32_grads, not a.no_dist_opt, a.recompute))
