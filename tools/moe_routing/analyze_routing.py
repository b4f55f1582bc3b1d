#!/usr/bin/env python
"""Offline analysis of ``RouterTracer`` output (reference ``tools/moe_routing/analyze_routing*.py``).

    python tools/moe_routing/analyze_routing.py /traces/rank_00000 --num-experts 64 [--json out.json]

Per layer: the load of every expert (fraction of routed slots), the max/mean imbalance, the normalised entropy of the load
distribution (1 = uniform), the share of slots taken by the busiest 10 % of experts (concentration), and — when a layer was
traced over several steps — how *predictable* routing is: the fraction of (position, slot) decisions that equal the previous
step's (high values mean a cached routing / replay would rarely be wrong).
"""
import argparse
import json
import math
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from megatron_b200.core.transformer.moe.router_trace import load_indices_for_record  # noqa: E402


def load_index(trace_dir):
    with open(os.path.join(trace_dir, "index.jsonl")) as f:
        return [json.loads(l) for l in f]


def analyze(trace_dir: str, num_experts: int):
    by_layer = defaultdict(list)
    for rec in load_index(trace_dir):
        by_layer[(rec["block"], rec["mtp_index"], rec["layer"])].append(rec)
    report = {}
    for key, recs in sorted(by_layer.items(), key=lambda kv: str(kv[0])):
        recs.sort(key=lambda r: (r["step"], r["microbatch"]))
        counts = torch.zeros(num_experts, dtype=torch.float64)
        prev, same, seen = None, 0, 0
        for r in recs:
            idx = load_indices_for_record(r, trace_dir).long()
            valid = idx >= 0
            counts += torch.bincount(idx[valid].flatten(), minlength=num_experts).double()
            if prev is not None and prev.shape == idx.shape:
                same += int(((prev == idx) & valid).sum())
                seen += int(valid.sum())
            prev = idx
        load = counts / counts.sum().clamp(min=1)
        nz = load[load > 0]
        entropy = float(-(nz * nz.log()).sum() / math.log(num_experts)) if num_experts > 1 else 1.0
        top = max(1, num_experts // 10)
        report["/".join(str(k) for k in key if k is not None)] = {
            "records": len(recs), "slots": int(counts.sum()), "load": [round(x, 6) for x in load.tolist()],
            "imbalance_max_over_mean": float(load.max() * num_experts), "entropy": entropy,
            "top10pct_share": float(load.sort(descending=True).values[:top].sum()), "dead_experts": int((counts == 0).sum()),
            "repeat_fraction": (same / seen) if seen else None,
        }
    return report


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("trace_dir")
    ap.add_argument("--num-experts", type=int, required=True)
    ap.add_argument("--json", default=None)
    a = ap.parse_args(argv)
    rep = analyze(a.trace_dir, a.num_experts)
    for layer, r in rep.items():
        rf = "-" if r["repeat_fraction"] is None else f"{r['repeat_fraction']:.3f}"
        print(f"{layer:24s} slots {r['slots']:9d}  imbalance {r['imbalance_max_over_mean']:.2f}  entropy {r['entropy']:.3f}  top10% {r['top10pct_share']:.3f}  "
              f"dead {r['dead_experts']:3d}  repeat {rf}")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rep, f, indent=1)
    return rep


if __name__ == "__main__":
    main()
