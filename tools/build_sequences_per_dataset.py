#!/usr/bin/env python
"""Write ``{dataset prefix: (num sequences, num documents)}`` for a blend so training can skip opening every ``.idx`` to size the
blend (reference ``tools/build_sequences_per_dataset.py``; consumed through ``--per-dataset-sequences-path``).

    python tools/build_sequences_per_dataset.py --data-path 0.3 /data/a 0.7 /data/b --per-dataset-sequences-path seqs.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatron_b200.core.datasets.indexed_dataset import _IndexReader, get_idx_path  # noqa: E402
from megatron_b200.core.datasets.utils import get_blend_from_list  # noqa: E402


def paths_from_args(a):
    paths = []
    for lst in (a.data_path, a.train_data_path, a.valid_data_path, a.test_data_path):
        if lst:
            prefixes, _ = get_blend_from_list(lst)
            paths += prefixes
    if a.per_split_data_args_path:
        with open(a.per_split_data_args_path) as f:
            per = json.load(f)
        for split in ("train", "valid", "test"):
            if per.get(split):
                paths += get_blend_from_list(per[split])[0]
    return sorted(set(paths))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-path", nargs="*")
    ap.add_argument("--train-data-path", nargs="*")
    ap.add_argument("--valid-data-path", nargs="*")
    ap.add_argument("--test-data-path", nargs="*")
    ap.add_argument("--per-split-data-args-path")
    ap.add_argument("--per-dataset-sequences-path", required=True)
    a = ap.parse_args(argv)
    out = {}
    for p in paths_from_args(a):
        idx = _IndexReader(get_idx_path(p), multimodal=False)
        out[p] = (int(len(idx.sequence_lengths)), int(len(idx.document_indices) - 1))
    with open(a.per_dataset_sequences_path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote sequence counts of {len(out)} datasets to {a.per_dataset_sequences_path}")
    return out


if __name__ == "__main__":
    main()
