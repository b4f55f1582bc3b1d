#!/usr/bin/env python
"""Concatenate several ``.bin/.idx`` indexed datasets into one (reference ``tools/merge_datasets.py``).

    python tools/merge_datasets.py --input /data/shards --output-prefix /data/merged
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatron_b200.core.datasets.indexed_dataset import IndexedDataset, IndexedDatasetBuilder, get_bin_path, get_idx_path  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True, help="directory with <prefix>.bin / <prefix>.idx pairs")
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--multimodal", action="store_true")
    a = ap.parse_args(argv)
    prefixes = sorted({os.path.join(a.input, f[:-4]) for f in os.listdir(a.input) if f.endswith(".idx") and os.path.exists(os.path.join(a.input, f[:-4] + ".bin"))})
    if not prefixes:
        raise SystemExit(f"no .bin/.idx pairs under {a.input}")
    first = IndexedDataset(prefixes[0], multimodal=a.multimodal)
    builder = IndexedDatasetBuilder(get_bin_path(a.output_prefix), dtype=first.index.dtype, multimodal=a.multimodal)
    del first
    for p in prefixes:
        builder.add_index(p)
        print(f"merged {p}", flush=True)
    builder.finalize(get_idx_path(a.output_prefix))
    print(f"wrote {a.output_prefix}.bin/.idx from {len(prefixes)} datasets")


if __name__ == "__main__":
    main()
