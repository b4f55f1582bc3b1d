#!/usr/bin/env python
"""Tokenise JSONL documents into the ``.bin/.idx`` indexed-dataset format (reference ``tools/preprocess_data.py``).

    python tools/preprocess_data.py --input corpus.jsonl --output-prefix out/corpus --tokenizer-type NullTokenizer \
        --vocab-size 50256 --append-eod --workers 8
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatron_b200.core.datasets.indexed_dataset import DType, IndexedDatasetBuilder  # noqa: E402
from megatron_b200.core.tokenizers import build_tokenizer  # noqa: E402

_TOK = None


def _init(args):
    global _TOK
    _TOK = build_tokenizer(args.tokenizer_type, vocab_size=args.vocab_size, tokenizer_model=args.tokenizer_model)


def _encode(line_and_args):
    line, key, append_eod = line_and_args
    try:
        text = json.loads(line)[key]
    except Exception:
        return None, len(line)
    ids = _TOK.tokenize(text)
    if append_eod:
        ids.append(_TOK.eod)
    return ids, len(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True)
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--json-key", default="text")
    ap.add_argument("--tokenizer-type", default="NullTokenizer")
    ap.add_argument("--tokenizer-model", default=None)
    ap.add_argument("--vocab-size", type=int, default=None)
    ap.add_argument("--append-eod", action="store_true")
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--log-interval", type=int, default=10000)
    args = ap.parse_args()
    _init(args)
    dtype = DType.optimal_dtype(_TOK.vocab_size)
    os.makedirs(os.path.dirname(os.path.abspath(args.output_prefix)), exist_ok=True)
    builder = IndexedDatasetBuilder(f"{args.output_prefix}_{args.json_key}_document.bin", dtype=dtype)
    t0, nbytes, ndocs = time.time(), 0, 0
    with open(args.input, encoding="utf-8") as f:
        items = ((line, args.json_key, args.append_eod) for line in f)
        if args.workers > 1:
            pool = mp.Pool(args.workers, initializer=_init, initargs=(args,))
            results = pool.imap(_encode, items, 32)
        else:
            results = map(_encode, items)
        for ids, n in results:
            nbytes += n
            if not ids:
                continue
            builder.add_item(np.asarray(ids, dtype=dtype))
            builder.end_document()
            ndocs += 1
            if ndocs % args.log_interval == 0:
                dt = time.time() - t0
                print(f"processed {ndocs} documents ({ndocs / dt:.1f} docs/s, {nbytes / dt / 2**20:.2f} MiB/s)", flush=True)
    builder.finalize(f"{args.output_prefix}_{args.json_key}_document.idx")
    print(f"done: {ndocs} documents → {args.output_prefix}_{args.json_key}_document.{{bin,idx}}")


if __name__ == "__main__":
    main()
