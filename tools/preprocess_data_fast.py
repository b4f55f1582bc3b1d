#!/usr/bin/env python
"""Parallel JSONL → ``.bin/.idx`` tokenisation without an inter-process token pipe (reference ``tools/preprocess_data_fast.py``,
which delegates to the external ``gigatoken`` library; the same job is done here with byte-range sharding).

``preprocess_data.py`` sends every document through a ``multiprocessing`` pipe twice (text out, token list back) and serialises the
writes in the parent.  Here the input file is cut into ``workers × chunks_per_worker`` byte ranges aligned to line starts; every
worker tokenises its ranges and writes its OWN partial ``.bin/.idx``; the parent only concatenates the partial files in range order
(``IndexedDatasetBuilder.add_index`` = a file append + index offset fix-up).  Document order is the input order.

    python tools/preprocess_data_fast.py --input corpus.jsonl --output-prefix out/corpus --json-keys text --workers 32 \
        --tokenizer-type NullTokenizer --vocab-size 50257 --append-eod
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatron_b200.core.datasets.indexed_dataset import DType, IndexedDatasetBuilder, get_bin_path, get_idx_path  # noqa: E402
from megatron_b200.core.tokenizers import build_tokenizer  # noqa: E402


def line_aligned_ranges(path: str, n: int):
    """``n`` byte ranges ``[start, end)`` that each begin at a line start and together cover the file."""
    size = os.path.getsize(path)
    cuts = [0]
    with open(path, "rb") as f:
        for i in range(1, n):
            f.seek(max(cuts[-1], size * i // n))
            f.readline()                                    # finish the line we landed in
            pos = min(f.tell(), size)
            if pos > cuts[-1]:
                cuts.append(pos)
    cuts.append(size)
    return [(a, b) for a, b in zip(cuts, cuts[1:]) if b > a]


def _work(job):
    idx, (start, end), a = job
    tok = build_tokenizer(a["tokenizer_type"], vocab_size=a["vocab_size"], tokenizer_model=a["tokenizer_model"])
    dtype = DType.optimal_dtype(tok.vocab_size)
    builders = {k: IndexedDatasetBuilder(get_bin_path(f"{a['tmp']}_{k}_{idx:06d}"), dtype=dtype) for k in a["keys"]}
    ndocs = 0
    with open(a["input"], "rb") as f:
        f.seek(start)
        while f.tell() < end:
            line = f.readline()
            if not line.strip():
                continue
            try:
                doc = json.loads(line)
            except Exception:
                continue
            for k in a["keys"]:
                text = doc.get(k)
                if not text and not a["keep_empty"]:
                    continue
                ids = tok.tokenize(text or "")
                if a["append_eod"]:
                    ids.append(tok.eod)
                builders[k].add_item(np.asarray(ids, dtype=dtype))
                builders[k].end_document()
            ndocs += 1
    for k, b in builders.items():
        b.finalize(get_idx_path(f"{a['tmp']}_{k}_{idx:06d}"))
    return idx, ndocs, end - start


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True)
    ap.add_argument("--output-prefix", required=True)
    ap.add_argument("--json-keys", nargs="+", default=["text"])
    ap.add_argument("--tokenizer-type", default="NullTokenizer")
    ap.add_argument("--tokenizer-model", default=None)
    ap.add_argument("--vocab-size", type=int, default=None)
    ap.add_argument("--append-eod", action="store_true")
    ap.add_argument("--keep-empty", action="store_true")
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--chunks-per-worker", type=int, default=4, help="more chunks = better balance when documents vary in length")
    ap.add_argument("--level", default="document")
    args = ap.parse_args(argv)
    os.makedirs(os.path.dirname(os.path.abspath(args.output_prefix)), exist_ok=True)
    tmp = args.output_prefix + ".part"
    a = dict(tokenizer_type=args.tokenizer_type, vocab_size=args.vocab_size, tokenizer_model=args.tokenizer_model, keys=args.json_keys,
             append_eod=args.append_eod, keep_empty=args.keep_empty, input=args.input, tmp=tmp)
    ranges = line_aligned_ranges(args.input, max(1, args.workers * args.chunks_per_worker))
    jobs = [(i, r, a) for i, r in enumerate(ranges)]
    t0 = time.time()
    if args.workers > 1:
        with mp.get_context("spawn").Pool(args.workers) as pool:
            done = list(pool.imap_unordered(_work, jobs))
    else:
        done = [_work(j) for j in jobs]
    ndocs, nbytes = sum(d[1] for d in done), sum(d[2] for d in done)
    tok = build_tokenizer(args.tokenizer_type, vocab_size=args.vocab_size, tokenizer_model=args.tokenizer_model)
    dtype = DType.optimal_dtype(tok.vocab_size)
    for k in args.json_keys:
        out = f"{args.output_prefix}_{k}_{args.level}"
        builder = IndexedDatasetBuilder(get_bin_path(out), dtype=dtype)
        for i in range(len(jobs)):
            part = f"{tmp}_{k}_{i:06d}"
            builder.add_index(part)
            os.remove(get_bin_path(part)), os.remove(get_idx_path(part))
        builder.finalize(get_idx_path(out))
    dt = time.time() - t0
    print(f"done: {ndocs} documents, {nbytes / 2**20:.1f} MiB in {dt:.1f} s ({ndocs / dt:.0f} docs/s, {nbytes / dt / 2**20:.1f} MiB/s) with {args.workers} workers")
    return ndocs


if __name__ == "__main__":
    main()
