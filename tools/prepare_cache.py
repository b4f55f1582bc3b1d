#!/usr/bin/env python
"""Build the GPT dataset index caches (document / sample / shuffle indices of every blend component and split) ahead of a training job (reference
``tools/prepare_cache.py``): a one-process CPU job on the login node instead of minutes of rank-0 work while N GPUs wait.

Takes the SAME arguments as ``pretrain_gpt.py`` (sizes, ``--data-path`` / ``--train-data-path`` …, ``--split``, ``--seed``, ``--train-iters``, ``--global-batch-size``,
``--eval-iters`` / ``--eval-interval``, ``--data-cache-path``) because the cache keys hash the dataset configuration and the requested sample counts.

    python tools/prepare_cache.py --data-path corpus_text_document --data-cache-path /cache --seq-length 8192 --train-iters 1000 --global-batch-size 128 ...
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    from megatron_b200.core.datasets import BlendedMegatronDatasetBuilder, GPTDatasetConfig
    from megatron_b200.core.datasets.gpt_dataset import GPTDataset
    from megatron_b200.core.datasets.utils import get_blend_from_list
    from megatron_b200.core.tokenizers import build_tokenizer_from_args
    from megatron_b200.training.arguments import parse_args, validate_args
    from megatron_b200.training.training import get_train_valid_test_num_samples

    args = parse_args(argv)
    if args.tokenizer_type is None:
        args.tokenizer_type = "NullTokenizer"
    validate_args(args, world_size=max(args.tensor_model_parallel_size * args.pipeline_model_parallel_size * args.context_parallel_size, 1))
    for bad in ("mock_data", "fim_data", "sft"):
        if getattr(args, bad, False):
            raise SystemExit(f"prepare_cache: --{bad.replace('_', '-')} has nothing to cache")
    if not args.data_cache_path:
        raise SystemExit("prepare_cache: --data-cache-path is required (that is where the indices go)")
    tokenizer = build_tokenizer_from_args(args)
    per_split = [getattr(args, k, None) for k in ("train_data_path", "valid_data_path", "test_data_path")]
    use_per_split = any(per_split)
    cfg = GPTDatasetConfig(
        random_seed=args.seed, sequence_length=args.seq_length, blend=None if use_per_split else get_blend_from_list(args.data_path), split=None if use_per_split else args.split,
        blend_per_split=[get_blend_from_list(p) if p else None for p in per_split] if use_per_split else None, path_to_cache=args.data_cache_path, tokenizer=tokenizer,
        reset_position_ids=args.reset_position_ids, reset_attention_mask=args.reset_attention_mask, eod_mask_loss=args.eod_mask_loss, create_attention_mask=False,
        mmap_bin_files=getattr(args, "mmap_bin_files", True))
    counts = get_train_valid_test_num_samples(args)
    t0 = time.time()
    splits = BlendedMegatronDatasetBuilder(GPTDataset, counts, lambda: True, cfg).build()
    files = sorted(os.listdir(args.data_cache_path))
    report = {"requested_samples": dict(zip(("train", "valid", "test"), counts)), "built": {n: (len(d) if d is not None else None) for n, d in zip(("train", "valid", "test"), splits)},
              "cache_dir": args.data_cache_path, "cache_files": len(files), "seconds": round(time.time() - t0, 2)}
    print(json.dumps(report))
    return report


if __name__ == "__main__":
    main()
