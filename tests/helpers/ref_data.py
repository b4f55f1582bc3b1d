"""Build GPT datasets with the UNMODIFIED reference (baseline/_ref) from .bin/.idx files and dump samples (test helper, run as a script).

    argv: out.pt cache_dir seq_len seed split n_train n_valid n_test weight1 prefix1 [weight2 prefix2 ...]
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]


class _Tok:
    """The dataset only needs the end-of-document id and an identity for its cache hash."""

    def __init__(self, eod):
        self.eod, self.eod_id, self.vocab_size = eod, eod, eod + 1
        self.unique_identifiers = {"class": "TestTokenizer", "eod": eod}
        self.pad = None


def main():
    out, cache, seq, seed, split = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    counts = [int(x) for x in sys.argv[6:9]]
    blend_args = sys.argv[9:]
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29721")
    dist.init_process_group("gloo", rank=0, world_size=1)
    from megatron.core.datasets.blended_megatron_dataset_builder import BlendedMegatronDatasetBuilder
    from megatron.core.datasets.gpt_dataset import GPTDataset, GPTDatasetConfig
    from megatron.core.datasets.utils import get_blend_from_list

    cfg = GPTDatasetConfig(random_seed=seed, sequence_length=seq, blend=get_blend_from_list(blend_args), split=split, path_to_cache=cache, tokenizer=_Tok(99),
                           reset_position_ids=True, reset_attention_mask=False, eod_mask_loss=True, create_attention_mask=False, mmap_bin_files=False)
    splits = BlendedMegatronDatasetBuilder(GPTDataset, counts, lambda: True, cfg).build()
    dump = {}
    for name, ds in zip(("train", "valid", "test"), splits):
        if ds is None:
            dump[name] = None
            continue
        idx = sorted(set([0, 1, 2, len(ds) // 2, len(ds) - 1]))
        dump[name] = {"len": len(ds), "samples": {i: {k: torch.as_tensor(v).clone() for k, v in ds[i].items() if k in ("tokens", "labels", "loss_mask", "position_ids")} for i in idx}}
    torch.save(dump, out)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
