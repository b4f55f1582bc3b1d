#!/usr/bin/env python
"""Pretrain BERT (masked LM + sentence-order head); drop-in for the reference's ``pretrain_bert.py``.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 pretrain_bert.py --num-layers 12 --hidden-size 768 --num-attention-heads 12 \
        --seq-length 512 --micro-batch-size 4 --global-batch-size 32 --train-iters 100 --lr 1e-4 --mock-data --vocab-size 30522
"""
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from megatron_b200.core.datasets.masked_dataset import BERTMaskedDataset, MaskedDatasetConfig  # noqa: E402
from megatron_b200.core.models.bert.bert_layer_specs import bert_layer_local_spec  # noqa: E402
from megatron_b200.core.models.bert.bert_model import BertModel  # noqa: E402
from megatron_b200.training.arguments import core_transformer_config_from_args  # noqa: E402
from megatron_b200.training.data import get_batch_on_this_tp_rank  # noqa: E402
from megatron_b200.training.training import get_args, pretrain, print_rank_0  # noqa: E402

KEYS = ("text", "types", "labels", "is_random", "loss_mask", "padding_mask")


def model_provider(pre_process=True, post_process=True, vp_stage=None):
    args = get_args()
    config = core_transformer_config_from_args(args)
    return BertModel(config=config, num_tokentypes=2, transformer_layer_spec=bert_layer_local_spec, vocab_size=args.padded_vocab_size,
                     max_sequence_length=args.max_position_embeddings or args.seq_length, pre_process=pre_process, post_process=post_process,
                     share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights, parallel_output=True, add_binary_head=True)


def loss_func(loss_mask, sentence_order, output_tensor):
    lm_loss_, sop_logits = output_tensor
    lm = torch.sum(lm_loss_.float().view(-1) * loss_mask.view(-1)) / loss_mask.sum().clamp(min=1)
    if sop_logits is not None:
        sop = torch.nn.functional.cross_entropy(sop_logits.view(-1, 2).float(), sentence_order.view(-1), ignore_index=-1)
        return lm + sop, {"lm loss": lm.detach(), "sop loss": sop.detach()}
    return lm, {"lm loss": lm.detach()}


def forward_step(data_iterator, model):
    b = get_batch_on_this_tp_rank(data_iterator, keys=KEYS)
    out = model(b["text"], b["padding_mask"], tokentype_ids=b["types"], lm_labels=b["labels"])
    return out, partial(loss_func, b["loss_mask"].float(), b["is_random"])


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    print_rank_0("> building BERT datasets (synthetic token stream unless --data-path is given) ...")
    cfg = MaskedDatasetConfig(sequence_length=args.seq_length, vocab_size=args.padded_vocab_size, random_seed=args.seed)
    indexed = None
    if not args.mock_data and args.data_path:
        from megatron_b200.core.datasets.indexed_dataset import IndexedDataset

        indexed = IndexedDataset(args.data_path[-1])
    return tuple(BERTMaskedDataset(cfg, indexed, max(n, 1)) for n in num_samples)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, forward_step, args_defaults={"tokenizer_type": "NullTokenizer", "position_embedding_type": "learned_absolute"})
