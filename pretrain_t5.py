#!/usr/bin/env python
"""Pretrain T5 with span corruption; drop-in for the reference's ``pretrain_t5.py``.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 pretrain_t5.py --num-layers 12 --hidden-size 768 --num-attention-heads 12 \
        --seq-length 512 --decoder-seq-length 128 --micro-batch-size 4 --global-batch-size 32 --train-iters 100 --mock-data --vocab-size 32128
"""
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from megatron_b200.core.datasets.masked_dataset import MaskedDatasetConfig, T5MaskedDataset  # noqa: E402
from megatron_b200.core.models.T5.t5_model import T5Model  # noqa: E402
from megatron_b200.core.models.T5.t5_spec import get_t5_decoder_with_local_block_spec, get_t5_encoder_with_local_block_spec  # noqa: E402
from megatron_b200.training.arguments import core_transformer_config_from_args  # noqa: E402
from megatron_b200.training.data import get_batch_on_this_tp_rank  # noqa: E402
from megatron_b200.training.training import get_args, pretrain, print_rank_0  # noqa: E402

KEYS = ("text_enc", "text_dec", "labels", "loss_mask", "enc_mask", "dec_mask")


def _extra_args(parser):
    g = parser.add_argument_group("t5")
    g.add_argument("--decoder-seq-length", type=int, default=128)
    g.add_argument("--decoder-num-layers", type=int, default=None)
    return parser


def model_provider(pre_process=True, post_process=True, vp_stage=None):
    args = get_args()
    config = core_transformer_config_from_args(args)
    nd = args.decoder_num_layers or args.num_layers
    return T5Model(config=config, encoder_config=config, transformer_encoder_layer_spec=get_t5_encoder_with_local_block_spec(args.num_layers),
                   transformer_decoder_layer_spec=get_t5_decoder_with_local_block_spec(nd), vocab_size=args.padded_vocab_size,
                   max_sequence_length=args.max_position_embeddings or args.seq_length, pre_process=pre_process, post_process=post_process,
                   share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights, parallel_output=True)


def loss_func(loss_mask, output_tensor):
    lm = torch.sum(output_tensor.float().view(-1) * loss_mask.view(-1)) / loss_mask.sum().clamp(min=1)
    return lm, {"lm loss": lm.detach()}


def forward_step(data_iterator, model):
    b = get_batch_on_this_tp_rank(data_iterator, keys=KEYS)
    em, dm = b["enc_mask"].float(), b["dec_mask"].float()
    enc_mask = em.unsqueeze(1) * em.unsqueeze(2)                     # [b, se, se]
    dec_mask = dm.unsqueeze(1) * dm.unsqueeze(2)                     # [b, sd, sd]  (causality added inside the model)
    x_mask = dm.unsqueeze(2) * em.unsqueeze(1)                       # [b, sd, se]
    out = model(b["text_enc"], b["text_dec"], enc_mask, dec_mask, x_mask, lm_labels=b["labels"])
    return out, partial(loss_func, b["loss_mask"].float())


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    print_rank_0("> building T5 datasets ...")
    cfg = MaskedDatasetConfig(sequence_length=args.seq_length, sequence_length_decoder=args.decoder_seq_length, vocab_size=args.padded_vocab_size, random_seed=args.seed)
    indexed = None
    if not args.mock_data and args.data_path:
        from megatron_b200.core.datasets.indexed_dataset import IndexedDataset

        indexed = IndexedDataset(args.data_path[-1])
    return tuple(T5MaskedDataset(cfg, indexed, max(n, 1)) for n in num_samples)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, forward_step, extra_args_provider=_extra_args,
             args_defaults={"tokenizer_type": "NullTokenizer", "position_embedding_type": "learned_absolute"})
