#!/usr/bin/env python
"""Pretrain a hybrid Mamba-2 / attention / MLP language model; drop-in for the reference's ``pretrain_mamba.py`` / ``pretrain_hybrid.py``.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 pretrain_mamba.py --num-layers 8 --hidden-size 1024 --num-attention-heads 8 \
        --hybrid-override-pattern 'M-M-M*-M' --seq-length 2048 --micro-batch-size 2 --global-batch-size 16 --train-iters 100 --mock-data --vocab-size 32000
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import pretrain_gpt as base  # noqa: E402
from megatron_b200.core.models.mamba.mamba_layer_specs import mamba_stack_spec  # noqa: E402
from megatron_b200.core.models.mamba.mamba_model import MambaModel  # noqa: E402
from megatron_b200.training.arguments import core_transformer_config_from_args  # noqa: E402
from megatron_b200.training.training import get_args, pretrain  # noqa: E402


def _extra_args(parser):
    g = parser.add_argument_group("hybrid")
    g.add_argument("--hybrid-override-pattern", default=None, help="per-layer symbols: M mamba, * attention, - MLP")
    g.add_argument("--hybrid-attention-ratio", type=float, default=0.0)
    g.add_argument("--hybrid-mlp-ratio", type=float, default=0.0)
    g.add_argument("--mamba-state-dim", type=int, default=128)
    g.add_argument("--mamba-head-dim", type=int, default=64)
    g.add_argument("--mamba-num-groups", type=int, default=8)
    return parser


def model_provider(pre_process=True, post_process=True, vp_stage=None):
    args = get_args()
    config = core_transformer_config_from_args(args)
    config.mamba_state_dim, config.mamba_head_dim, config.mamba_num_groups = args.mamba_state_dim, args.mamba_head_dim, args.mamba_num_groups
    return MambaModel(config=config, mamba_stack_spec=mamba_stack_spec, vocab_size=args.padded_vocab_size,
                      max_sequence_length=args.max_position_embeddings or args.seq_length, pre_process=pre_process, post_process=post_process,
                      hybrid_attention_ratio=args.hybrid_attention_ratio, hybrid_mlp_ratio=args.hybrid_mlp_ratio,
                      hybrid_override_pattern=args.hybrid_override_pattern, parallel_output=True,
                      share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
                      position_embedding_type="none" if args.position_embedding_type == "learned_absolute" else args.position_embedding_type,
                      rotary_percent=args.rotary_percent, rotary_base=args.rotary_base)


if __name__ == "__main__":
    pretrain(base.train_valid_test_datasets_provider, model_provider, base.forward_step, extra_args_provider=_extra_args,
             args_defaults={"tokenizer_type": "NullTokenizer"})
