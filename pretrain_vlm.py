#!/usr/bin/env python
"""Pretrain a LLaVA-style vision-language model (CLIP-ViT → MLP projector → GPT); drop-in for the reference's ``pretrain_vlm.py``.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 pretrain_vlm.py --num-layers 8 --hidden-size 1024 --num-attention-heads 8 --seq-length 1024 \
        --img-h 336 --img-w 336 --patch-dim 14 --vision-num-layers 6 --vision-hidden-size 512 --micro-batch-size 1 --global-batch-size 8 --train-iters 50 --mock-data
"""
import os
import sys
from dataclasses import replace

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec  # noqa: E402
from megatron_b200.core.models.multimodal.llava_model import DEFAULT_IMAGE_TOKEN_INDEX, LLaVAModel  # noqa: E402
from megatron_b200.core.models.vision.clip_vit_model import get_num_image_embeddings  # noqa: E402
from megatron_b200.core.models.vision.vit_layer_specs import get_vit_layer_with_local_spec  # noqa: E402
from megatron_b200.core.tensor_parallel.layers import ColumnParallelLinear, RowParallelLinear  # noqa: E402
from megatron_b200.core.transformer.mlp import MLPSubmodules  # noqa: E402
from megatron_b200.training.arguments import core_transformer_config_from_args  # noqa: E402
from megatron_b200.training.data import get_batch_on_this_tp_rank  # noqa: E402
from megatron_b200.training.training import get_args, pretrain, print_rank_0  # noqa: E402


def _extra_args(parser):
    g = parser.add_argument_group("vlm")
    g.add_argument("--img-h", type=int, default=336)
    g.add_argument("--img-w", type=int, default=336)
    g.add_argument("--patch-dim", type=int, default=14)
    g.add_argument("--vision-num-layers", type=int, default=24)
    g.add_argument("--vision-hidden-size", type=int, default=1024)
    g.add_argument("--vision-num-attention-heads", type=int, default=16)
    g.add_argument("--freeze-LM", action="store_true")
    g.add_argument("--freeze-ViT", action="store_true")
    g.add_argument("--disable-vision-class-token", action="store_true")
    return parser


def model_provider(pre_process=True, post_process=True, vp_stage=None):
    args = get_args()
    lcfg = core_transformer_config_from_args(args)
    vcfg = replace(lcfg, num_layers=args.vision_num_layers, hidden_size=args.vision_hidden_size, num_attention_heads=args.vision_num_attention_heads,
                   num_query_groups=args.vision_num_attention_heads, ffn_hidden_size=4 * args.vision_hidden_size, kv_channels=args.vision_hidden_size // args.vision_num_attention_heads,
                   gated_linear_unit=False, activation_func=torch.nn.functional.gelu, normalization="LayerNorm", add_bias_linear=True, add_qkv_bias=True,
                   sequence_parallel=False, recompute_granularity=None, recompute_modules=None)
    pcfg = replace(lcfg, num_layers=1, ffn_hidden_size=lcfg.hidden_size, gated_linear_unit=False, activation_func=torch.nn.functional.gelu, add_bias_linear=True,
                   sequence_parallel=False, recompute_granularity=None, recompute_modules=None)
    n_img = get_num_image_embeddings(args.img_h, args.img_w, args.patch_dim, "clip", args.disable_vision_class_token, 1)
    model = LLaVAModel(
        language_transformer_config=lcfg, language_transformer_layer_spec=get_gpt_layer_local_spec(normalization=args.normalization), language_vocab_size=args.padded_vocab_size,
        language_max_sequence_length=(args.max_position_embeddings or args.seq_length) + n_img, vision_transformer_config=vcfg,
        vision_transformer_layer_spec=get_vit_layer_with_local_spec(), drop_vision_class_token=args.disable_vision_class_token, vision_projection_config=pcfg,
        vision_projection_layer_spec=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear), vision_projection_type="mlp",
        parallel_output=True, share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
        language_position_embedding_type=args.position_embedding_type, language_rotary_percent=args.rotary_percent, language_rotary_base=args.rotary_base,
        pre_process=pre_process, post_process=post_process, img_h=args.img_h, img_w=args.img_w, patch_dim=args.patch_dim,
    )
    model.freeze(args.freeze_LM, args.freeze_ViT, False)
    return model


class MockVLMDataset(torch.utils.data.Dataset):
    """One image placeholder at the start of each caption; pixels and tokens are synthetic (``--mock-data``)."""

    def __init__(self, n, seq, vocab, h, w, seed):
        self.n, self.seq, self.vocab, self.h, self.w, self.seed = n, seq, vocab, h, w, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        toks = torch.randint(1, self.vocab, (self.seq + 1,), generator=g)
        tokens, labels = toks[:-1].clone(), toks[1:].clone()
        tokens[0] = DEFAULT_IMAGE_TOKEN_INDEX
        return {"tokens": tokens, "labels": labels, "loss_mask": torch.ones(self.seq), "position_ids": torch.arange(self.seq),
                "image": torch.randn(3, self.h, self.w, generator=g)}


def loss_func(output):
    losses, mask = output
    loss = torch.sum(losses.float() * mask) / mask.sum().clamp(min=1)
    return loss, {"lm loss": loss.detach()}


def forward_step(data_iterator, model):
    b = next(data_iterator)
    dev = next(model.parameters()).device
    b = {k: v.to(dev) for k, v in b.items()}
    out = model(b["image"].to(next(model.parameters()).dtype), b["tokens"], b["position_ids"], None, labels=b["labels"], loss_mask=b["loss_mask"].float())
    return out, loss_func


def train_valid_test_datasets_provider(num_samples):
    args = get_args()
    print_rank_0("> building synthetic VLM datasets ...")
    return tuple(MockVLMDataset(max(n, 1), args.seq_length, args.padded_vocab_size, args.img_h, args.img_w, args.seed) for n in num_samples)


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, forward_step, extra_args_provider=_extra_args, args_defaults={"tokenizer_type": "NullTokenizer"})
