#!/usr/bin/env python
"""Pretrain a GPT/Llama-family model (drop-in for the reference's ``pretrain_gpt.py``).

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 pretrain_gpt.py --model llama3_8b --tensor-model-parallel-size 8 \
        --sequence-parallel --bf16 --micro-batch-size 1 --global-batch-size 16 --train-iters 100 --lr 3e-4 --mock-data \
        --tokenizer-type NullTokenizer --vocab-size 128255 --use-distributed-optimizer --log-interval 10
"""
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from megatron_b200.core import parallel_state as ps  # noqa: E402
from megatron_b200.core.datasets import BlendedMegatronDatasetBuilder, GPTDatasetConfig, MockGPTDataset  # noqa: E402
from megatron_b200.core.datasets.gpt_dataset import GPTDataset  # noqa: E402
from megatron_b200.core.datasets.utils import get_blend_from_list  # noqa: E402
from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_decoder_block_spec, get_gpt_layer_local_spec  # noqa: E402
from megatron_b200.core.models.gpt.gpt_model import GPTModel  # noqa: E402
from megatron_b200.core.rerun_state_machine import get_rerun_state_machine  # noqa: E402
from megatron_b200.core.tokenizers import build_tokenizer  # noqa: E402
from megatron_b200.core.utils import get_batch_on_this_cp_rank  # noqa: E402
from megatron_b200.training.arguments import core_transformer_config_from_args  # noqa: E402
from megatron_b200.training.data import get_batch_on_this_tp_rank  # noqa: E402
from megatron_b200.training.training import get_args, pretrain, print_rank_0  # noqa: E402


def model_provider(pre_process=True, post_process=True, vp_stage=None) -> GPTModel:
    args = get_args()
    config = core_transformer_config_from_args(args)
    if args.num_experts:
        spec = get_gpt_decoder_block_spec(config, vp_stage=vp_stage)
    else:
        spec = get_gpt_layer_local_spec(normalization=args.normalization, qk_layernorm=args.qk_layernorm, multi_latent_attention=args.multi_latent_attention)
    return GPTModel(
        config=config, transformer_layer_spec=spec, vocab_size=args.padded_vocab_size, max_sequence_length=args.max_position_embeddings or args.seq_length,
        pre_process=pre_process, post_process=post_process, parallel_output=True, share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
        position_embedding_type=args.position_embedding_type, rotary_percent=args.rotary_percent, rotary_base=args.rotary_base,
        rope_scaling=args.use_rope_scaling, rope_scaling_factor=args.rope_scaling_factor, vp_stage=vp_stage,
    )


def get_batch(data_iterator):
    if not (ps.is_pipeline_first_stage(ignore_virtual=True) or ps.is_pipeline_last_stage(ignore_virtual=True)):
        return None, None, None, None
    b = get_batch_on_this_tp_rank(data_iterator)
    b = get_batch_on_this_cp_rank(b)
    return b["tokens"], b["labels"], b["loss_mask"], b["position_ids"]


def loss_func(loss_mask: torch.Tensor, output_tensor: torch.Tensor):
    args = get_args()
    losses = output_tensor.float().view(-1)
    mask = loss_mask.view(-1).float()
    total = torch.sum(losses * mask)
    ntok = mask.sum()
    rsm = get_rerun_state_machine()
    if args.check_for_nan_in_loss_and_grad or args.rerun_mode != "disabled":
        rsm.validate_result(total, rejection_func=lambda x: not bool(torch.isfinite(torch.as_tensor(x)).all()), message="found NaN/Inf in local forward loss", fatal=True)
    local_total, local_ntok = total, ntok
    if args.context_parallel_size > 1:
        t = torch.stack([total, ntok])
        torch.distributed.all_reduce(t, group=ps.get_context_parallel_group())
        total, ntok = t[0], t[1]
    if args.calculate_per_token_loss:
        # the schedule / finalize_model_grads sum num_tokens over dp x cp themselves: hand back the LOCAL sum and count
        # (reference pretrain_gpt.py loss_func); the CP-reduced mean is for logging only
        return local_total, local_ntok.detach().to(torch.int), {"lm loss": (total / ntok.clamp(min=1)).detach()}
    loss = total / ntok.clamp(min=1)
    return loss, {"lm loss": loss.detach()}


def forward_step(data_iterator, model: GPTModel):
    tokens, labels, loss_mask, position_ids = get_batch(data_iterator)
    out = model(tokens, position_ids, None, labels=labels)
    return out, partial(loss_func, loss_mask)


def train_valid_test_datasets_provider(train_val_test_num_samples):
    args = get_args()
    tokenizer = build_tokenizer(args.tokenizer_type, vocab_size=args.vocab_size, tokenizer_model=args.tokenizer_model)
    cfg = GPTDatasetConfig(
        random_seed=args.seed, sequence_length=args.seq_length, blend=None if args.mock_data else get_blend_from_list(args.data_path), split=args.split,
        path_to_cache=args.data_cache_path, tokenizer=tokenizer, reset_position_ids=args.reset_position_ids, reset_attention_mask=args.reset_attention_mask,
        eod_mask_loss=args.eod_mask_loss, create_attention_mask=False,
    )
    cls = MockGPTDataset if args.mock_data else GPTDataset
    print_rank_0("> building train, validation, and test datasets for GPT ...")
    is_built = lambda: ps.get_tensor_model_parallel_rank() == 0 and (ps.is_pipeline_first_stage(ignore_virtual=True) or ps.is_pipeline_last_stage(ignore_virtual=True))  # noqa: E731
    return BlendedMegatronDatasetBuilder(cls, train_val_test_num_samples, lambda: True, cfg).build()


if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, forward_step, args_defaults={"tokenizer_type": "NullTokenizer"})
