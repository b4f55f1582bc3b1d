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
from megatron_b200.core.tokenizers import build_tokenizer, build_tokenizer_from_args  # noqa: E402,F401
from megatron_b200.core.utils import get_batch_on_this_cp_rank  # noqa: E402
from megatron_b200.training.arguments import core_transformer_config_from_args  # noqa: E402
from megatron_b200.training.data import get_batch_on_this_tp_rank  # noqa: E402
from megatron_b200.training.training import get_args, pretrain, print_rank_0  # noqa: E402


def model_provider(pre_process=True, post_process=True, vp_stage=None) -> GPTModel:
    args = get_args()
    config = core_transformer_config_from_args(args)
    if getattr(args, "spec", None):
        # --spec module function: a user-supplied layer spec (reference pretrain_gpt.py model_provider: import_module(args.spec))
        import importlib

        mod, fn = args.spec[:2] if isinstance(args.spec, (list, tuple)) else str(args.spec).rsplit(".", 1)
        spec = getattr(importlib.import_module(mod), fn)
        spec = spec(config) if callable(spec) and not hasattr(spec, "submodules") else spec
    elif getattr(config, "experimental_attention_variant", None):
        # --experimental-attention-variant gdn|dsa: linear-attention / sparse-attention layers following --linear-attention-freq and the MoE pattern
        from megatron_b200.core.models.gpt.experimental_attention_variant_module_specs import get_transformer_block_with_experimental_attention_variant_spec

        spec = get_transformer_block_with_experimental_attention_variant_spec(config, vp_stage=vp_stage)
    elif args.num_experts:
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
    if getattr(args, "check_for_spiky_loss", False):
        # a loss far above anything seen in the first 100 observations: not fatal, but it goes through the rerun protocol (transient vs persistent fault)
        rsm.validate_result((total / ntok.clamp(min=1)).detach(), rejection_func=lambda x: rsm.is_unexpectedly_large(float(x), threshold=10.0, context="loss"),
                            message="Spiky loss", tolerance=0.0, fatal=False)
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


_EOD = {}


def _eod_token(args) -> int:
    """End-of-document id of the run's tokenizer (built once; the dataset provider builds its own instance on the data ranks only)."""
    if "id" not in _EOD:
        _EOD["id"] = build_tokenizer_from_args(args).eod
    return _EOD["id"]


def forward_step(data_iterator, model: GPTModel):
    args = get_args()
    tokens, labels, loss_mask, position_ids = get_batch(data_iterator)
    if getattr(args, "reset_attention_mask", False) and getattr(args, "reset_position_ids", False) and tokens is not None \
            and args.pipeline_model_parallel_size == 1 and args.context_parallel_size == 1:
        # documents must not attend to each other: flatten the micro-batch into ONE packed row and hand the boundaries to the attention kernels as cu_seqlens
        # (band mask inside the tcgen05 kernels, RoPE restarting per document) instead of materialising a dense [b, 1, s, s] mask
        from megatron_b200.core.packed_seq_params import packed_seq_params_from_documents

        b, s = tokens.shape
        psp = packed_seq_params_from_documents(tokens, _eod_token(args))
        out = model(tokens.reshape(1, b * s), position_ids.reshape(1, b * s), None, labels=labels.reshape(1, b * s), packed_seq_params=psp)
        return out.reshape(b, s), partial(loss_func, loss_mask)
    out = model(tokens, position_ids, None, labels=labels)
    return out, partial(loss_func, loss_mask)


def train_valid_test_datasets_provider(train_val_test_num_samples):
    args = get_args()
    tokenizer = build_tokenizer_from_args(args)
    per_split = [getattr(args, k, None) for k in ("train_data_path", "valid_data_path", "test_data_path")]
    use_per_split = not args.mock_data and any(per_split)            # --train-data-path / --valid-data-path / --test-data-path instead of --data-path + --split
    cfg = GPTDatasetConfig(
        random_seed=args.seed, sequence_length=args.seq_length,
        blend=None if (args.mock_data or use_per_split) else get_blend_from_list(args.data_path), split=None if use_per_split else args.split,
        blend_per_split=[get_blend_from_list(p) if p else None for p in per_split] if use_per_split else None,
        path_to_cache=args.data_cache_path, tokenizer=tokenizer, reset_position_ids=args.reset_position_ids, reset_attention_mask=args.reset_attention_mask,
        eod_mask_loss=args.eod_mask_loss, create_attention_mask=False, mmap_bin_files=getattr(args, "mmap_bin_files", True),
        num_dataset_builder_threads=getattr(args, "num_dataset_builder_threads", 1) or 1,
    )
    cls = MockGPTDataset if args.mock_data else GPTDataset
    print_rank_0("> building train, validation, and test datasets for GPT ...")
    is_built = lambda: ps.get_tensor_model_parallel_rank() == 0 and (ps.is_pipeline_first_stage(ignore_virtual=True) or ps.is_pipeline_last_stage(ignore_virtual=True))  # noqa: E731
    out = BlendedMegatronDatasetBuilder(cls, train_val_test_num_samples, lambda: True, cfg).build()
    if getattr(args, "fim_data", False):
        # fill-in-the-middle: documents are re-ordered (prefix / suffix / middle) with sentinel tokens (reference pretrain_gpt.py --fim-data, GPTFIMDataset)
        from megatron_b200.training.datasets.fim_dataset import FIMConfig, GPTFIMDataset

        def tok_id(flag, default):
            v = getattr(args, flag, None)
            if v is None:
                return default
            if isinstance(v, int) or str(v).lstrip("-").isdigit():
                return int(v)
            try:
                ids = tokenizer.tokenize(v) if hasattr(tokenizer, "tokenize") else []
            except (ValueError, KeyError):                      # the tokenizer has no such special token (e.g. NullTokenizer): reserve ids at the top of the vocabulary
                ids = []
            return int(ids[0]) if len(ids) == 1 else default

        v = args.padded_vocab_size if getattr(args, "padded_vocab_size", None) else args.vocab_size
        fim = FIMConfig(fim_rate=getattr(args, "fim_rate", 0.5), fim_spm_rate=getattr(args, "fim_spm_rate", 0.5), prefix_id=tok_id("fim_prefix_token", v - 5),
                        middle_id=tok_id("fim_middle_token", v - 4), suffix_id=tok_id("fim_suffix_token", v - 3), pad_id=tok_id("fim_pad_token", v - 2),
                        eod_id=tok_id("fim_eod_token", getattr(tokenizer, "eod", v - 1)))
        out = tuple(GPTFIMDataset(d, fim, seed=args.seed) if i == 0 and d is not None else d for i, d in enumerate(out))
    return out


if __name__ == "__main__":
    run = pretrain
    if "--inprocess-restart" in sys.argv:
        # recoverable failures restart the training function inside this process (reference pretrain_gpt.py: inprocess_restart.maybe_wrap_for_inprocess_restart)
        from megatron_b200.training.inprocess_restart import maybe_wrap_for_inprocess_restart

        budget = int(sys.argv[sys.argv.index("--inprocess-max-iterations") + 1]) if "--inprocess-max-iterations" in sys.argv else 5
        run = maybe_wrap_for_inprocess_restart(pretrain, max_restarts=budget)
    run(train_valid_test_datasets_provider, model_provider, forward_step, args_defaults={"tokenizer_type": "NullTokenizer"})
