"""Reference-compatible command line: every flag of the reference's ``megatron/training`` parsers that this framework does not define itself is accepted with the
reference's name, kind, type and default (``reference_flags_table.py``, generated), so an existing launch script parses unchanged.

Three classes of flags, decided HERE (nothing is silently ignored):

* **wired** – ``apply_reference_compat`` translates the value into the option / config field this framework uses (``WIRED`` lists the target);
* **native** – the dest has the same name as a field of a config dataclass (``TransformerConfig`` / ``MLATransformerConfig`` / optimizer / DDP) or is read by
  ``getattr(args, name)`` somewhere in ``training/`` — it flows through by name;
* **inert** – accepted for script compatibility, no effect in this build; ``inert_flags_in_use`` reports the ones a run actually set, and ``parse_args`` prints
  them on rank 0 (``--strict-reference-flags`` turns that report into an error)."""
from __future__ import annotations

import argparse
from typing import Dict, List

from .reference_flags_table import REFERENCE_FLAG_TABLE

_TYPES = {"int": int, "float": float, "str": str, "bool": lambda s: str(s).lower() in ("1", "true", "yes"), None: None}


def add_reference_compat_flags(parser: argparse.ArgumentParser) -> List[str]:
    """Register the table's flags that the parser does not know yet.  Returns the dests added."""
    known = {s for a in parser._actions for s in a.option_strings}
    dests = {a.dest for a in parser._actions}
    group = parser.add_argument_group("reference-compatible flags (generated table)")
    added = []
    for flags, kind, typ, default, extra, _src in REFERENCE_FLAG_TABLE:
        flags = tuple(f for f in flags if f not in known)
        if not flags:
            continue
        kw = dict(extra)
        dest = kw.get("dest") or flags[0].lstrip("-").replace("-", "_")
        if kind == "store_false" and dest.startswith("no_") and "dest" not in kw:
            dest = flags[0].lstrip("-").replace("-", "_")
        if dest in dests:
            kw["dest"] = dest                                   # alias of an option that already exists under another spelling
            default = argparse.SUPPRESS
        if kind == "store_true":
            group.add_argument(*flags, action="store_true", **{k: v for k, v in kw.items() if k == "dest"}, default=argparse.SUPPRESS if default is argparse.SUPPRESS else (bool(default) if default is not None else False))
        elif kind == "store_false":
            group.add_argument(*flags, action="store_false", **{k: v for k, v in kw.items() if k == "dest"}, default=argparse.SUPPRESS if default is argparse.SUPPRESS else (True if default is None else default))
        elif kind == "bool_opt":
            group.add_argument(*flags, action=argparse.BooleanOptionalAction, **{k: v for k, v in kw.items() if k == "dest"}, default=default)
        else:
            opts = {k: v for k, v in kw.items() if k in ("dest", "nargs", "choices", "const")}
            if _TYPES.get(typ) is not None:
                opts["type"] = _TYPES[typ]
            opts["default"] = default
            group.add_argument(*flags, **opts)
        known.update(flags)
        dests.add(dest)
        added.append(dest)
    # the config dataclass default of this switch is False while the reference's COMMAND LINE default is on (--no-rope-fusion turns it off): follow the command line
    parser.set_defaults(apply_rope_fusion=True)
    if "--strict-reference-flags" not in known:
        group.add_argument("--strict-reference-flags", action="store_true", help="fail when a reference flag that has no effect in this build is set")
    return added


# dest of a table flag -> what consumes it here.  ``tests/test_training_utils_cpu.py::test_reference_flag_table`` checks that every name listed is really read
# (by ``apply_reference_compat`` / ``engine_kwargs_from_args`` below or somewhere else in the tree), so this list cannot drift into wishful thinking.
WIRED: Dict[str, str] = {
    "verify_integrity": "SHA-256 checkpoint manifest written on save / checked on load (training.checkpointing.configure)",
    # spellings of options that exist under another name
    "tp_size": "tensor_model_parallel_size", "model_parallel_size": "tensor_model_parallel_size (legacy)", "ep_size": "expert_model_parallel_size", "batch_size": "micro_batch_size (legacy)",
    "warmup": "lr_warmup_fraction (legacy)", "checkpoint_activations": "recompute_granularity = full (deprecated spelling)", "grad_reduce_in_bf16": "accumulate_allreduce_grads_in_fp32 = False",
    "no_position_embedding": "position_embedding_type = none", "encoder_seq_length": "seq_length", "encoder_num_layers": "num_layers",
    "decoder_first_pipeline_num_layers": "num_layers_in_first_pipeline_stage", "decoder_last_pipeline_num_layers": "num_layers_in_last_pipeline_stage",
    "hybrid_layer_pattern": "hybrid_override_pattern", "muon_num_ns_steps": "muon_ns_steps", "no_one_logger": "enable_one_logger = False", "train_data_path": "per-split blend (pretrain_gpt.py)",
    "valid_data_path": "per-split blend", "test_data_path": "per-split blend", "disable_symmetric_registration": "MEGATRON_B200_DP_COMM=nccl (no symmetric-heap DP collectives)",
    "yarn_beta_fast": "MLATransformerConfig.beta_fast", "yarn_beta_slow": "MLATransformerConfig.beta_slow", "yarn_original_max_position_embeddings": "MLATransformerConfig.original_max_position_embeddings",
    "ddp_average_in_collective": "DistributedDataParallelConfig.average_in_collective", "ddp_pad_buckets_for_high_nccl_busbw": "DistributedDataParallelConfig.pad_buckets_for_high_nccl_busbw",
    "ddp_reduce_scatter_with_fp32_accumulation": "DistributedDataParallelConfig.reduce_scatter_with_fp32_accumulation", "use_nccl_ub": "DistributedDataParallelConfig.nccl_ub",
    "ddp_num_buckets": "DDP bucket_size = ceil(parameters / n) (training.py)",
    # read directly
    "apply_rope_fusion": "TransformerConfig.apply_rope_fusion (--no-rope-fusion)", "bias_gelu_fusion": "TransformerConfig.bias_activation_fusion (--no-bias-gelu-fusion)",
    "bias_swiglu_fusion": "TransformerConfig.bias_activation_fusion (--no-bias-swiglu-fusion)", "openai_gelu": "tanh-approximated GeLU", "quick_geglu": "quick-GeGLU activation",
    "init_method_xavier_uniform": "xavier-uniform init", "spec": "user layer spec (module, function) in pretrain_gpt.py", "mmap_bin_files": "GPTDatasetConfig.mmap_bin_files (--no-mmap-bin-files)",
    "num_dataset_builder_threads": "GPTDatasetConfig.num_dataset_builder_threads", "fim_data": "GPTFIMDataset around the train split", "fim_rate": "FIMConfig.fim_rate", "fim_spm_rate": "FIMConfig.fim_spm_rate",
    "fim_prefix_token": "FIMConfig.prefix_id", "fim_middle_token": "FIMConfig.middle_id", "fim_suffix_token": "FIMConfig.suffix_id", "fim_pad_token": "FIMConfig.pad_id", "fim_eod_token": "FIMConfig.eod_id",
    "ft_num_warmup_iters": "FaultToleranceMonitor.min_samples", "te_precision_config_file": "TransformerConfig.quant_recipe (per-layer precision YAML)",
    "kitchen_config_file": "TransformerConfig.quant_recipe (per-layer precision YAML)",
    # training loop (training/training.py)
    "step_batch_size_schedule": "StepBatchsizeNumMicroBatchesCalculator (thresholds in tokens)", "decrease_batch_size_if_needed": "micro-batch calculator rounds the batch down",
    "iterations_to_skip": "train loop fast-forwards the data iterator", "train_sync_interval": "device synchronize every N iterations", "skip_train": "pretrain() goes straight to evaluation",
    "start_eval_at_iter": "first in-training evaluation", "manual_gc_eval": "gc.collect around evaluation (--no-manual-gc-eval)", "empty_unused_memory_level": "empty_cache after fwd/bwd (1) and after the optimizer (2)",
    "exit_signal": "signal that requests checkpoint-and-exit", "lr_wsd_decay_samples": "OptimizerParamScheduler.wsd_decay_steps (sample-based runs)",
    "override_opt_param_scheduler": "OptimizerParamScheduler.override_opt_param_scheduler", "use_checkpoint_opt_param_scheduler": "OptimizerParamScheduler.use_checkpoint_opt_param_scheduler",
    "error_injection_type": "RerunErrorInjector kind", "check_for_spiky_loss": "loss_func spike validation through the rerun state machine (pretrain_gpt.py)",
    "use_tp_pp_dp_mapping": "initialize_model_parallel(order='tp-cp-ep-pp-dp')", "nccl_communicator_config_path": "initialize_model_parallel(nccl_communicator_config_path=)",
    "high_priority_stream_groups": "initialize_model_parallel(high_priority_stream_groups=)",
    "data_parallel_random_init": "seed + 10 * dp_rank, parameters broadcast from the first replica", "batch_invariant_mode": "enable_batch_invariant_mode()", "logging_level": "root logger level",
    "profile_ranks": "ranks that start the profiler", "record_shapes": "torch profiler record_shapes",
    "tensorboard_log_interval": "writer cadence", "log_timers_to_tensorboard": "Timers.write", "log_loss_scale_to_tensorboard": "loss-scale scalar (--no-log-loss-scale-to-tensorboard)",
    "log_validation_ppl_to_tensorboard": "validation ppl scalars", "log_memory_to_tensorboard": "allocator scalars", "log_memory_interval": "allocator scalar cadence",
    "log_world_size_to_tensorboard": "world-size scalar", "wandb_entity": "wandb.init(entity=)",
    "fault_injector_ranks": "FaultInjectorConfig.from_args", "fault_injector_num_ranks": "FaultInjectorConfig.from_args", "fault_injector_fault_types": "FaultInjectorConfig.from_args",
    "fault_injector_fault_probabilities": "FaultInjectorConfig.from_args", "fault_injector_fault_delay": "FaultInjectorConfig.from_args",
    "fault_injector_delay_start_iteration": "FaultInjectorConfig.from_args", "fault_injector_mtti_seconds": "FaultInjectorConfig.from_args",
    "fault_injector_offset_seconds": "FaultInjectorConfig.from_args", "fault_injector_seed": "FaultInjectorConfig.from_args",
    "replication": "local checkpoints keep copies of other ranks' blobs", "replication_jump": "spacing between a rank and the ranks holding its replicas",
    "replication_factor": "number of holders of each blob",
    "use_mp_args_from_checkpoint_args": "load_args_from_checkpoint(model_parallel=True)",
    "enable_experimental": "core.config.set_experimental_flag(True)",
    "save_retain_interval": "save_checkpoint(retain_interval=): the previous checkpoint is deleted unless its iteration is a multiple",
    # tokenizer / vocabulary
    "padded_vocab_size": "pins the padded vocabulary", "pad_vocab_size": "--no-pad-vocab-size: no rounding", "vocab_extra_ids": "added before padding",
    "null_tokenizer_eod_id": "NullTokenizer eod", "null_tokenizer_pad_id": "NullTokenizer pad", "tiktoken_pattern": "TikTokenTokenizer splitter (v1 / v2 / regex)",
    "tiktoken_num_special_tokens": "TikTokenTokenizer special slots", "tokenizer_special_tokens": "tokenizer special tokens", "tokenizer_hf_no_use_fast": "AutoTokenizer use_fast=False",
    "tokenizer_hf_no_include_special_tokens": "HF encode without special tokens", "trust_remote_code": "AutoTokenizer trust_remote_code", "chat_template": "tokenizer chat template",
    "tokenizer_metadata": "MegatronTokenizer.from_pretrained(metadata)", "tokenizer_sentencepiece_legacy": "accepted by the SentencePiece wrapper",
    # serving (tools/run_text_generation_server.py through engine_kwargs_from_args)
    "inference_dynamic_batching_block_size": "engine block_size", "inference_dynamic_batching_max_requests": "engine max_running", "inference_max_requests": "engine max_running",
    "inference_dynamic_batching_max_tokens": "engine max_prefill_tokens_per_step", "enable_chunked_prefill": "engine max_prefill_tokens_per_step (2048 when no budget is given)",
    "inference_dynamic_batching_enable_prefix_caching": "engine enable_prefix_caching", "inference_dynamic_batching_num_cuda_graphs": "engine decode buckets + CUDA graphs",
    "decode_only_cuda_graphs": "engine enable_cuda_graphs",
}

# accepted and already the behaviour of this build (setting them changes nothing, and that is correct)
ALWAYS_ON: Dict[str, str] = {
    "use_dist_ckpt": "checkpoints are always distributed (torch_dist)", "dist_ckpt_format": "torch_dist is the only format", "calc_ft_timeouts": "section timeouts are always learned and persisted",
    "use_mcore_models": "there is no legacy model path", "inference_dynamic_batching": "the dynamic engine is the default engine", "perform_rl_step": "train_rl.py always performs RL steps",
    "inprocess_restart": "handled by pretrain_gpt.py before argument parsing (sys.argv)", "inprocess_max_iterations": "handled by pretrain_gpt.py before argument parsing (sys.argv)",
    "ckpt_fully_parallel_save": "fully-parallel save is the default; --no-ckpt-fully-parallel-save is read by name",
    "auto_detect_ckpt_format": "the loader always detects the format from the checkpoint metadata", "use_dist_ckpt_deprecated": "checkpoints are always distributed",
    "dist_ckpt_format_deprecated": "torch_dist is the only format", "deprecated_use_mcore_models": "there is no legacy model path", "local_rank": "LOCAL_RANK comes from the launcher environment",
    "lazy_mpu_init": "model-parallel state is always initialised by initialize_megatron", "disable_jit_fuser": "there is no JIT fuser: fusions are CUDA kernels",
"ckpt_load_validate_sharding_integrity": "sharding integrity is always validated on load",
}


def apply_reference_compat(args) -> None:
    """Translate reference spellings into this framework's options (only when the reference flag was actually given)."""
    g = lambda n, d=None: getattr(args, n, d)  # noqa: E731

    if g("tp_size") and g("tensor_model_parallel_size", 1) == 1:
        args.tensor_model_parallel_size = g("tp_size")
    if g("model_parallel_size") and g("tensor_model_parallel_size", 1) == 1:
        args.tensor_model_parallel_size = g("model_parallel_size")
    if g("ep_size") and g("expert_model_parallel_size", 1) == 1:
        args.expert_model_parallel_size = g("ep_size")
    if g("batch_size") and not g("micro_batch_size"):
        args.micro_batch_size = g("batch_size")
    if g("warmup") is not None and g("lr_warmup_fraction") is None:
        args.lr_warmup_fraction = g("warmup")
    if g("checkpoint_activations") and not g("recompute_granularity"):
        args.recompute_granularity, args.recompute_method = "full", g("recompute_method") or "uniform"
        if not g("recompute_num_layers"):
            args.recompute_num_layers = 1
    if g("grad_reduce_in_bf16"):
        args.accumulate_allreduce_grads_in_fp32 = False
    if g("no_position_embedding"):
        args.position_embedding_type = "none"
    if g("encoder_seq_length") and not g("seq_length"):
        args.seq_length = g("encoder_seq_length")
    if g("encoder_num_layers") and not g("num_layers"):
        args.num_layers = g("encoder_num_layers")
    if g("decoder_first_pipeline_num_layers") is not None:
        args.num_layers_in_first_pipeline_stage = g("decoder_first_pipeline_num_layers")
    if g("decoder_last_pipeline_num_layers") is not None:
        args.num_layers_in_last_pipeline_stage = g("decoder_last_pipeline_num_layers")
    if g("hybrid_layer_pattern") and not g("hybrid_override_pattern"):
        args.hybrid_override_pattern = g("hybrid_layer_pattern")
    if g("muon_num_ns_steps") is not None:
        args.muon_ns_steps = g("muon_num_ns_steps")
    for ref, ours in (("yarn_beta_fast", "beta_fast"), ("yarn_beta_slow", "beta_slow"), ("yarn_original_max_position_embeddings", "original_max_position_embeddings")):
        if g(ref) is not None:
            setattr(args, ours, g(ref))
    for ref, ours in (("ddp_average_in_collective", "average_in_collective"), ("ddp_pad_buckets_for_high_nccl_busbw", "pad_buckets_for_high_nccl_busbw"),
                      ("ddp_reduce_scatter_with_fp32_accumulation", "reduce_scatter_with_fp32_accumulation"), ("use_nccl_ub", "nccl_ub")):
        if g(ref):
            setattr(args, ours, True)
    if g("disable_symmetric_registration"):
        import os

        os.environ["MEGATRON_B200_DP_COMM"] = "nccl"
    if g("no_one_logger"):
        args.enable_one_logger = False
    if g("enable_chunked_prefill") and not g("inference_dynamic_batching_max_tokens"):
        args.inference_dynamic_batching_max_tokens = 2048
    if g("inference_max_requests") and not g("inference_dynamic_batching_max_requests"):
        args.inference_dynamic_batching_max_requests = g("inference_max_requests")
    if g("train_data_path") and not g("data_path"):
        args.data_path = list(g("train_data_path"))
    if g("ckpt_fully_parallel_save") is False:
        args.ckpt_fully_parallel_save = False


def engine_kwargs_from_args(args) -> dict:
    """``DynamicInferenceEngine`` keyword arguments from the reference's ``--inference-dynamic-batching-*`` flags."""
    g = lambda n, d=None: getattr(args, n, d)  # noqa: E731
    kw = {}
    if g("inference_dynamic_batching_block_size"):
        kw["block_size"] = g("inference_dynamic_batching_block_size")
    if g("inference_dynamic_batching_max_requests"):
        kw["max_running"] = g("inference_dynamic_batching_max_requests")
    if g("inference_dynamic_batching_max_tokens"):
        kw["max_prefill_tokens_per_step"] = g("inference_dynamic_batching_max_tokens")
    if g("inference_dynamic_batching_enable_prefix_caching"):
        kw["enable_prefix_caching"] = True
    n_graphs = g("inference_dynamic_batching_num_cuda_graphs")
    if n_graphs or g("decode_only_cuda_graphs"):
        kw["enable_cuda_graphs"] = True
        if n_graphs and kw.get("max_running"):
            mr = kw["max_running"]
            kw["decode_batch_buckets"] = sorted({max(1, round(mr * (i + 1) / n_graphs)) for i in range(n_graphs)})
    return kw


def inert_flags_in_use(args, parser: argparse.ArgumentParser) -> List[str]:
    """Reference flags a run set to a non-default value that nothing in this build consumes."""
    import dataclasses

    from .arguments import config_classes

    native = set()
    for cls in config_classes():
        native |= {f.name for f in dataclasses.fields(cls)}
    try:
        from ..core.transformer.transformer_config import MLATransformerConfig

        native |= {f.name for f in dataclasses.fields(MLATransformerConfig)}
    except ImportError:
        pass
    out = []
    for grp in parser._action_groups:
        if grp.title != "reference-compatible flags (generated table)":
            continue
        for a in grp._group_actions:
            if a.default is argparse.SUPPRESS or a.dest == "strict_reference_flags":
                continue                                                  # another spelling of an option this framework defines itself
            if a.dest in WIRED or a.dest in ALWAYS_ON or a.dest in native:
                continue
            if getattr(args, a.dest, a.default) != a.default:
                out.append(a.option_strings[0])
    return sorted(out)
