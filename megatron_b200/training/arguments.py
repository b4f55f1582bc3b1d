"""Command-line interface (reference ``megatron/training/arguments.py``: ~1000 flags in ~30 groups).

The flag names follow the reference so launch scripts carry over; flags that selected
TransformerEngine/Apex code paths are accepted and ignored (there is one native back end).
``core_transformer_config_from_args`` maps the namespace onto ``TransformerConfig``.
"""
from __future__ import annotations

import argparse
import os
from typing import Optional

import torch
import torch.nn.functional as F


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="megatron_b200 arguments", allow_abbrev=False)
    g = p.add_argument_group("network size")
    g.add_argument("--num-layers", type=int, default=None)
    g.add_argument("--hidden-size", type=int, default=None)
    g.add_argument("--ffn-hidden-size", type=int, default=None)
    g.add_argument("--num-attention-heads", type=int, default=None)
    g.add_argument("--kv-channels", type=int, default=None)
    g.add_argument("--group-query-attention", action="store_true")
    g.add_argument("--num-query-groups", type=int, default=None)
    g.add_argument("--max-position-embeddings", type=int, default=None)
    g.add_argument("--position-embedding-type", default="learned_absolute", choices=["learned_absolute", "rope", "yarn", "none"])
    g.add_argument("--use-rotary-position-embeddings", action="store_true")
    g.add_argument("--rotary-base", type=int, default=10000)
    g.add_argument("--rotary-percent", type=float, default=1.0)
    g.add_argument("--use-rope-scaling", action="store_true")
    g.add_argument("--rope-scaling-factor", type=float, default=8.0)
    g.add_argument("--normalization", default="LayerNorm", choices=["LayerNorm", "RMSNorm"])
    g.add_argument("--norm-epsilon", type=float, default=1e-5)
    g.add_argument("--swiglu", action="store_true")
    g.add_argument("--squared-relu", action="store_true")
    g.add_argument("--disable-bias-linear", action="store_false", dest="add_bias_linear")
    g.add_argument("--add-qkv-bias", action="store_true")
    g.add_argument("--untie-embeddings-and-output-weights", action="store_true")
    g.add_argument("--make-vocab-size-divisible-by", type=int, default=128)
    g.add_argument("--qk-layernorm", action="store_true")
    g.add_argument("--multi-latent-attention", action="store_true")
    g.add_argument("--model", default=None, help="named preset from megatron_b200.models.presets (overrides the size flags)")

    g = p.add_argument_group("regularization / init")
    g.add_argument("--attention-dropout", type=float, default=0.1)
    g.add_argument("--hidden-dropout", type=float, default=0.1)
    g.add_argument("--weight-decay", type=float, default=0.01)
    g.add_argument("--clip-grad", type=float, default=1.0)
    g.add_argument("--adam-beta1", type=float, default=0.9)
    g.add_argument("--adam-beta2", type=float, default=0.999)
    g.add_argument("--adam-eps", type=float, default=1e-8)
    g.add_argument("--init-method-std", type=float, default=0.02)
    g.add_argument("--seed", type=int, default=1234)

    g = p.add_argument_group("training")
    g.add_argument("--micro-batch-size", type=int, default=1)
    g.add_argument("--global-batch-size", type=int, default=None)
    g.add_argument("--rampup-batch-size", nargs="*", default=None)
    g.add_argument("--seq-length", type=int, default=None)
    g.add_argument("--train-iters", type=int, default=None)
    g.add_argument("--train-samples", type=int, default=None)
    g.add_argument("--exit-interval", type=int, default=None)
    g.add_argument("--exit-duration-in-mins", type=int, default=None)
    g.add_argument("--optimizer", default="adam", choices=["adam", "sgd", "lion", "muon", "soap"])
    g.add_argument("--soap-shampoo-beta", type=float, default=0.95)
    g.add_argument("--soap-precondition-frequency", type=int, default=10)
    g.add_argument("--soap-max-precond-dim", type=int, default=8192)
    g.add_argument("--muon-momentum", type=float, default=0.95)
    g.add_argument("--muon-ns-steps", type=int, default=5)
    g.add_argument("--muon-tp-mode", default="blockwise", choices=["blockwise", "duplicated"])
    g.add_argument("--recompute-granularity", default=None, choices=["full", "selective"])
    g.add_argument("--recompute-method", default=None, choices=["uniform", "block"])
    g.add_argument("--recompute-num-layers", type=int, default=None)
    g.add_argument("--recompute-modules", nargs="*", default=None)
    g.add_argument("--recompute-activations", action="store_true")
    g.add_argument("--no-gradient-accumulation-fusion", action="store_false", dest="gradient_accumulation_fusion")
    g.add_argument("--use-flash-attn", action="store_true")
    g.add_argument("--attention-backend", default="auto")
    g.add_argument("--transformer-impl", default="b200", choices=["b200", "local", "transformer_engine"])
    g.add_argument("--deterministic-mode", action="store_true")
    g.add_argument("--check-for-nan-in-loss-and-grad", action="store_true")
    g.add_argument("--manual-gc", action="store_true")
    g.add_argument("--manual-gc-interval", type=int, default=0)
    g.add_argument("--cross-entropy-loss-fusion", action="store_true")
    g.add_argument("--calculate-per-token-loss", action="store_true")

    g = p.add_argument_group("learning rate")
    g.add_argument("--lr", type=float, default=None)
    g.add_argument("--min-lr", type=float, default=0.0)
    g.add_argument("--lr-decay-style", default="linear", choices=["constant", "linear", "cosine", "inverse-square-root", "WSD"])
    g.add_argument("--lr-decay-iters", type=int, default=None)
    g.add_argument("--lr-decay-samples", type=int, default=None)
    g.add_argument("--lr-warmup-fraction", type=float, default=None)
    g.add_argument("--lr-warmup-iters", type=int, default=0)
    g.add_argument("--lr-warmup-samples", type=int, default=0)
    g.add_argument("--lr-warmup-init", type=float, default=0.0)
    g.add_argument("--lr-wsd-decay-style", default="exponential")
    g.add_argument("--lr-wsd-decay-iters", type=int, default=None)
    g.add_argument("--start-weight-decay", type=float, default=None)
    g.add_argument("--end-weight-decay", type=float, default=None)
    g.add_argument("--weight-decay-incr-style", default="constant", choices=["constant", "linear", "cosine"])

    g = p.add_argument_group("checkpointing")
    g.add_argument("--save", default=None)
    g.add_argument("--load", default=None)
    g.add_argument("--save-interval", type=int, default=None)
    g.add_argument("--no-save-optim", action="store_true")
    g.add_argument("--no-load-optim", action="store_true")
    g.add_argument("--no-load-rng", action="store_true")
    g.add_argument("--finetune", action="store_true")
    g.add_argument("--ckpt-format", default="torch_dist", choices=["torch_dist"])
    g.add_argument("--async-save", action="store_true")
    g.add_argument("--ckpt-fully-parallel-save", action="store_true", default=True)
    g.add_argument("--dist-ckpt-optim-fully-reshardable", action="store_true", default=True)
    g.add_argument("--keep-last-checkpoints", type=int, default=None)
    g.add_argument("--ckpt-assume-constant-structure", action="store_true", help="reuse the previous save's plan and metadata when the checkpoint structure did not change")
    g.add_argument("--ckpt-fully-parallel-load", action="store_true", help="each DP-replicated shard is read from storage once per dp-cp group and exchanged")
    g.add_argument("--dist-ckpt-strictness", default="assume_ok_unexpected",
                   choices=["assume_ok_unexpected", "log_unexpected", "log_all", "raise_unexpected", "raise_all", "return_unexpected", "return_all", "ignore_all"])
    g.add_argument("--ckpt-step", type=int, default=None, help="load this iteration instead of the newest one")
    g.add_argument("--exit-on-missing-checkpoint", action="store_true", help="with --load set and nothing to load: exit instead of training from scratch")
    g.add_argument("--pretrained-checkpoint", default=None, help="weights to start from when --load holds no checkpoint (finetuning)")
    g.add_argument("--no-save-rng", action="store_true")
    g.add_argument("--non-persistent-ckpt-type", default=None, choices=["global", "local", "in_memory"], help="kind of the frequent recovery checkpoint")
    g.add_argument("--non-persistent-global-ckpt-dir", default=None)
    g.add_argument("--non-persistent-local-ckpt-algo", default="fully_parallel", choices=["fully_parallel", "atomic"])
    g.add_argument("--use-persistent-ckpt-worker", action="store_true", default=True, help="async saves write from one long-lived worker process (default)")

    g = p.add_argument_group("mixed precision")
    g.add_argument("--fp16", action="store_true")
    g.add_argument("--bf16", action="store_true")
    g.add_argument("--no-bf16-tiny", dest="bf16", action="store_false", help="examples' TINY (CPU smoke) mode: undo an earlier --bf16")
    g.add_argument("--loss-scale", type=float, default=None)
    g.add_argument("--initial-loss-scale", type=float, default=2**32)
    g.add_argument("--min-loss-scale", type=float, default=1.0)
    g.add_argument("--loss-scale-window", type=float, default=1000)
    g.add_argument("--hysteresis", type=int, default=2)
    g.add_argument("--accumulate-allreduce-grads-in-fp32", action="store_true")
    g.add_argument("--fp8-format", default=None, choices=["e4m3", "hybrid"])
    g.add_argument("--fp8-recipe", default="mxfp8")

    g = p.add_argument_group("distributed")
    g.add_argument("--tensor-model-parallel-size", type=int, default=1)
    g.add_argument("--pipeline-model-parallel-size", type=int, default=1)
    g.add_argument("--num-layers-per-virtual-pipeline-stage", type=int, default=None)
    g.add_argument("--num-virtual-stages-per-pipeline-rank", type=int, default=None)
    g.add_argument("--context-parallel-size", type=int, default=1)
    g.add_argument("--cp-comm-type", default="p2p")
    g.add_argument("--expert-model-parallel-size", type=int, default=1)
    g.add_argument("--expert-tensor-parallel-size", type=int, default=None)
    g.add_argument("--sequence-parallel", action="store_true")
    g.add_argument("--use-distributed-optimizer", action="store_true")
    g.add_argument("--overlap-grad-reduce", action="store_true")
    g.add_argument("--overlap-param-gather", action="store_true")
    g.add_argument("--overlap-p2p-communication", action="store_true", dest="overlap_p2p_comm")
    g.add_argument("--tp-comm-overlap", action="store_true", help="fused in-kernel NVLink AG→GEMM / GEMM→RS")
    g.add_argument("--tp-comm", default="auto", choices=["auto", "nccl", "nvlink", "fused"])
    g.add_argument("--distributed-backend", default="nccl", choices=["nccl", "gloo"])
    g.add_argument("--distributed-timeout-minutes", type=int, default=10)
    g.add_argument("--ddp-bucket-size", type=int, default=None)
    g.add_argument("--use-cpu-initialization", action="store_true")

    g = p.add_argument_group("moe")
    g.add_argument("--num-experts", type=int, default=None)
    g.add_argument("--moe-router-topk", type=int, default=2)
    g.add_argument("--moe-router-load-balancing-type", default="aux_loss")
    g.add_argument("--moe-aux-loss-coeff", type=float, default=0.0)
    g.add_argument("--moe-z-loss-coeff", type=float, default=None)
    g.add_argument("--moe-token-dispatcher-type", default="alltoall", choices=["allgather", "alltoall", "flex"])
    g.add_argument("--moe-grouped-gemm", action="store_true")
    g.add_argument("--moe-ffn-hidden-size", type=int, default=None)
    g.add_argument("--moe-shared-expert-intermediate-size", type=int, default=None)
    g.add_argument("--moe-expert-capacity-factor", type=float, default=None)
    g.add_argument("--moe-layer-freq", type=int, default=1)

    g = p.add_argument_group("data")
    g.add_argument("--data-path", nargs="*", default=None)
    g.add_argument("--split", default="969,30,1")
    g.add_argument("--mock-data", action="store_true")
    g.add_argument("--tokenizer-type", default="NullTokenizer")
    g.add_argument("--tokenizer-model", default=None)
    g.add_argument("--vocab-size", type=int, default=None)
    g.add_argument("--vocab-file", default=None)
    g.add_argument("--merge-file", default=None)
    g.add_argument("--num-workers", type=int, default=2)
    g.add_argument("--data-cache-path", default=None)
    g.add_argument("--reset-position-ids", action="store_true")
    g.add_argument("--reset-attention-mask", action="store_true")
    g.add_argument("--eod-mask-loss", action="store_true")

    g = p.add_argument_group("logging / validation")
    g.add_argument("--log-interval", type=int, default=100)
    g.add_argument("--log-throughput", action="store_true")
    g.add_argument("--timing-log-level", type=int, default=0, choices=[0, 1, 2])
    g.add_argument("--timing-log-option", default="minmax", choices=["max", "minmax", "all"])
    g.add_argument("--tensorboard-dir", default=None)
    g.add_argument("--wandb-project", default=None)
    g.add_argument("--wandb-exp-name", default=None)
    g.add_argument("--wandb-save-dir", default=None)
    g.add_argument("--tensorboard-queue-size", type=int, default=1000)
    g.add_argument("--enable-one-logger", action="store_true")
    g.add_argument("--enable-ft-package", action="store_true", help="section-timeout hang detection (training/ft_integration.py)")
    g.add_argument("--ft-timeout-step", type=float, default=None)
    g.add_argument("--ft-timeout-setup", type=float, default=None)
    g.add_argument("--ft-timeout-checkpointing", type=float, default=None)
    g.add_argument("--simulated-fault", default=None, help="kind:delay_s[:rank], kind in {rank_killed, rank_hung}")
    g.add_argument("--log-activations-interval", type=int, default=0)
    g.add_argument("--log-dgrad-interval", type=int, default=0)
    g.add_argument("--log-wgrad-interval", type=int, default=0)
    g.add_argument("--activation-log-dir", default=None)
    g.add_argument("--exit-signal-handler", action="store_true")
    g.add_argument("--log-params-norm", action="store_true")
    g.add_argument("--use-pytorch-profiler", action="store_true", help="torch.profiler chrome trace for iterations [--profile-step-start, --profile-step-end)")
    g.add_argument("--pytorch-profiler-collect-shapes", action="store_true")
    g.add_argument("--pytorch-profiler-collect-callstack", action="store_true")
    g.add_argument("--use-torch-fsdp2", action="store_true", help="shard with torch.distributed.fsdp.fully_shard instead of the in-house FSDP units")
    g.add_argument("--log-progress", action="store_true", help="append job start / checkpoint lines with cumulative FLOPs to <save>/progress.txt")
    g.add_argument("--record-memory-history", action="store_true", help="torch.cuda.memory._record_memory_history + snapshot dump at exit")
    g.add_argument("--memory-snapshot-path", default="snapshot.pickle")
    g.add_argument("--trace-spans", default=None, help="JSON-lines file for job/startup/train spans (core/telemetry)")
    g.add_argument("--prometheus-port", type=int, default=None)
    g.add_argument("--use-checkpoint-args", action="store_true", help="take the model architecture arguments from the checkpoint in --load")
    g.add_argument("--non-persistent-save-interval", type=int, default=None, help="local (node-storage) recovery checkpoint every N iterations")
    g.add_argument("--non-persistent-local-ckpt-dir", default=None)
    g.add_argument("--nccl-flight-recorder-dir", default=None, help="dump the last collectives of every rank here when the NCCL watchdog fires")
    g.add_argument("--check-weight-hash-across-dp-replicas-interval", type=int, default=None)
    g.add_argument("--log-straggler", action="store_true")
    g.add_argument("--eval-iters", type=int, default=0)
    g.add_argument("--eval-interval", type=int, default=1000)
    g.add_argument("--profile", action="store_true")
    g.add_argument("--profile-step-start", type=int, default=10)
    g.add_argument("--profile-step-end", type=int, default=12)
    g.add_argument("--error-injection-rate", type=int, default=0)
    g.add_argument("--rerun-mode", default="disabled", choices=["disabled", "validate_results", "report_stats"])
    return p


def validate_args(args, world_size: Optional[int] = None):
    world_size = world_size or int(os.environ.get("WORLD_SIZE", "1"))
    mp = args.tensor_model_parallel_size * args.pipeline_model_parallel_size * args.context_parallel_size
    if world_size % mp != 0:
        raise ValueError(f"world size {world_size} is not divisible by tp*pp*cp = {mp}")
    args.world_size = world_size
    args.data_parallel_size = world_size // mp
    args._global_batch_size_given = args.global_batch_size is not None
    if getattr(args, "step_batch_size_schedule", None) is not None:
        from ..core.num_microbatches_calculator import StepBatchsizeNumMicroBatchesCalculator as _Step

        if args._global_batch_size_given:
            raise ValueError("Cannot specify both --step-batch-size-schedule and --global-batch-size")
        args.global_batch_size = _Step._parse_schedule(args.step_batch_size_schedule, args.seq_length)[-1][1]   # the final (steady-state) batch size
    if args.global_batch_size is None:
        args.global_batch_size = args.micro_batch_size * args.data_parallel_size
    if args.global_batch_size % (args.micro_batch_size * args.data_parallel_size) != 0:
        raise ValueError("global batch size must be divisible by micro-batch-size × data-parallel size")
    if args.fp16 and args.bf16:
        raise ValueError("--fp16 and --bf16 are mutually exclusive")
    args.params_dtype = torch.bfloat16 if args.bf16 else (torch.float16 if args.fp16 else torch.float32)
    if args.bf16:
        args.accumulate_allreduce_grads_in_fp32 = args.accumulate_allreduce_grads_in_fp32
    if args.num_layers_per_virtual_pipeline_stage is not None:
        per_stage = args.num_layers // args.pipeline_model_parallel_size
        if per_stage % args.num_layers_per_virtual_pipeline_stage != 0:
            raise ValueError("layers per pipeline stage must be divisible by --num-layers-per-virtual-pipeline-stage")
        args.virtual_pipeline_model_parallel_size = per_stage // args.num_layers_per_virtual_pipeline_stage
    elif args.num_virtual_stages_per_pipeline_rank is not None:
        args.virtual_pipeline_model_parallel_size = args.num_virtual_stages_per_pipeline_rank
    else:
        args.virtual_pipeline_model_parallel_size = None
    if args.tensor_model_parallel_size == 1:
        args.sequence_parallel = False
    if args.use_rotary_position_embeddings:
        args.position_embedding_type = "rope"
    if args.recompute_activations:
        args.recompute_granularity = "selective"
    if args.overlap_param_gather and not args.use_distributed_optimizer:
        raise ValueError("--overlap-param-gather requires --use-distributed-optimizer")
    if args.train_iters is None and args.train_samples is None:
        args.train_iters = 0
    if args.train_iters is None:
        args.train_iters = args.train_samples // args.global_batch_size
        args._train_iters_from_samples = True              # recomputed against the batch-size schedule once the calculator exists (update_train_iters)
    if args.lr_warmup_fraction is not None and (args.lr_warmup_iters or args.lr_warmup_samples):
        raise ValueError("--lr-warmup-fraction is exclusive with --lr-warmup-iters/--lr-warmup-samples")
    if args.vocab_size is not None and getattr(args, "padded_vocab_size", None) is None:      # --padded-vocab-size pins it (e.g. to match a checkpoint)
        if getattr(args, "pad_vocab_size", True):
            m = args.make_vocab_size_divisible_by * args.tensor_model_parallel_size
            args.padded_vocab_size = (args.vocab_size + (getattr(args, "vocab_extra_ids", 0) or 0) + m - 1) // m * m
        else:
            args.padded_vocab_size = args.vocab_size + (getattr(args, "vocab_extra_ids", 0) or 0)      # --no-pad-vocab-size
    if args.num_query_groups is None:
        args.num_query_groups = args.num_attention_heads
    if args.expert_model_parallel_size > 1 and args.num_experts is None:
        raise ValueError("--expert-model-parallel-size > 1 requires --num-experts")
    return args


def config_classes():
    """The dataclasses whose fields become command-line options (after the hand-written, reference-named flags)."""
    from ..core.distributed import DistributedDataParallelConfig
    from ..core.optimizer import OptimizerConfig
    from ..core.transformer.transformer_config import TransformerConfig

    return (TransformerConfig, OptimizerConfig, DistributedDataParallelConfig)          # TransformerConfig includes ModelParallelConfig's fields


def build_full_parser(extra_args_provider=None) -> argparse.ArgumentParser:
    """Hand-written flags (reference names) + one generated flag for every remaining config-dataclass field + ``--yaml-cfg``."""
    from .argument_utils import add_dataclass_arguments

    parser = build_parser()
    if not any(a.dest == "yaml_cfg" for a in parser._actions):
        parser.add_argument("--yaml-cfg", type=str, default=None, help="YAML file whose (optionally nested) keys are argument names; command-line flags win")
    if extra_args_provider is not None:
        parser = extra_args_provider(parser)
    for cls in config_classes():
        add_dataclass_arguments(parser, cls, title=f"{cls.__name__} (generated)")
    if not os.environ.get("MB200_NO_REFERENCE_FLAGS"):
        from .reference_flags import add_reference_compat_flags

        add_reference_compat_flags(parser)             # the reference's remaining flag names (accepted; wired / native / inert: see reference_flags.py)
    return parser


def parse_args(argv=None, extra_args_provider=None, ignore_unknown_args: bool = False):
    import sys as _sys

    parser = build_full_parser(extra_args_provider)
    argv_list = list(_sys.argv[1:] if argv is None else argv)
    args, unknown = parser.parse_known_args(argv_list)
    if unknown and not ignore_unknown_args:
        parser.error(f"unrecognized arguments: {' '.join(unknown)}")
    if getattr(args, "yaml_cfg", None):
        from .argument_utils import apply_yaml, explicit_dests, load_yaml_config

        apply_yaml(args, load_yaml_config(args.yaml_cfg), parser, explicit_dests(parser, argv_list))
    if not os.environ.get("MB200_NO_REFERENCE_FLAGS"):
        from .reference_flags import apply_reference_compat, inert_flags_in_use

        apply_reference_compat(args)
        inert = inert_flags_in_use(args, parser)
        if inert:
            msg = f"reference flags accepted but without effect in this build: {' '.join(inert)}"
            if getattr(args, "strict_reference_flags", False):
                parser.error(msg)
            if int(os.environ.get("RANK", "0")) == 0:
                print(f"WARNING: {msg}", flush=True)
    if args.model is not None:
        from ..models.presets import PRESETS

        p = PRESETS[args.model]
        for k_arg, k_p in (("num_layers", "num_layers"), ("hidden_size", "hidden_size"), ("ffn_hidden_size", "ffn_hidden_size"),
                           ("num_attention_heads", "num_attention_heads"), ("num_query_groups", "num_query_groups"), ("kv_channels", "kv_channels"),
                           ("seq_length", "seq_length"), ("vocab_size", "vocab_size")):
            if getattr(args, k_arg, None) is None:
                setattr(args, k_arg, p[k_p])
        args.normalization = p["normalization"]
        args.swiglu = p["swiglu"]
        args.add_bias_linear = p["bias"]
        args.untie_embeddings_and_output_weights = p["untie"]
        if p["rotary_base"]:
            args.position_embedding_type, args.rotary_base = "rope", p["rotary_base"]
        if p.get("num_moe_experts") and args.num_experts is None:
            args.num_experts, args.moe_router_topk = p["num_moe_experts"], p["moe_router_topk"]
        if args.max_position_embeddings is None:
            args.max_position_embeddings = args.seq_length
    return args


def parse_and_validate_args(argv=None, **kw):
    args = parse_args(argv, **kw)
    if getattr(args, "use_checkpoint_args", False) or getattr(args, "use_mp_args_from_checkpoint_args", False):
        # --use-checkpoint-args: the checkpoint describes the architecture; --use-mp-args-from-checkpoint-args: and the model-parallel layout it was saved with
        from .checkpointing import load_args_from_checkpoint

        load_args_from_checkpoint(args, architecture=bool(getattr(args, "use_checkpoint_args", False)),
                                  model_parallel=bool(getattr(args, "use_mp_args_from_checkpoint_args", False)))
    return validate_args(args)


def core_transformer_config_from_args(args):
    from ..core.transformer.transformer_config import TransformerConfig

    act = F.silu if args.swiglu else F.gelu
    if args.squared_relu:
        from ..ops.reference import squared_relu as act  # noqa: F811
    if getattr(args, "openai_gelu", False):                      # tanh-approximated GeLU
        act = lambda x: F.gelu(x, approximate="tanh")            # noqa: E731
    if getattr(args, "quick_geglu", False):                      # x * sigmoid(1.702 x), gated
        act = lambda x: x * torch.sigmoid(1.702 * x)             # noqa: E731
    kw = dict(
        num_layers=args.num_layers, hidden_size=args.hidden_size, ffn_hidden_size=args.ffn_hidden_size, num_attention_heads=args.num_attention_heads,
        num_query_groups=args.num_query_groups, kv_channels=args.kv_channels, hidden_dropout=args.hidden_dropout, attention_dropout=args.attention_dropout,
        layernorm_epsilon=args.norm_epsilon, add_bias_linear=args.add_bias_linear, add_qkv_bias=args.add_qkv_bias, gated_linear_unit=args.swiglu,
        activation_func=act, normalization=args.normalization, qk_layernorm=args.qk_layernorm, init_method_std=args.init_method_std,
        tensor_model_parallel_size=args.tensor_model_parallel_size, pipeline_model_parallel_size=args.pipeline_model_parallel_size,
        virtual_pipeline_model_parallel_size=args.virtual_pipeline_model_parallel_size, context_parallel_size=args.context_parallel_size,
        expert_model_parallel_size=args.expert_model_parallel_size, expert_tensor_parallel_size=args.expert_tensor_parallel_size,
        sequence_parallel=args.sequence_parallel, fp16=args.fp16, bf16=args.bf16, params_dtype=args.params_dtype,
        pipeline_dtype=args.params_dtype if args.pipeline_model_parallel_size > 1 else None,
        use_cpu_initialization=args.use_cpu_initialization or not torch.cuda.is_available(),
        gradient_accumulation_fusion=args.gradient_accumulation_fusion, recompute_granularity=args.recompute_granularity,
        recompute_method=args.recompute_method, recompute_num_layers=args.recompute_num_layers, recompute_modules=args.recompute_modules,
        deterministic_mode=args.deterministic_mode, overlap_p2p_comm=args.overlap_p2p_comm, batch_p2p_comm=not args.overlap_p2p_comm,
        calculate_per_token_loss=args.calculate_per_token_loss, tp_comm_overlap=args.tp_comm_overlap, cp_comm_type=args.cp_comm_type,
        bias_activation_fusion=getattr(args, "bias_swiglu_fusion", True) if args.swiglu else getattr(args, "bias_gelu_fusion", True), bias_dropout_fusion=True,
        apply_rope_fusion=getattr(args, "apply_rope_fusion", True), masked_softmax_fusion=True,
        fp8=args.fp8_format, fp8_recipe=args.fp8_recipe if args.fp8_format else "delayed",
    )
    if args.num_experts:
        kw.update(num_moe_experts=args.num_experts, moe_router_topk=args.moe_router_topk, moe_router_load_balancing_type=args.moe_router_load_balancing_type,
                  moe_aux_loss_coeff=args.moe_aux_loss_coeff, moe_z_loss_coeff=args.moe_z_loss_coeff, moe_token_dispatcher_type=args.moe_token_dispatcher_type,
                  moe_grouped_gemm=args.moe_grouped_gemm, moe_ffn_hidden_size=args.moe_ffn_hidden_size, moe_layer_freq=args.moe_layer_freq,
                  moe_shared_expert_intermediate_size=args.moe_shared_expert_intermediate_size, moe_expert_capacity_factor=args.moe_expert_capacity_factor)
    # every other TransformerConfig field that has a (generated) flag and was given a non-default value
    import dataclasses as _dc

    for f in _dc.fields(TransformerConfig):
        if f.name in kw or not f.init or not hasattr(args, f.name):
            continue
        v = getattr(args, f.name)
        default = f.default if f.default is not _dc.MISSING else (f.default_factory() if f.default_factory is not _dc.MISSING else None)
        if v is not None and v != default:
            kw[f.name] = v
    if getattr(args, "quick_geglu", False):
        kw["gated_linear_unit"] = True
    if getattr(args, "openai_gelu", False) or getattr(args, "quick_geglu", False):
        kw["bias_activation_fusion"] = False                      # the fused bias+activation kernels cover gelu / silu / quick_gelu / squared_relu only
    if getattr(args, "init_method_xavier_uniform", False):
        kw["init_method"] = torch.nn.init.xavier_uniform_
        kw["output_layer_init_method"] = torch.nn.init.xavier_uniform_
    recipe_file = getattr(args, "te_precision_config_file", None) or getattr(args, "kitchen_config_file", None)
    if recipe_file:
        kw["quant_recipe"] = recipe_file                        # YAML of per-layer precision matchers (core/quantization)
    if getattr(args, "multi_latent_attention", False):
        # DeepSeek-style attention: its dimensions / YaRN parameters come from the reference-named flags (--q-lora-rank, --kv-lora-rank, --qk-head-dim,
        # --qk-pos-emb-head-dim, --v-head-dim, --rope-type, --rotary-scaling-factor, --mscale, --mscale-all-dim, --yarn-*) when given
        from ..core.transformer.transformer_config import MLATransformerConfig

        for f in _dc.fields(MLATransformerConfig):
            if f.name in kw or not f.init:
                continue
            v = getattr(args, f.name, None)
            if v is not None and f.name not in {ff.name for ff in _dc.fields(TransformerConfig)}:
                kw[f.name] = v
        if kw.get("rope_type", "yarn") != "yarn":
            kw["apply_rope_fusion"] = False
        return MLATransformerConfig(**kw)
    return TransformerConfig(**kw)


def ddp_config_from_args(args):
    """``DistributedDataParallelConfig`` from the hand-written flags plus every same-named (generated or reference-compatible) option."""
    import dataclasses as _dc

    from ..core.distributed import DistributedDataParallelConfig

    kw = dict(grad_reduce_in_fp32=args.accumulate_allreduce_grads_in_fp32, overlap_grad_reduce=args.overlap_grad_reduce, overlap_param_gather=args.overlap_param_gather,
              use_distributed_optimizer=args.use_distributed_optimizer, check_for_nan_in_grad=args.check_for_nan_in_loss_and_grad, bucket_size=args.ddp_bucket_size)
    for f in _dc.fields(DistributedDataParallelConfig):
        if f.name in kw or not hasattr(args, f.name):
            continue
        v = getattr(args, f.name)
        default = f.default if f.default is not _dc.MISSING else None
        if v is not None and v != default:
            kw[f.name] = v
    return DistributedDataParallelConfig(**kw)
