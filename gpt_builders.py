"""GPT / Llama / Mixtral model builder from parsed arguments (reference ``gpt_builders.py``)."""
from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_decoder_block_spec, get_gpt_layer_local_spec
from megatron_b200.core.models.gpt.gpt_model import GPTModel
from megatron_b200.training.arguments import core_transformer_config_from_args


def gpt_builder(args, pre_process=True, post_process=True, vp_stage=None, config=None, pg_collection=None):
    config = config or core_transformer_config_from_args(args)
    if args.num_experts:
        spec = get_gpt_decoder_block_spec(config, vp_stage=vp_stage)
    else:
        spec = get_gpt_layer_local_spec(normalization=args.normalization, qk_layernorm=args.qk_layernorm, multi_latent_attention=args.multi_latent_attention)
    mtp_spec = None
    if getattr(args, "mtp_num_layers", None):
        from megatron_b200.core.transformer.multi_token_prediction import get_mtp_block_spec

        mtp_spec = get_mtp_block_spec(config, spec if not args.num_experts else get_gpt_layer_local_spec(normalization=args.normalization), vp_stage=vp_stage)
    return GPTModel(
        config=config, transformer_layer_spec=spec, vocab_size=args.padded_vocab_size, max_sequence_length=args.max_position_embeddings or args.seq_length,
        pre_process=pre_process, post_process=post_process, parallel_output=True, share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
        position_embedding_type=args.position_embedding_type, rotary_percent=args.rotary_percent, rotary_base=args.rotary_base, rope_scaling=args.use_rope_scaling,
        rope_scaling_factor=args.rope_scaling_factor, vp_stage=vp_stage, mtp_block_spec=mtp_spec, pg_collection=pg_collection,
    )
