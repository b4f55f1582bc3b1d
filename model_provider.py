"""Top-level ``model_provider`` (reference ``model_provider.py``): dispatch to the family-specific builder by ``args.model_type`` / flags."""
from gpt_builders import gpt_builder


def model_provider(model_builder=gpt_builder, pre_process=True, post_process=True, vp_stage=None, config=None, pg_collection=None):
    from megatron_b200.training.training import get_args

    return model_builder(get_args(), pre_process, post_process, vp_stage, config=config, pg_collection=pg_collection)
