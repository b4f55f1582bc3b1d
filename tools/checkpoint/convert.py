"""Checkpoint conversion CLI (reference ``tools/checkpoint/convert.py`` with loader/saver plug-ins).

    python tools/checkpoint/convert.py --from hf --to megatron --load HF_DIR --save OUT_DIR --preset llama3_8b
    python tools/checkpoint/convert.py --from megatron --to hf --load CKPT_DIR --save OUT_DIR --preset llama3_8b [--target-tp 1]

HF side: a directory of ``*.safetensors`` / ``pytorch_model*.bin`` holding ``LlamaForCausalLM`` weights.  Megatron side: this
framework's ``torch_dist`` distributed checkpoint (TP/PP resharding happens on load through the sharded state dict)."""
import argparse
import glob
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _load_hf_dir(path):
    sd = {}
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if files:
        from safetensors.torch import load_file

        for f in files:
            sd.update(load_file(f))
        return sd
    for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))) or sorted(glob.glob(os.path.join(path, "*.pt"))):
        sd.update(torch.load(f, map_location="cpu"))
    if not sd:
        raise FileNotFoundError(f"no HF weight files under {path}")
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from", dest="src", choices=["hf", "megatron"], required=True)
    ap.add_argument("--to", dest="dst", choices=["hf", "megatron"], required=True)
    ap.add_argument("--load", required=True)
    ap.add_argument("--save", required=True)
    ap.add_argument("--preset", default="llama3_8b")
    args = ap.parse_args()
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29581")
    dist.init_process_group("gloo", rank=0, world_size=1)
    from megatron_b200.core import dist_checkpointing
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.export.hf_llama import hf_llama_to_megatron, megatron_to_hf_llama
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.models.presets import build_gpt_model

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(0)
    model, cfg, p = build_gpt_model(args.preset, use_cpu_initialization=True, perform_initialization=False)
    os.makedirs(args.save, exist_ok=True)
    if args.src == "hf":
        sd = hf_llama_to_megatron(_load_hf_dir(args.load), cfg.num_attention_heads, cfg.num_query_groups, cfg.kv_channels, tie_embeddings=not p["untie"])
        missing, unexpected = model.load_state_dict(sd, strict=False)
        missing = [k for k in missing if "_extra_state" not in k]
        assert not missing, f"missing keys: {missing[:5]}"
        dist_checkpointing.save(model.sharded_state_dict(), args.save)
        print(f"wrote megatron checkpoint to {args.save}")
    else:
        loaded = dist_checkpointing.load(model.sharded_state_dict(), args.load)
        model.load_state_dict(loaded, strict=False)
        sd = {k: v for k, v in model.state_dict().items() if isinstance(v, torch.Tensor) and "_extra_state" not in k}
        hf = megatron_to_hf_llama(sd, cfg.num_attention_heads, cfg.num_query_groups, cfg.kv_channels)
        torch.save(hf, os.path.join(args.save, "pytorch_model.bin"))
        print(f"wrote {len(hf)} HF tensors to {args.save}/pytorch_model.bin")


if __name__ == "__main__":
    main()
