#!/usr/bin/env python
"""The smallest complete training program on the core API (reference ``examples/run_simple_mcore_train_loop.py``): parallel state → GPT model → DDP →
optimizer → mock dataset → forward/backward schedule → distributed checkpoint save / load.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/run_simple_mcore_train_loop.py --tp 2      # GPUs (NCCL) or CPUs (gloo) alike
"""
import argparse
import os
import sys
import tempfile
from functools import partial

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from megatron_b200.core import dist_checkpointing, parallel_state as ps  # noqa: E402
from megatron_b200.core.datasets import BlendedMegatronDatasetBuilder, GPTDatasetConfig, MockGPTDataset  # noqa: E402
from megatron_b200.core.distributed import DistributedDataParallel, DistributedDataParallelConfig, finalize_model_grads  # noqa: E402
from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec  # noqa: E402
from megatron_b200.core.models.gpt.gpt_model import GPTModel  # noqa: E402
from megatron_b200.core.optimizer import OptimizerConfig, get_megatron_optimizer  # noqa: E402
from megatron_b200.core.pipeline_parallel.schedules import get_forward_backward_func  # noqa: E402
from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed  # noqa: E402
from megatron_b200.core.tokenizers import build_tokenizer  # noqa: E402
from megatron_b200.core.transformer.transformer_config import TransformerConfig  # noqa: E402

SEQ = 64


def initialize_distributed(tp: int, pp: int):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29571")
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl" if cuda else "gloo", rank=rank, world_size=world)
    ps.initialize_model_parallel(tensor_model_parallel_size=tp, pipeline_model_parallel_size=pp)
    model_parallel_cuda_manual_seed(123)


def model_provider(cuda: bool) -> GPTModel:
    cfg = TransformerConfig(num_layers=2, hidden_size=64, num_attention_heads=4, ffn_hidden_size=128, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                            normalization="RMSNorm", use_cpu_initialization=not cuda, hidden_dropout=0.0, attention_dropout=0.0,
                            tensor_model_parallel_size=ps.get_tensor_model_parallel_world_size(), sequence_parallel=ps.get_tensor_model_parallel_world_size() > 1,
                            pipeline_dtype=torch.float32)
    return GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=128, max_sequence_length=SEQ, position_embedding_type="rope",
                    pre_process=ps.is_pipeline_first_stage(), post_process=ps.is_pipeline_last_stage())


def get_train_data_iterator(batch: int):
    tok = build_tokenizer("NullTokenizer", vocab_size=127)
    cfg = GPTDatasetConfig(random_seed=0, sequence_length=SEQ, blend=None, split="1000,0,0", tokenizer=tok, reset_position_ids=False, reset_attention_mask=False,
                           eod_mask_loss=False, create_attention_mask=False)
    train, _, _ = BlendedMegatronDatasetBuilder(MockGPTDataset, [1000, None, None], lambda: True, cfg).build()
    return iter(torch.utils.data.DataLoader(train, batch_size=batch, shuffle=False))


def forward_step_func(data_iterator, model, device):
    def loss_func(loss_mask, output):
        losses = output.float().view(-1)
        m = loss_mask.view(-1).float()
        loss = (losses * m).sum() / m.sum()
        return loss, {"lm loss": loss.detach()}

    b = next(data_iterator)
    tokens, labels, mask, pos = (b[k].to(device) for k in ("tokens", "labels", "loss_mask", "position_ids"))
    return model(tokens, pos, None, labels=labels), partial(loss_func, mask)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    initialize_distributed(a.tp, a.pp)
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    model = model_provider(cuda).to(device)
    ddp = DistributedDataParallel(model.config, DistributedDataParallelConfig(overlap_grad_reduce=False, use_distributed_optimizer=False), model)
    optim = get_megatron_optimizer(OptimizerConfig(optimizer="adam", lr=1e-3, bf16=False, fp16=False), [ddp])
    it = get_train_data_iterator(4)
    fwd_bwd = get_forward_backward_func()
    model.config.finalize_model_grads_func = finalize_model_grads
    first = last = None
    for i in range(a.iters):
        optim.zero_grad()
        ddp.zero_grad_buffer()
        losses = fwd_bwd(forward_step_func=partial(forward_step_func, device=device), data_iterator=it, model=ddp, num_microbatches=1, seq_length=SEQ, micro_batch_size=4,
                         decoder_seq_length=SEQ, forward_only=False)
        optim.step()
        if ps.is_pipeline_last_stage():
            last = float(losses[0]["lm loss"])
            first = last if first is None else first
            if dist.get_rank() == dist.get_world_size() - 1:
                print(f"iteration {i}: lm loss {last:.4f}", flush=True)
    # distributed checkpoint round trip
    ckpt = os.environ.get("CKPT_DIR") or os.path.join(tempfile.gettempdir(), "mb200_simple_ckpt")
    if dist.get_rank() == 0:
        import shutil

        shutil.rmtree(ckpt, ignore_errors=True)      # dist_checkpointing refuses to write into a non-empty directory
        os.makedirs(ckpt, exist_ok=True)
    dist.barrier()
    dist_checkpointing.save({"model": model.sharded_state_dict(prefix="")}, ckpt)
    before = {k: v.detach().clone() for k, v in model.state_dict().items() if isinstance(v, torch.Tensor)}
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
    loaded = dist_checkpointing.load({"model": model.sharded_state_dict(prefix="")}, ckpt)
    model.load_state_dict(loaded["model"])
    assert all(torch.equal(v, model.state_dict()[k]) for k, v in before.items()), "checkpoint round trip changed a tensor"
    if dist.get_rank() == dist.get_world_size() - 1:
        print(f"checkpoint round trip ok; loss {first:.4f} -> {last:.4f}", flush=True)
    ps.destroy_model_parallel()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
