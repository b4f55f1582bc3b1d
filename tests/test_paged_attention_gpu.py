"""Paged-KV decode kernels (csrc/paged_attention.cu) vs the fp32 formulation: fused K/V append through the block table, flash-decoding with split-KV and GQA."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B, hq, hk, d, block_size, lens, num_blocks=None, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    width = max((l + block_size - 1) // block_size for l in lens)
    num_blocks = num_blocks or B * width + 3
    k_pool = torch.randn(num_blocks, block_size, hk, d, device=dev).bfloat16()
    v_pool = torch.randn(num_blocks, block_size, hk, d, device=dev).bfloat16()
    perm = torch.randperm(num_blocks)[: B * width].view(B, width).to(dev)      # scattered, non-contiguous pages
    q = torch.randn(B, hq, d, device=dev).bfloat16()
    return q, k_pool, v_pool, perm.to(torch.int32), torch.tensor(lens, device=dev, dtype=torch.int32)


def _ref(q, k_pool, v_pool, table, lengths, scale):
    B, hq, d = q.shape
    hk, bs = k_pool.shape[2], k_pool.shape[1]
    out = torch.zeros(B, hq, d, device=q.device)
    for b in range(B):
        L = int(lengths[b])
        nb = (L + bs - 1) // bs
        K = k_pool[table[b, :nb].long()].reshape(-1, hk, d)[:L].float()
        V = v_pool[table[b, :nb].long()].reshape(-1, hk, d)[:L].float()
        for h in range(hq):
            s = (K[:, h // (hq // hk)] @ q[b, h].float()) * scale
            out[b, h] = torch.softmax(s, 0) @ V[:, h // (hq // hk)]
    return out


@pytest.mark.parametrize("B,hq,hk,d,bs,lens", [
    (4, 8, 2, 128, 16, [1, 17, 300, 1000]),          # GQA 4:1, ragged lengths, one-token history
    (3, 4, 4, 128, 64, [64, 65, 4096]),              # MHA, block-boundary lengths, long history (several splits)
    (2, 16, 2, 64, 32, [129, 777]),                  # GQA 8:1, d 64
    (32, 32, 8, 128, 16, [500 + 13 * i for i in range(32)]),   # serving-sized batch (Llama-3 8B heads)
])
def test_paged_decode_matches_reference(B, hq, hk, d, bs, lens):
    from megatron_b200 import ops

    assert hasattr(ops.ext(), "paged_decode"), "paged attention kernels not built"
    q, kp, vp, table, lengths = _setup(B, hq, hk, d, bs, lens)
    scale = 1.0 / math.sqrt(d)
    out = ops.paged_attention_decode(q, kp, vp, table, lengths, scale, max(lens))
    ref = _ref(q, kp, vp, table, lengths, scale)
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err < 2e-2, err


def test_paged_kv_append_writes_the_right_page():
    from megatron_b200 import ops

    B, hk, d, bs = 5, 2, 128, 16
    q, kp, vp, table, _ = _setup(B, 4, hk, d, bs, [40] * B)
    k0, v0 = kp.clone(), vp.clone()
    pos = torch.tensor([0, 15, 16, 31, 39], device="cuda", dtype=torch.int32)
    kn = torch.randn(B, hk, d, device="cuda").bfloat16()
    vn = torch.randn(B, hk, d, device="cuda").bfloat16()
    ops.paged_kv_append(kn, vn, kp, vp, table, pos)
    torch.cuda.synchronize()
    for b in range(B):
        blk, off = int(table[b, int(pos[b]) // bs]), int(pos[b]) % bs
        assert torch.equal(kp[blk, off], kn[b]) and torch.equal(vp[blk, off], vn[b])
        k0[blk, off], v0[blk, off] = kn[b], vn[b]
    assert torch.equal(kp, k0) and torch.equal(vp, v0), "append touched other slots"


def test_dynamic_engine_cuda_graph_decode_matches_eager():
    """Greedy generation with the CUDA-graphed decode step equals the eager engine token for token (tiny Llama, 6 requests of different lengths)."""
    import os

    import torch.distributed as dist

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.inference.engine import DynamicInferenceEngine
    from megatron_b200.core.inference.sampling import SamplingParams
    from megatron_b200.models.presets import build_gpt_model

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29993")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    if not ps.is_initialized():
        ps.initialize_model_parallel()
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed

    model_parallel_cuda_manual_seed(7)
    m, cfg, p = build_gpt_model("tiny_llama", bf16=True, params_dtype=torch.bfloat16, kv_channels=128, num_attention_heads=2, num_query_groups=1, hidden_size=256)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(3)
    prompts = [torch.randint(0, p["vocab_size"], (n,), generator=g).tolist() for n in (5, 17, 33, 8, 64, 21)]
    outs = []
    for graphs in (False, True):
        eng = DynamicInferenceEngine(m, num_blocks=128, block_size=16, max_running=8, vocab_size=p["vocab_size"], enable_cuda_graphs=graphs)
        ids = [eng.add_request(pr, SamplingParams(temperature=0.0, num_tokens_to_generate=40)) for pr in prompts]
        done = eng.run_until_done()
        outs.append([done[i].generated_tokens for i in ids])
        if graphs:
            assert eng.graph_replays > 0 and len(eng._graphs) >= 1
    assert outs[0] == outs[1]


def test_paged_decode_is_batch_invariant_in_the_mode():
    """Batch-invariant mode: a request's decode attention is bitwise the same alone, in a batch of 8, and next to a much longer request (which changes the
    default split sizing); without the mode the bits may differ (they do for this shape) while values agree to rounding."""
    import math

    from megatron_b200 import ops
    from megatron_b200.core.transformer.custom_layers.batch_invariant_kernels import set_batch_invariant_mode

    torch.manual_seed(0)
    hq, hk, d, bs = 32, 8, 128, 16
    lens = [700, 1500, 333, 4096, 64, 2049, 900, 17]
    width = (max(lens) + bs - 1) // bs
    nb = sum((l + bs - 1) // bs for l in lens) + 1
    kp = torch.randn(nb, bs, hk, d, device="cuda").bfloat16()
    vp = torch.randn_like(kp)
    table = torch.zeros(len(lens), width, dtype=torch.int32, device="cuda")
    nxt = 1
    for i, l in enumerate(lens):
        n = (l + bs - 1) // bs
        table[i, :n] = torch.arange(nxt, nxt + n, dtype=torch.int32, device="cuda")
        nxt += n
    q = torch.randn(len(lens), hq, d, device="cuda").bfloat16()
    L = torch.tensor(lens, dtype=torch.int32, device="cuda")
    scale = 1 / math.sqrt(d)

    def run(idx):
        sel = torch.tensor(idx, device="cuda")
        return ops.paged_attention_decode(q[sel], kp, vp, table[sel], L[sel], scale, max(lens[i] for i in idx))

    with set_batch_invariant_mode(True):
        full = run(list(range(8)))
        for i in (0, 2, 3, 7):
            assert torch.equal(run([i])[0], full[i]), f"request {i} alone differs from the batch of 8"
        assert torch.equal(run([2, 3])[0], full[2]) and torch.equal(run([7, 0, 5])[1], full[0])
    loose = run(list(range(8)))
    assert (loose.float() - full.float()).abs().max().item() < 2e-2
