"""RL sequence packing: bin packing, packed (THD) log-probs == padded log-probs through a GPT model (RoPE restarts per sequence, attention does not cross
boundaries), and a GRPO step in packed mode equals the padded step."""
import copy

import torch
import torch.nn.functional as F

from dist_utils import run_distributed


def test_pack_sequences_bins():
    from megatron_b200.rl.sequence_packing_utils import pack_sequences

    lens = [30, 70, 10, 50, 64, 5, 90]
    bins = pack_sequences(lens, 100)
    assert sorted(i for b in bins for i in b) == list(range(7)) and all(sum(lens[i] for i in b) <= 100 for b in bins)
    assert len(bins) == 4                                                      # FFD: [90, 10] [70, 30] [64, 5] [50] — the lower bound is ceil(319 / 100) = 4
    fifo = pack_sequences(lens, 100, algo="fifo")
    assert [i for b in fifo for i in b] == list(range(7)) and all(sum(lens[i] for i in b) <= 100 for b in fifo)
    assert all(len(b) <= 2 for b in pack_sequences(lens, 1000, max_sequences_per_bin=2))
    assert pack_sequences([150], 100) == [[0]]


def _packing_worker(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig
    from megatron_b200.rl.grpo import CountTokenEnv, GRPOConfig, GRPOTrainer, sequence_logprobs
    from megatron_b200.rl.sequence_packing_utils import build_packed_batch, pack_sequences, packed_sequence_logprobs, unpack

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(2)
    torch.manual_seed(2)
    cfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                            normalization="RMSNorm", use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, bias_dropout_fusion=False)
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=64, position_embedding_type="rope")
    g = torch.Generator().manual_seed(0)
    seqs = [torch.randint(1, 64, (n,), generator=g).tolist() for n in (7, 19, 3, 12, 25)]
    # padded reference: every sequence alone (right padding never influences earlier positions under a causal mask)
    want = []
    for s in seqs:
        t = torch.tensor([s])
        want.append(sequence_logprobs(model, t, 64)[0])
    bins = pack_sequences([len(s) for s in seqs], 40)
    assert len(bins) == 2
    for idx in bins:
        pb = build_packed_batch(seqs, [2] * len(seqs), idx)
        assert pb.tokens.shape[1] == sum(len(seqs[i]) for i in idx) and pb.position_ids[0, len(seqs[idx[0]])] == 0      # positions restart
        got = unpack(packed_sequence_logprobs(model, pb, 64), pb)
        for i, lp in zip(pb.seq_index, got):
            assert torch.allclose(lp, want[i], atol=1e-5), (i, (lp - want[i]).abs().max())
        # the loss mask covers exactly the completion targets and never a boundary
        cuts = torch.tensor(pb.lengths).cumsum(0) - 1
        assert pb.loss_mask[cuts[:-1]].sum() == 0 and pb.loss_mask.sum() == sum(len(seqs[i]) - 2 for i in idx)
    # a GRPO step on packed rows == the padded step (same rollouts: both trainers share the sampling RNG state)
    env = CountTokenEnv(64)
    stats = {}
    for packed in (False, True):
        m = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=64, position_embedding_type="rope")
        ref = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=64, position_embedding_type="rope")
        m.load_state_dict(model.state_dict())
        ref.load_state_dict(model.state_dict())
        opt = torch.optim.SGD(m.parameters(), lr=0.1)
        tr = GRPOTrainer(m, ref, opt, copy.deepcopy(env), GRPOConfig(group_size=2, max_new_tokens=6, use_sequence_packing=packed, packing_bin_size=32), vocab_size=64)
        torch.manual_seed(11)
        s = tr.step(2)
        stats[packed] = (float(s["loss"]), float(s["kl"]), [p.detach().clone() for p in m.parameters()])
    assert abs(stats[True][0] - stats[False][0]) < 1e-5 and abs(stats[True][1] - stats[False][1]) < 1e-6
    assert all(torch.allclose(a, b, atol=1e-5) for a, b in zip(stats[True][2], stats[False][2]))
    return True


def test_packed_logprobs_and_grpo_step_match_padded():
    assert run_distributed(_packing_worker, 1) == [True]


def _doc_mask_worker(rank, world):
    """--reset-attention-mask --reset-position-ids through the packed path == every document run on its own (per-token losses and gradients)."""
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.datasets.gpt_dataset import _get_ltor_masks_and_position_ids
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel
    from megatron_b200.core.packed_seq_params import packed_seq_params_from_documents
    from megatron_b200.core.tensor_parallel.random import model_parallel_cuda_manual_seed
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    ps.initialize_model_parallel()
    model_parallel_cuda_manual_seed(4)
    torch.manual_seed(4)
    cfg = TransformerConfig(num_layers=2, hidden_size=32, num_attention_heads=4, ffn_hidden_size=64, gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False,
                            normalization="RMSNorm", use_cpu_initialization=True, hidden_dropout=0.0, attention_dropout=0.0, bias_dropout_fusion=False)
    model = GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=64, max_sequence_length=32, position_embedding_type="rope")
    eod, b, s = 63, 2, 32
    tokens = torch.randint(0, 62, (b, s), generator=torch.Generator().manual_seed(0))
    tokens[0, 9] = tokens[0, 20] = tokens[1, 4] = tokens[1, 31] = eod            # 3 + 2 documents; one eod exactly at the end of a row
    labels = tokens.roll(-1, 1)
    pid = torch.stack([_get_ltor_masks_and_position_ids(tokens[i], eod, True, True, False, False)[2] for i in range(b)])
    # reference: every document alone through the model (a causal model with a dense mask would need the unfused path; per-document runs are unambiguous)
    bounds = [(0, 0, 10), (0, 10, 21), (0, 21, 32), (1, 0, 5), (1, 5, 32)]
    dense = torch.zeros(b, s)
    for r, lo, hi in bounds:
        t = tokens[r:r + 1, lo:hi]
        dense[r, lo:hi] = model(t, torch.arange(hi - lo)[None], None, labels=labels[r:r + 1, lo:hi])[0]
    dense.sum().backward()
    g_dense = [p.grad.clone() for p in model.parameters()]
    model.zero_grad()
    psp = packed_seq_params_from_documents(tokens, eod)
    assert psp.cu_seqlens_q.tolist() == [0, 10, 21, 32, 37, 64] and psp.max_seqlen_q == 27
    packed = model(tokens.reshape(1, b * s), pid.reshape(1, b * s), None, labels=labels.reshape(1, b * s), packed_seq_params=psp).reshape(b, s)
    packed.sum().backward()
    assert torch.allclose(packed, dense, atol=1e-5), (packed - dense).abs().max()
    assert all(torch.allclose(p.grad, g, atol=1e-5) for p, g in zip(model.parameters(), g_dense))
    return True


def test_document_masking_through_packed_attention_matches_per_document_runs():
    assert run_distributed(_doc_mask_worker, 1) == [True]
