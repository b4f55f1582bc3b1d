"""SOAP, parameter-layout rules, safe unpickling, storage shim, typing helpers, rank / SLURM probes."""
import io
import os
import pickle

import numpy as np
import pytest
import torch


def test_soap_reduces_to_adam_without_bases_and_beats_it_on_ill_conditioned_least_squares():
    from megatron_b200.core.optimizer.soap import init_soap_state, is_soap_param, soap_step

    torch.manual_seed(0)
    assert is_soap_param(torch.nn.Parameter(torch.zeros(4, 4))) and not is_soap_param(torch.nn.Parameter(torch.zeros(4)))
    # both sides above max_precond_dim → identity bases → exactly AdamW
    W, Wa = torch.randn(6, 5), None
    Wa = W.clone().requires_grad_(True)
    opt = torch.optim.AdamW([Wa], lr=0.01, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    st = init_soap_state(W, max_precond_dim=0)
    assert st == {}
    m, v = torch.zeros_like(W), torch.zeros_like(W)
    for step in range(1, 6):
        g = torch.randn(6, 5)
        Wa.grad = g.clone()
        opt.step()
        soap_step(W, g, m, v, st, lr=0.01, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=step)
    assert torch.allclose(W, Wa.detach(), atol=1e-6)

    # ill-conditioned least squares  min ‖X W − Y‖²,  X with a 1e3 spread of feature scales rotated by a random orthogonal matrix:
    # Adam's diagonal preconditioner cannot undo the rotation, SOAP's eigenbasis can
    Q, _ = torch.linalg.qr(torch.randn(16, 16))
    X = (torch.randn(256, 16) * torch.logspace(0, -3, 16)) @ Q
    Y = X @ torch.randn(16, 8)

    def run(kind):
        W = torch.zeros(16, 8)
        m, v = torch.zeros_like(W), torch.zeros_like(W)
        st = init_soap_state(W, 64 if kind == "soap" else 0)
        for step in range(1, 201):
            r = X @ W - Y
            g = 2 * X.t() @ r / r.numel()
            soap_step(W, g, m, v, st, lr=0.05, beta1=0.9, beta2=0.95, eps=1e-12, weight_decay=0.0, step=step, precondition_frequency=5)
        return ((X @ W - Y) ** 2).mean().item()

    adam, soap = run("adam"), run("soap")
    assert soap < 0.5 * adam, (soap, adam)


def test_soap_through_the_optimizer_stack_with_state_roundtrip():
    from megatron_b200.core.optimizer.optimizer import FP32Optimizer
    from megatron_b200.core.optimizer.optimizer_config import OptimizerConfig

    torch.manual_seed(1)
    lin = torch.nn.Linear(8, 4)
    cfg = OptimizerConfig(optimizer="soap", lr=0.01, weight_decay=0.0, clip_grad=0.0, soap_precondition_frequency=2)
    groups = [{"params": list(lin.parameters()), "lr": 0.01, "weight_decay": 0.0}]
    opt = FP32Optimizer(groups, cfg)
    x = torch.randn(32, 8)
    for _ in range(5):
        lin.zero_grad()
        lin(x).square().mean().backward()
        opt.step()
    w_slot = next(s for s in opt.slots if s.master.dim() == 2)
    assert {"L", "R", "QL", "QR"} <= set(w_slot.extra) and next(s for s in opt.slots if s.master.dim() == 1).extra is None
    q = w_slot.extra["QL"]
    assert torch.allclose(q.t() @ q, torch.eye(4), atol=1e-4)
    sd = opt.state_dict()
    lin2 = torch.nn.Linear(8, 4)
    lin2.load_state_dict(lin.state_dict())
    opt2 = FP32Optimizer([{"params": list(lin2.parameters()), "lr": 0.01, "weight_decay": 0.0}], cfg)
    opt2.load_state_dict(sd)
    for o, l in ((opt, lin), (opt2, lin2)):
        l.zero_grad()
        l(x).square().mean().backward()
        o.step()
    assert torch.allclose(lin.weight, lin2.weight, atol=1e-7)


def test_param_layout_alignment_buckets_and_whole_param_shards():
    from megatron_b200.core.optimizer import param_layout as pl

    assert pl.pad_to_divisor(65, 64) == 128 and pl.pad_param_start(64) == 64
    assert pl.bucket_end_divisor(6, False) == 384 and pl.bucket_end_divisor(8, True) == 65536
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (100, 30, 700, 64, 5)]
    lay = pl.compute_per_buffer_layout(ps, data_parallel_world_size=4, bucket_size=600)
    # reverse order, 64-aligned starts, bucket ends divisible by lcm(4, 128)
    order = [p for p in lay.param_index_map]
    assert order[0] is ps[-1]
    for p, (s, e, b) in lay.param_index_map.items():
        assert s % 64 == 0 and e - s == p.numel() and lay.bucket_indices[b][0] <= s and e <= lay.bucket_indices[b][1]
    assert all(e % 128 == 0 and (e - s) % 4 == 0 for s, e in lay.bucket_indices) and len(lay.bucket_indices) == 2
    assert sum(lay.per_bucket_numel_unpadded) >= sum(p.numel() for p in ps)
    assert lay.shard_range(0, 1)[0] == lay.bucket_indices[0][0] + (lay.bucket_indices[0][1] - lay.bucket_indices[0][0]) // 4
    # whole-parameter shards: nothing straddles a shard boundary
    lw = pl.compute_per_buffer_layout(ps, data_parallel_world_size=2, whole_params_per_shard=True)
    cap = lw.bucket_indices[0][1] // 2
    spans = sorted((s, e) for s, e, _ in lw.param_index_map.values())
    assert all(s // cap == (e - 1) // cap for s, e in spans) and all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    full = pl.compute_full_layout([torch.nn.Parameter(torch.zeros(8, dtype=torch.bfloat16)), torch.nn.Parameter(torch.zeros(8))], 2)
    assert len(full.layouts) == 2


class _Evil:
    def __reduce__(self):
        return (os.system, ("echo pwned",))


def test_safe_unpickler_accepts_tensors_and_rejects_callables(tmp_path):
    from megatron_b200.core import safe_globals as sg

    payload = {"scale": torch.arange(4.0), "amax_history": [1, 2.5, "x"], "shape": torch.Size([2, 3])}
    out = sg.safe_pickle_loads(pickle.dumps(payload))
    assert torch.equal(out["scale"], payload["scale"]) and out["amax_history"] == [1, 2.5, "x"]
    with pytest.raises(pickle.UnpicklingError):
        sg.safe_pickle_loads(pickle.dumps(_Evil()))
    p = tmp_path / "obj.npy"
    np.save(p, np.array([{"a": 1}, [1, 2]], dtype=object), allow_pickle=True)
    arr = sg.safe_numpy_load(p, allow_pickle=True)
    assert arr[0] == {"a": 1}
    np.save(p, np.array([_Evil()], dtype=object), allow_pickle=True)
    with pytest.raises(pickle.UnpicklingError):
        sg.safe_numpy_load(p, allow_pickle=True)
    sg.register_safe_globals()
    from argparse import Namespace

    buf = io.BytesIO()
    torch.save({"args": Namespace(lr=0.1)}, buf)
    assert torch.load(io.BytesIO(buf.getvalue()), weights_only=True)["args"].lr == 0.1


def test_msc_shim_typed_helpers_and_rank_probes(tmp_path, monkeypatch):
    from megatron_b200.core import _rank_utils, _slurm_utils, typed_torch
    from megatron_b200.core.msc_utils import MaybeMultiStorageClient, MultiStorageClientFeature

    assert not MultiStorageClientFeature.is_enabled()
    msc = MaybeMultiStorageClient()
    f = tmp_path / "a.txt"
    with msc.open(str(f), "w") as fh:
        fh.write("hi")
    assert msc.os.path.exists(str(f)) and msc.path_isdir(str(tmp_path)) and MultiStorageClientFeature.is_msc_url("msc://p/x")
    MultiStorageClientFeature.enable()
    try:
        with pytest.raises(ImportError):
            msc.open
    finally:
        MultiStorageClientFeature.disable()

    lin = torch.nn.Linear(2, 2)
    assert typed_torch.apply_module(lin)(torch.zeros(1, 2)).shape == (1, 2)
    with pytest.raises(TypeError):
        typed_torch.apply_module(lambda x: x)
    assert typed_torch.not_none(3) == 3
    with pytest.raises(ValueError):
        typed_torch.not_none(None)

    def src(a: int, b: str = "x") -> float:
        return 0.0

    @typed_torch.copy_signature(src)
    def wrapper(*args, **kwargs):
        return src(*args, **kwargs)

    import inspect

    assert list(inspect.signature(wrapper).parameters) == ["a", "b"] and wrapper(1) == 0.0

    for k in ("RANK", "WORLD_SIZE", "SLURM_JOB_ID", "SLURM_PROCID", "SLURM_NTASKS", "SLURM_LOCALID"):
        monkeypatch.delenv(k, raising=False)
    assert _rank_utils.safe_get_rank() == 0 and _rank_utils.safe_get_world_size() == 1 and not _slurm_utils.is_slurm_job()
    monkeypatch.setenv("SLURM_JOB_ID", "7"), monkeypatch.setenv("SLURM_PROCID", "3"), monkeypatch.setenv("SLURM_NTASKS", "8"), monkeypatch.setenv("SLURM_LOCALID", "1")
    assert (_rank_utils.safe_get_rank(), _rank_utils.safe_get_world_size(), _slurm_utils.resolve_slurm_local_rank()) == (3, 8, 1)
    monkeypatch.setenv("RANK", "5")
    assert _rank_utils.safe_get_rank() == 5                         # torchrun's variables win over SLURM's
    import logging

    seen = []
    lg = logging.getLogger("t")
    lg.log = lambda *a, **k: seen.append(a)
    _rank_utils.log_single_rank(lg, logging.INFO, "x", rank=5)
    _rank_utils.log_single_rank(lg, logging.INFO, "y", rank=0)
    _rank_utils.log_single_rank(lg, logging.INFO, "z", rank=-3)
    assert [a[1] for a in seen] == ["x", "z"]


def test_mtp_placement_follows_the_pipeline_layout():
    from megatron_b200.core.enums import LayerType
    from megatron_b200.core.transformer.multi_token_prediction import get_mtp_layer_offset, get_mtp_num_layers_to_build, mtp_on_this_rank
    from megatron_b200.core.transformer.pipeline_parallel_layer_layout import PipelineParallelLayerLayout
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    lay = PipelineParallelLayerLayout.from_str("Et*2|t*2|m|mL", 4)             # MTP depth 0 on a stage of its own, depth 1 with the loss
    lay.validate_layer_layout(num_layers=4, mtp_num_layers=2)
    assert lay.mtp_standalone_stages() == [2]
    cfg = TransformerConfig(num_layers=4, hidden_size=16, num_attention_heads=2, mtp_num_layers=2, pipeline_model_parallel_size=4, pipeline_dtype=torch.float32, use_cpu_initialization=True)
    cfg.pipeline_model_parallel_layout = lay
    assert [mtp_on_this_rank(cfg, pp_rank=r) for r in range(4)] == [False, False, True, True]
    assert [get_mtp_num_layers_to_build(cfg, pp_rank=r) for r in range(4)] == [0, 0, 1, 1] and [get_mtp_layer_offset(cfg, pp_rank=r) for r in (2, 3)] == [0, 1]
    cfg.pipeline_model_parallel_layout = "Et*2|t*2|m|mL"                       # the string form is parsed on demand
    assert mtp_on_this_rank(cfg, pp_rank=2) and not mtp_on_this_rank(cfg, pp_rank=1)
    cfg.pipeline_model_parallel_layout = None
    assert mtp_on_this_rank(cfg) and get_mtp_num_layers_to_build(cfg) == 2     # no layout, no process groups: single stage holds them
    cfg.mtp_num_layers = None
    assert not mtp_on_this_rank(cfg)
    bad = PipelineParallelLayerLayout.from_str("Etm|ttL", 2)
    try:
        bad.validate_layer_layout(num_layers=3, mtp_num_layers=1)
        raise SystemExit("an MTP layer in front of decoder layers was accepted")
    except AssertionError:
        pass
    vpp = PipelineParallelLayerLayout.from_str("Et|t|t|t|m|L", 2)              # 2 ranks x 3 virtual chunks; MTP on rank 0's third chunk
    cfg2 = TransformerConfig(num_layers=4, hidden_size=16, num_attention_heads=2, mtp_num_layers=1, pipeline_model_parallel_size=2, pipeline_dtype=torch.float32, use_cpu_initialization=True)
    cfg2.pipeline_model_parallel_layout = vpp
    assert vpp.layout[0][2] == [LayerType.mtp] and mtp_on_this_rank(cfg2, pp_rank=0) and not mtp_on_this_rank(cfg2, pp_rank=1)
    assert mtp_on_this_rank(cfg2, ignore_virtual=False, vp_stage=2, pp_rank=0) and not mtp_on_this_rank(cfg2, ignore_virtual=False, vp_stage=0, pp_rank=0)
