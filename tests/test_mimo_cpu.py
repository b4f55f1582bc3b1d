"""MIMO model: colocated forward/backward, and encoders / language model on disjoint rank grids through the multi-module communicator."""
import torch
import torch.nn.functional as F

from dist_utils import run_distributed

SEQ, VOCAB, H = 24, 96, 32
IMG_TOK, AUD_TOK = 90, 91


def _tcfg(layers=1, hidden=H, heads=2, **kw):
    from megatron_b200.core.transformer.transformer_config import TransformerConfig

    base = dict(num_layers=layers, hidden_size=hidden, num_attention_heads=heads, ffn_hidden_size=2 * hidden, use_cpu_initialization=True,
                hidden_dropout=0.0, attention_dropout=0.0, add_bias_linear=True)
    base.update(kw)
    return TransformerConfig(**base)


def _vision():
    from megatron_b200.core.models.mimo import VisionModalitySubmodules
    from megatron_b200.core.models.vision.clip_vit_model import CLIPViTModel
    from megatron_b200.core.models.vision.vit_layer_specs import get_vit_layer_with_local_spec

    enc = CLIPViTModel(_tcfg(hidden=16), get_vit_layer_with_local_spec(), add_class_token=False, patch_dim=4, img_h=8, img_w=8)   # 4 tokens / image
    return VisionModalitySubmodules(encoders={"clip": enc}, input_projections=[torch.nn.Linear(16, H)])


def _audio():
    from megatron_b200.core.models.audio import AudioEncoderModel
    from megatron_b200.core.models.mimo import AudioModalitySubmodules
    from megatron_b200.core.models.vision.vit_layer_specs import get_vit_layer_with_local_spec

    enc = AudioEncoderModel(_tcfg(hidden=16), get_vit_layer_with_local_spec(), n_mels=8, max_frames=16)   # 6 frames → 3 tokens
    return AudioModalitySubmodules(encoders={"whisper": enc}, input_projections=[torch.nn.Linear(16, H)])


def _language():
    from megatron_b200.core.models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
    from megatron_b200.core.models.gpt.gpt_model import GPTModel

    cfg = _tcfg(layers=2, normalization="RMSNorm", gated_linear_unit=True, activation_func=F.silu, add_bias_linear=False)
    return GPTModel(cfg, get_gpt_layer_local_spec(normalization="RMSNorm"), vocab_size=VOCAB, max_sequence_length=SEQ, position_embedding_type="rope",
                    share_embeddings_and_output_weights=False)


def _batch(seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 80, (2, SEQ), generator=g)
    ids[0, 2:6] = IMG_TOK          # one image (4 tokens) per sample
    ids[1, 10:14] = IMG_TOK
    ids[0, 8:11] = AUD_TOK         # one clip (3 tokens) per sample
    ids[1, 1:4] = AUD_TOK
    return dict(ids=ids, labels=torch.randint(0, 80, (2, SEQ), generator=g), pos=torch.arange(SEQ)[None].expand(2, -1),
                images=torch.randn(2, 3, 8, 8, generator=g), mel=torch.randn(2, 8, 6, generator=g))


def _build(seed, grids=None):
    from megatron_b200.core.models.mimo import MimoModel, MimoModelConfig
    from megatron_b200.core.transformer.spec_utils import ModuleSpec

    torch.manual_seed(seed)
    parts = {"vision": _vision(), "audio": _audio(), "language": _language()}     # always built in this order → identical weights on every rank
    spec = lambda m: ModuleSpec(module=lambda: m)  # noqa: E731
    cfg = MimoModelConfig(language_model_spec=ModuleSpec(module=_Const, params={"value": parts["language"]}),
                          modality_submodules_spec={"vision": ModuleSpec(module=_Const, params={"value": parts["vision"]}),
                                                    "audio": ModuleSpec(module=_Const, params={"value": parts["audio"]})},
                          special_token_ids={"vision": IMG_TOK, "audio": AUD_TOK}, module_to_grid_map=grids)
    return MimoModel(cfg)


class _Const:
    """Spec adapter: ``build_module`` instantiates classes, the tests hand it pre-built modules."""

    def __new__(cls, value):
        return value


def _colocated(rank, world):
    from megatron_b200.core import parallel_state as ps

    ps.initialize_model_parallel()
    m = _build(7)
    b = _batch()
    loss = m(b["ids"], b["pos"], None, labels=b["labels"], modality_inputs={"vision": {"clip": {"x": b["images"]}}, "audio": {"whisper": {"mel": b["mel"]}}})
    loss = loss.float().mean()
    loss.backward()
    # manual reference: encoders → rows → scatter by hand
    v = m.modality_submodules["vision"]({"clip": {"x": b["images"]}})
    a = m.modality_submodules["audio"]({"whisper": {"mel": b["mel"]}})
    assert v.shape == (8, H) and a.shape == (6, H)
    ids0 = torch.where((b["ids"] == IMG_TOK) | (b["ids"] == AUD_TOK), torch.zeros_like(b["ids"]), b["ids"])
    x = m.language_model.embedding(input_ids=ids0, position_ids=b["pos"]).clone()       # [s, b, h]
    x[2:6, 0], x[10:14, 1] = v[:4], v[4:]
    x[8:11, 0], x[1:4, 1] = a[:3], a[3:]
    ref = m.language_model(input_ids=None, position_ids=b["pos"], attention_mask=None, decoder_input=x, labels=b["labels"]).float().mean()
    assert abs(ref.item() - loss.item()) < 1e-5
    grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    assert any(n.startswith("modality_submodules.vision") for n in grads) and any(n.startswith("modality_submodules.audio") for n in grads)
    return loss.item(), grads


def _distributed(rank, world):
    import torch.distributed as dist

    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.hyper_comm_grid import HyperCommGrid
    from megatron_b200.core.pipeline_parallel.multimodule_communicator import MultiModulePipelineCommunicator

    ps.initialize_model_parallel()
    grids = {"vision": HyperCommGrid([1, 1], ["tp", "dp"], rank_offset=0), "audio": HyperCommGrid([1, 1], ["tp", "dp"], rank_offset=1),
             "language": HyperCommGrid([1, 1], ["tp", "dp"], rank_offset=2)}
    topo = {"vision": ["language"], "audio": ["language"], "language": []}
    comm = MultiModulePipelineCommunicator(grids, topo, dim_mapping={"b": 0})
    assert comm.total_stages == 2 and comm.is_pp_first_stage == (rank < 2) and comm.is_pp_last_stage == (rank == 2)
    m = _build(7, grids)
    assert (m.language_model is not None) == (rank == 2) and set(m.modality_submodules) == ({"vision"} if rank == 0 else {"audio"} if rank == 1 else set())
    b = _batch()
    shapes = {"vision": (8, H), "audio": (6, H)}
    loss = None
    if rank < 2:
        name = "vision" if rank == 0 else "audio"
        inputs = {"vision": {"clip": {"x": b["images"]}}} if rank == 0 else {"audio": {"whisper": {"mel": b["mel"]}}}
        embs = m(b["ids"], modality_inputs=inputs)
        g = comm.send_forward_recv_backward(embs, shapes)
        embs[name].backward(g[name])
    else:
        got = comm.recv_forward(shapes)
        loss = m(b["ids"], b["pos"], None, labels=b["labels"], modality_embeddings=got).float().mean()
        loss.backward()
        comm.send_backward({k: v.grad for k, v in got.items()})
    dist.barrier()
    return (loss.item() if loss is not None else None), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}


def test_mimo_colocated_and_multigrid_agree():
    (ref_loss, ref_grads), = run_distributed(_colocated, 1)
    res = run_distributed(_distributed, 3)
    assert abs(res[2][0] - ref_loss) < 1e-5
    seen = set()
    for r in range(3):
        for n, g in res[r][1].items():
            assert torch.allclose(g, ref_grads[n], atol=1e-5, rtol=1e-4), (r, n)
            seen.add(n)
    assert seen == set(ref_grads)


def test_align_embeddings_masked_scatter_order():
    from megatron_b200.core.models.mimo import MimoModel

    ids = torch.tensor([[5, 90, 90, 7], [90, 6, 91, 91]])
    text = torch.zeros(4, 2, 3)
    v = torch.arange(9.0).reshape(3, 3) + 1
    a = -(torch.arange(6.0).reshape(2, 3) + 1)
    out = MimoModel.align_embeddings_by_token_positions({"vision": v, "audio": a}, ids, {"vision": 90, "audio": 91}, text_embeddings=text)
    assert out.shape == (4, 2, 3)
    assert torch.equal(out[1, 0], v[0]) and torch.equal(out[2, 0], v[1]) and torch.equal(out[0, 1], v[2])
    assert torch.equal(out[2, 1], a[0]) and torch.equal(out[3, 1], a[1]) and torch.all(out[0, 0] == 0)
