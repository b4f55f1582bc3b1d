"""Model-parallel topology and process-group registry.

Capability parity with the reference's ``megatron/core/parallel_state.py``
(``initialize_model_parallel`` :601, ``RankGenerator`` :465, getters :1697-2470,
``destroy_model_parallel`` :2506) but built differently: the world is an n-D numpy
grid of ranks, every named group is a *projection* of that grid described by a
row in ``_GROUP_TABLE``, and the ~90 accessor functions are generated from the
table instead of being written out by hand.  One 8xB200 NVSwitch box is the
design point, so each group additionally records its rank list; the symmetric
memory runtime (``megatron_b200.parallel.symm``) keys its peer-mapped heaps on
that list.
"""
from __future__ import annotations

import warnings
from datetime import timedelta
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

# ----------------------------------------------------------------------------
# Rank grid
# ----------------------------------------------------------------------------


def generate_masked_orthogonal_rank_groups(
    world_size: int, parallel_size: List[int], mask: List[bool]
) -> List[List[int]]:
    """All groups obtained by varying the masked axes and fixing the others.

    ``parallel_size[0]`` is the fastest-varying axis (adjacent ranks).  Same
    contract as reference ``parallel_state.py:269`` but done with a reshape +
    transpose of ``arange(world_size)``.
    """
    assert int(np.prod(parallel_size)) == world_size, (parallel_size, world_size)
    n = len(parallel_size)
    # numpy is row-major: last axis fastest => reverse the size list
    grid = np.arange(world_size).reshape(list(reversed(parallel_size)))
    ax = lambda i: n - 1 - i  # noqa: E731  order-index -> numpy axis
    masked = [ax(i) for i in range(n) if mask[i]]
    unmasked = [ax(i) for i in range(n) if not mask[i]]
    # keep numpy-axis order (slow→fast) inside each half so that within a group
    # the fastest-varying order axis stays fastest
    perm = sorted(unmasked) + sorted(masked)
    gsize = int(np.prod([grid.shape[a] for a in masked])) if masked else 1
    out = grid.transpose(perm).reshape(-1, gsize)
    # Reference enumerates groups with the fastest unmasked axis first; order
    # of *groups* only matters for determinism of new_group calls, keep ours
    # sorted by first member for stable, rank-independent ordering.
    groups = [list(map(int, row)) for row in out]
    groups.sort(key=lambda g: g[0])
    return groups


class RankGenerator:
    """Maps an order string such as ``"tp-cp-ep-dp-pp"`` to rank groups.

    Parity: reference ``parallel_state.py:465-557``.
    """

    def __init__(self, tp: int, ep: int, dp: int, pp: int, cp: int, order: str, rank_offset: int = 0):
        assert ep == 1 or cp == 1, "ep and cp live in different generators"
        self.tp, self.ep, self.dp, self.pp, self.cp = tp, ep, dp, pp, cp
        self.rank_offset = rank_offset
        self.world_size = tp * dp * pp * cp * ep
        self.name_to_size = {"tp": tp, "pp": pp, "dp": dp, "ep": ep, "cp": cp}
        order = order.lower()
        for name, size in self.name_to_size.items():
            if name not in order:
                if size != 1:
                    raise RuntimeError(
                        f"size of '{name}' is {size} but it is missing from order '{order}'"
                    )
                order = order + "-" + name
        self.order = order
        self.ordered_size = [self.name_to_size[t] for t in order.split("-")]

    def get_mask(self, order: str, token: str) -> List[bool]:
        toks = token.split("-")
        return [t in toks for t in order.split("-")]

    def get_ranks(self, token: str) -> List[List[int]]:
        mask = self.get_mask(self.order, token)
        groups = generate_masked_orthogonal_rank_groups(self.world_size, self.ordered_size, mask)
        if self.rank_offset:
            groups = [[r + self.rank_offset for r in g] for g in groups]
        return groups


# ----------------------------------------------------------------------------
# Registry
# ----------------------------------------------------------------------------


class _Group:
    __slots__ = ("name", "pg", "ranks", "gloo", "all_rank_lists")

    def __init__(self, name, pg, ranks, gloo=None, all_rank_lists=None):
        self.name, self.pg, self.ranks, self.gloo = name, pg, ranks, gloo
        self.all_rank_lists = all_rank_lists


_GROUPS: Dict[str, _Group] = {}
_OVERRIDES: Dict[str, Optional[int]] = {}
_VIRTUAL_PP_RANK: Optional[int] = None
_VIRTUAL_PP_WORLD_SIZE: Optional[int] = None
_GLOBAL_MEMORY_BUFFER = None
_EMBEDDING_GLOBAL_RANKS: Optional[List[int]] = None
_POSITION_EMBEDDING_GLOBAL_RANKS: Optional[List[int]] = None
_PIPELINE_GLOBAL_RANKS: Optional[List[int]] = None
_HIERARCHICAL_CP_GROUPS: List = []
_HYBRID_DP_CP_GROUPS: Dict[int, object] = {}
_ALL_GATHER_GROUPS: Dict[str, object] = {}
_INITIALIZED = False
_TOPOLOGY: Dict[str, int] = {}

# name -> (generator: "dense"|"expert", token, make_gloo)
_GROUP_TABLE = {
    "tp": ("dense", "tp", False),
    "pp": ("dense", "pp", False),
    "cp": ("dense", "cp", False),
    "dp": ("dense", "dp", True),
    "dp_cp": ("dense", "dp-cp", True),
    "mp": ("dense", "tp-pp", False),
    "tp_dp_cp": ("dense", "tp-dp-cp", False),
    "tp_dp": ("dense", "tp-dp", False),
    "tp_cp": ("dense", "tp-cp", False),
    "ep": ("expert", "ep", False),
    "expt_tp": ("expert", "tp", False),
    "tp_ep": ("expert", "tp-ep", False),
    "tp_ep_pp": ("expert", "tp-ep-pp", False),
    "expt_dp": ("expert", "dp", True),
}


def _new_group(ranks, backend=None, timeout=None, desc=None, pg_options=None):
    kwargs = {}
    if timeout is not None:
        kwargs["timeout"] = timeout
    if backend is not None:
        kwargs["backend"] = backend
    if pg_options is not None:
        kwargs["pg_options"] = pg_options
    try:
        return dist.new_group(ranks, group_desc=desc, **kwargs)
    except TypeError:
        return dist.new_group(ranks, **kwargs)


_NCCL_OPTION_FIELDS = ("cga_cluster_size", "max_ctas", "min_ctas", "net_name")


def load_nccl_communicator_config(path: Optional[str]) -> dict:
    """YAML ``{group name: {cga_cluster_size, max_ctas, min_ctas, net_name, is_high_priority_stream}}`` (reference :168-200, :713-722).
    On an NVSwitch box the useful knobs are ``max_ctas`` (how many SMs a collective may take away from overlapped GEMMs) and the
    stream priority; unknown group names or fields are rejected so a typo does not silently tune nothing."""
    if not path:
        return {}
    import yaml

    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    known = set(_GROUP_TABLE) | {"embd", "pos_embd", "intra_dp_cp", "inter_dist_opt", "intra_dist_opt", "hcp", "default"}
    for name, fields in cfg.items():
        if name not in known:
            raise ValueError(f"{path}: unknown process group '{name}' (known: {sorted(known)})")
        bad = set(fields) - set(_NCCL_OPTION_FIELDS) - {"is_high_priority_stream"}
        if bad:
            raise ValueError(f"{path}: unknown NCCL option(s) {sorted(bad)} for group '{name}'")
        if str(fields.get("net_name", "ib")).lower() not in ("ib", "socket"):
            raise RuntimeError(f"net_name ({fields['net_name']}) is not supported; accepted values: 'IB' or 'socket'")
    return cfg


def get_nccl_options(pg_name: str, nccl_comm_cfgs: dict, high_priority: bool = False):
    """``ProcessGroupNCCL.Options`` for one group, or ``None`` to take NCCL's defaults."""
    fields = dict(nccl_comm_cfgs.get("default", {}))
    fields.update(nccl_comm_cfgs.get(pg_name, {}))
    if not fields and not high_priority:
        return None
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=bool(fields.get("is_high_priority_stream", high_priority)))
    for k in _NCCL_OPTION_FIELDS:
        if k in fields:
            setattr(opts.config, k, fields[k])
    return opts


def create_group(ranks=None, timeout=None, backend=None, pg_options=None, use_local_synchronization=False, group_desc=None):
    """Thin wrapper kept for API parity (reference ``parallel_state.py:232``)."""
    return _new_group(ranks, backend=backend, timeout=timeout, desc=group_desc, pg_options=pg_options)


def default_embedding_ranks(pp_ranks: Sequence[int]) -> List[int]:
    """First and last pipeline stage hold (tied) embeddings."""
    return [pp_ranks[0]] if len(pp_ranks) == 1 else [pp_ranks[0], pp_ranks[-1]]


def default_position_embedding_ranks(pp_ranks: Sequence[int]) -> List[int]:
    return [pp_ranks[0]]


def initialize_model_parallel(
    tensor_model_parallel_size: int = 1,
    pipeline_model_parallel_size: int = 1,
    virtual_pipeline_model_parallel_size: Optional[int] = None,
    pipeline_model_parallel_comm_backend: Optional[str] = None,
    use_sharp: bool = False,
    context_parallel_size: int = 1,
    hierarchical_context_parallel_sizes: Optional[List[int]] = None,
    hybrid_context_parallel: bool = False,
    expert_model_parallel_size: int = 1,
    gtp_remat_size: int = 1,
    expert_gtp_remat_size: int = 1,
    num_distributed_optimizer_instances: int = 1,
    expert_tensor_parallel_size: Optional[int] = None,
    nccl_communicator_config_path: Optional[str] = None,
    distributed_timeout_minutes: int = 30,
    order: str = "tp-cp-ep-dp-pp",
    get_embedding_ranks: Optional[Callable] = None,
    get_position_embedding_ranks: Optional[Callable] = None,
    create_gloo_process_groups: bool = True,
    high_priority_stream_groups: Optional[List[str]] = None,
    sharp_enabled_group: Optional[str] = None,
    rank_offset: int = 0,
    local_world_size: Optional[int] = None,
) -> None:
    """Create every model/data-parallel group (signature parity: reference :601-625)."""
    global _INITIALIZED, _VIRTUAL_PP_RANK, _VIRTUAL_PP_WORLD_SIZE
    global _EMBEDDING_GLOBAL_RANKS, _POSITION_EMBEDDING_GLOBAL_RANKS, _PIPELINE_GLOBAL_RANKS
    assert dist.is_initialized(), "torch.distributed must be initialised first"
    assert not _INITIALIZED, "model parallel groups are already initialised"
    if use_sharp or sharp_enabled_group:
        warnings.warn("SHARP is an InfiniBand feature; ignored on a single NVSwitch box")

    world = local_world_size or dist.get_world_size()
    rank = dist.get_rank()
    tp, pp, cp, ep = (
        tensor_model_parallel_size,
        pipeline_model_parallel_size,
        context_parallel_size,
        expert_model_parallel_size,
    )
    model_size = tp * pp * cp
    if world % model_size != 0:
        raise RuntimeError(f"world_size ({world}) is not divisible by tp*pp*cp ({model_size})")
    dp = world // model_size
    etp = tp if expert_tensor_parallel_size is None else expert_tensor_parallel_size
    expert_model_size = etp * ep * pp
    if world % expert_model_size != 0:
        raise RuntimeError(f"world_size ({world}) is not divisible by etp*ep*pp ({expert_model_size})")
    edp = world // expert_model_size

    if virtual_pipeline_model_parallel_size is not None:
        if pp < 2:
            raise RuntimeError("virtual pipeline requires pipeline_model_parallel_size >= 2")
        _VIRTUAL_PP_RANK = 0
        _VIRTUAL_PP_WORLD_SIZE = virtual_pipeline_model_parallel_size

    dense = RankGenerator(tp=tp, ep=1, dp=dp, pp=pp, cp=cp, order=order, rank_offset=rank_offset)
    expert = RankGenerator(tp=etp, ep=ep, dp=edp, pp=pp, cp=1, order=order, rank_offset=rank_offset)
    assert dense.get_ranks("pp") == expert.get_ranks("pp"), "dense and expert pipelines must coincide"
    gens = {"dense": dense, "expert": expert}
    timeout = timedelta(minutes=distributed_timeout_minutes)
    backend_is_cpu = dist.get_backend() == "gloo"

    _TOPOLOGY.update(dict(tp=tp, pp=pp, cp=cp, ep=ep, dp=dp, etp=etp, edp=edp, world=world))

    nccl_cfgs = load_nccl_communicator_config(nccl_communicator_config_path)
    hp_groups = set(high_priority_stream_groups or [])

    def register(name, rank_lists, gloo=False, backend=None):
        mine = None
        opts = None
        if not backend_is_cpu and (backend in (None, "nccl")):
            opts = get_nccl_options(name, nccl_cfgs, high_priority=name in hp_groups)
        for ranks in rank_lists:
            pg = _new_group(ranks, backend=backend, timeout=timeout, desc=name.upper(), pg_options=opts)
            gpg = None
            if gloo and create_gloo_process_groups:
                gpg = pg if backend_is_cpu else _new_group(ranks, backend="gloo", timeout=timeout, desc=name.upper() + "_GLOO")
            if rank in ranks:
                mine = _Group(name, pg, list(ranks), gpg, rank_lists)
        if mine is not None:
            _GROUPS[name] = mine

    for name, (gen, token, gloo) in _GROUP_TABLE.items():
        be = pipeline_model_parallel_comm_backend if name == "pp" else None
        register(name, gens[gen].get_ranks(token), gloo=gloo, backend=be)

    # pipeline-derived groups
    emb_fn = get_embedding_ranks or default_embedding_ranks
    pos_fn = get_position_embedding_ranks or default_position_embedding_ranks
    pp_lists = dense.get_ranks("pp")
    register("embd", [emb_fn(r) for r in pp_lists])
    register("pos_embd", [pos_fn(r) for r in pp_lists])
    for r in pp_lists:
        if rank in r:
            _PIPELINE_GLOBAL_RANKS = list(r)
            _EMBEDDING_GLOBAL_RANKS = emb_fn(r) if rank in emb_fn(r) else None
            _POSITION_EMBEDDING_GLOBAL_RANKS = pos_fn(r) if rank in pos_fn(r) else None

    # distributed-optimizer instances: split dp_cp into `n` contiguous replicas
    n_inst = num_distributed_optimizer_instances
    dpcp_lists = dense.get_ranks("dp-cp")
    assert (dp * cp) % n_inst == 0
    intra_size = (dp * cp) // n_inst
    intra, inter = [], []
    for lst in dpcp_lists:
        for i in range(n_inst):
            intra.append(lst[i * intra_size : (i + 1) * intra_size])
        for j in range(intra_size):
            inter.append(lst[j::intra_size])
    register("intra_dp_cp", intra, gloo=True)
    if n_inst > 1:
        register("inter_dist_opt", inter)

    # generalised TP weight rematerialisation: R adjacent ranks of every data-parallel group share one copy of each GTP weight (sharded along out-features);
    # the ranks holding the SAME shard form the orthogonal group over which that shard's gradient is data-parallel-reduced
    if gtp_remat_size > 1:
        assert (dp * cp) % gtp_remat_size == 0, f"gtp_remat_size ({gtp_remat_size}) must divide dp*cp ({dp * cp})"
        remat, ortho = [], []
        for lst in dense.get_ranks("dp-cp"):
            for i in range(0, len(lst), gtp_remat_size):
                remat.append(lst[i : i + gtp_remat_size])
            for j in range(gtp_remat_size):
                ortho.append(lst[j::gtp_remat_size])
        register("gtp_remat", remat)
        register("dp_no_gtp", ortho)
    _TOPOLOGY["gtp"] = gtp_remat_size
    # expert-side GTP: the same construction over the EXPERT data-parallel groups (expert weights have their own dp axis under EP)
    if expert_gtp_remat_size > 1:
        assert edp % expert_gtp_remat_size == 0, f"expert_gtp_remat_size ({expert_gtp_remat_size}) must divide the expert data-parallel size ({edp})"
        remat, ortho = [], []
        for lst in expert.get_ranks("dp"):
            for i in range(0, len(lst), expert_gtp_remat_size):
                remat.append(lst[i : i + expert_gtp_remat_size])
            for j in range(expert_gtp_remat_size):
                ortho.append(lst[j::expert_gtp_remat_size])
        register("egtp_remat", remat)
        register("expt_dp_no_egtp", ortho)
    _TOPOLOGY["egtp"] = expert_gtp_remat_size

    # hierarchical context parallel (a2a inside NVLink island, ring across)
    del _HIERARCHICAL_CP_GROUPS[:]
    if hierarchical_context_parallel_sizes:
        assert int(np.prod(hierarchical_context_parallel_sizes)) == cp
        for cp_ranks in dense.get_ranks("cp"):
            arr = np.array(cp_ranks).reshape(list(reversed(hierarchical_context_parallel_sizes)))
            nlev = len(hierarchical_context_parallel_sizes)
            for lev in range(nlev):
                axis = nlev - 1 - lev
                moved = np.moveaxis(arr, axis, -1).reshape(-1, arr.shape[axis])
                for sub in moved:
                    pg = _new_group(list(map(int, sub)), timeout=timeout, desc=f"HIERARCHICAL_CONTEXT_PARALLEL_GROUP_L{lev}")
                    if rank in sub:
                        while len(_HIERARCHICAL_CP_GROUPS) <= lev:
                            _HIERARCHICAL_CP_GROUPS.append(None)
                        _HIERARCHICAL_CP_GROUPS[lev] = pg

    _HYBRID_DP_CP_GROUPS.clear()
    _ALL_GATHER_GROUPS.clear()
    if hybrid_context_parallel:
        for lst in dense.get_ranks("dp-cp"):
            create_hybrid_dp_cp_groups(rank, [int(r) for r in lst], timeout=timeout)

    _INITIALIZED = True


def model_parallel_is_initialized() -> bool:
    return _INITIALIZED and "tp" in _GROUPS and "pp" in _GROUPS and "dp" in _GROUPS


def is_initialized() -> bool:
    return _INITIALIZED


def is_unitialized() -> bool:  # sic — the reference keeps this spelling
    return not _INITIALIZED


def destroy_model_parallel() -> None:
    """Forget every group (reference :2506). Groups themselves are destroyed lazily."""
    global _INITIALIZED, _VIRTUAL_PP_RANK, _VIRTUAL_PP_WORLD_SIZE, _GLOBAL_MEMORY_BUFFER
    global _EMBEDDING_GLOBAL_RANKS, _POSITION_EMBEDDING_GLOBAL_RANKS, _PIPELINE_GLOBAL_RANKS
    seen = set()
    for g in _GROUPS.values():
        for pg in (g.pg, g.gloo):
            if pg is not None and id(pg) not in seen and pg is not dist.group.WORLD:
                seen.add(id(pg))
                try:
                    dist.destroy_process_group(pg)
                except Exception:
                    pass
    _GROUPS.clear()
    _OVERRIDES.clear()
    _TOPOLOGY.clear()
    del _HIERARCHICAL_CP_GROUPS[:]
    _HYBRID_DP_CP_GROUPS.clear()
    _ALL_GATHER_GROUPS.clear()
    _INITIALIZED = False
    _VIRTUAL_PP_RANK = _VIRTUAL_PP_WORLD_SIZE = None
    _GLOBAL_MEMORY_BUFFER = None
    _EMBEDDING_GLOBAL_RANKS = _POSITION_EMBEDDING_GLOBAL_RANKS = _PIPELINE_GLOBAL_RANKS = None


# ----------------------------------------------------------------------------
# Accessors
# ----------------------------------------------------------------------------


def _grp(name: str, check: bool = True):
    g = _GROUPS.get(name)
    if g is None and check:
        raise AssertionError(f"process group '{name}' is not initialised")
    return g


def get_group(name: str, check_initialized: bool = True):
    g = _grp(name, check_initialized)
    return None if g is None else g.pg


def get_group_ranks(name: str) -> List[int]:
    return list(_grp(name).ranks)


def _ws(name: str) -> int:
    ov = _OVERRIDES.get(name + ".ws")
    if ov is not None:
        return ov
    g = _GROUPS.get(name)
    if g is None:
        return 1 if not _INITIALIZED else dist.get_world_size(group=_grp(name).pg)
    return len(g.ranks)


def _rk(name: str) -> int:
    ov = _OVERRIDES.get(name + ".rk")
    if ov is not None:
        return ov
    g = _GROUPS.get(name)
    if g is None:
        return 0
    return g.ranks.index(dist.get_rank())


# ---- accessors of the registry-backed groups (reference :1697-2470).  Written out one by one — they are the public API every
# model file imports by name — but each is a one-liner over the group registry (``get_group`` / ``_ws`` / ``_rk``).
def get_tensor_model_parallel_group(check_initialized: bool = True):
    """The tensor-model-parallel process group this rank belongs to."""
    return get_group("tp", check_initialized)


def get_tensor_model_parallel_world_size() -> int:
    return _ws("tp")


def get_tensor_model_parallel_rank() -> int:
    return _rk("tp")


def set_tensor_model_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["tp.ws"] = world_size


def set_tensor_model_parallel_rank(rank) -> None:
    _OVERRIDES["tp.rk"] = rank


def get_pipeline_model_parallel_group(check_initialized: bool = True):
    """The pipeline-model-parallel process group this rank belongs to."""
    return get_group("pp", check_initialized)


def get_pipeline_model_parallel_world_size() -> int:
    return _ws("pp")


def get_pipeline_model_parallel_rank() -> int:
    return _rk("pp")


def set_pipeline_model_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["pp.ws"] = world_size


def set_pipeline_model_parallel_rank(rank) -> None:
    _OVERRIDES["pp.rk"] = rank


def get_context_parallel_group(check_initialized: bool = True):
    """The context-parallel process group this rank belongs to."""
    return get_group("cp", check_initialized)


def get_context_parallel_world_size() -> int:
    return _ws("cp")


def get_context_parallel_rank() -> int:
    return _rk("cp")


def set_context_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["cp.ws"] = world_size


def set_context_parallel_rank(rank) -> None:
    _OVERRIDES["cp.rk"] = rank


def get_model_parallel_group(check_initialized: bool = True):
    """The model-parallel (tp x pp) process group this rank belongs to."""
    return get_group("mp", check_initialized)


def get_model_parallel_world_size() -> int:
    return _ws("mp")


def get_model_parallel_rank() -> int:
    return _rk("mp")


def set_model_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["mp.ws"] = world_size


def set_model_parallel_rank(rank) -> None:
    _OVERRIDES["mp.rk"] = rank


def get_tensor_and_context_parallel_group(check_initialized: bool = True):
    """The tensor-and-context-parallel process group this rank belongs to."""
    return get_group("tp_cp", check_initialized)


def get_tensor_and_context_parallel_world_size() -> int:
    return _ws("tp_cp")


def get_tensor_and_context_parallel_rank() -> int:
    return _rk("tp_cp")


def set_tensor_and_context_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["tp_cp.ws"] = world_size


def set_tensor_and_context_parallel_rank(rank) -> None:
    _OVERRIDES["tp_cp.rk"] = rank


def get_expert_model_parallel_group(check_initialized: bool = True):
    """The expert-model-parallel process group this rank belongs to."""
    return get_group("ep", check_initialized)


def get_expert_model_parallel_world_size() -> int:
    return _ws("ep")


def get_expert_model_parallel_rank() -> int:
    return _rk("ep")


def set_expert_model_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["ep.ws"] = world_size


def set_expert_model_parallel_rank(rank) -> None:
    _OVERRIDES["ep.rk"] = rank


def get_expert_tensor_parallel_group(check_initialized: bool = True):
    """The expert-tensor-parallel process group this rank belongs to."""
    return get_group("expt_tp", check_initialized)


def get_expert_tensor_parallel_world_size() -> int:
    return _ws("expt_tp")


def get_expert_tensor_parallel_rank() -> int:
    return _rk("expt_tp")


def set_expert_tensor_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["expt_tp.ws"] = world_size


def set_expert_tensor_parallel_rank(rank) -> None:
    _OVERRIDES["expt_tp.rk"] = rank


def get_expert_tensor_and_model_parallel_group(check_initialized: bool = True):
    """The expert tensor x expert model parallel process group this rank belongs to."""
    return get_group("tp_ep", check_initialized)


def get_expert_tensor_and_model_parallel_world_size() -> int:
    return _ws("tp_ep")


def get_expert_tensor_and_model_parallel_rank() -> int:
    return _rk("tp_ep")


def set_expert_tensor_and_model_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["tp_ep.ws"] = world_size


def set_expert_tensor_and_model_parallel_rank(rank) -> None:
    _OVERRIDES["tp_ep.rk"] = rank


def get_expert_tensor_model_pipeline_parallel_group(check_initialized: bool = True):
    """The expert tensor x model x pipeline parallel process group this rank belongs to."""
    return get_group("tp_ep_pp", check_initialized)


def get_expert_tensor_model_pipeline_parallel_world_size() -> int:
    return _ws("tp_ep_pp")


def get_expert_tensor_model_pipeline_parallel_rank() -> int:
    return _rk("tp_ep_pp")


def set_expert_tensor_model_pipeline_parallel_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["tp_ep_pp.ws"] = world_size


def set_expert_tensor_model_pipeline_parallel_rank(rank) -> None:
    _OVERRIDES["tp_ep_pp.rk"] = rank


def get_embedding_group(check_initialized: bool = True):
    """The embedding (first + last pipeline stage) process group this rank belongs to."""
    return get_group("embd", check_initialized)


def get_embedding_world_size() -> int:
    return _ws("embd")


def get_embedding_rank() -> int:
    return _rk("embd")


def set_embedding_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["embd.ws"] = world_size


def set_embedding_rank(rank) -> None:
    _OVERRIDES["embd.rk"] = rank


def get_position_embedding_group(check_initialized: bool = True):
    """The position-embedding process group this rank belongs to."""
    return get_group("pos_embd", check_initialized)


def get_position_embedding_world_size() -> int:
    return _ws("pos_embd")


def get_position_embedding_rank() -> int:
    return _rk("pos_embd")


def set_position_embedding_world_size(world_size) -> None:
    """Override (tests, checkpoint conversion tools): ``None`` restores the real value."""
    _OVERRIDES["pos_embd.ws"] = world_size


def set_position_embedding_rank(rank) -> None:
    _OVERRIDES["pos_embd.rk"] = rank


def get_data_parallel_group(with_context_parallel: bool = False, partial_data_parallel: bool = False):
    if partial_data_parallel:
        return get_group("intra_dp_cp")
    return get_group("dp_cp" if with_context_parallel else "dp")


def get_data_parallel_group_gloo(with_context_parallel: bool = False, partial_data_parallel: bool = False):
    name = "intra_dp_cp" if partial_data_parallel else ("dp_cp" if with_context_parallel else "dp")
    return _grp(name).gloo


def get_data_parallel_world_size(with_context_parallel: bool = False, partial_data_parallel: bool = False) -> int:
    if not _INITIALIZED:
        return 1
    name = "intra_dp_cp" if partial_data_parallel else ("dp_cp" if with_context_parallel else "dp")
    return _ws(name)


def get_data_parallel_rank(with_context_parallel: bool = False, partial_data_parallel: bool = False) -> int:
    if not _INITIALIZED:
        return 0
    name = "intra_dp_cp" if partial_data_parallel else ("dp_cp" if with_context_parallel else "dp")
    return _rk(name)


def get_expert_data_parallel_group(check_initialized: bool = True, partial_expert_data_parallel: bool = False):
    return get_group("expt_dp", check_initialized)


def get_expert_data_parallel_group_gloo(partial_expert_data_parallel: bool = False):
    return _grp("expt_dp").gloo


def get_expert_data_parallel_rank(partial_expert_data_parallel: bool = False) -> int:
    return _rk("expt_dp")


def get_expert_data_parallel_world_size(partial_expert_data_parallel: bool = False) -> int:
    return _ws("expt_dp")


def get_tensor_and_data_parallel_group(with_context_parallel: bool = False):
    return get_group("tp_dp_cp" if with_context_parallel else "tp_dp")


def get_tensor_and_data_parallel_world_size(with_context_parallel: bool = False) -> int:
    return _ws("tp_dp_cp" if with_context_parallel else "tp_dp")


def get_tensor_and_data_parallel_rank(with_context_parallel: bool = False) -> int:
    return _rk("tp_dp_cp" if with_context_parallel else "tp_dp")


def get_amax_reduction_group(with_context_parallel: bool = False, tp_only_amax_red: bool = False):
    """FP8 amax statistics are reduced over tp(-dp-cp) (reference :1877)."""
    if tp_only_amax_red:
        return get_group("tp_cp" if with_context_parallel else "tp")
    return get_tensor_and_data_parallel_group(with_context_parallel)


def get_intra_distributed_optimizer_instance_group():
    return get_group("intra_dp_cp")


def get_inter_distributed_optimizer_instance_group():
    return get_group("inter_dist_opt", check_initialized=False)


def get_gtp_weight_remat_group(check_initialized: bool = True):
    """Ranks that jointly hold one copy of every GTP-sharded weight (``None`` when ``gtp_remat_size == 1``)."""
    return get_group("gtp_remat", check_initialized=False)


def get_gtp_weight_remat_world_size() -> int:
    return _ws("gtp_remat") if "gtp_remat" in _GROUPS else 1


def get_gtp_weight_remat_rank() -> int:
    return _rk("gtp_remat") if "gtp_remat" in _GROUPS else 0


def get_data_parallel_group_without_gtp(check_initialized: bool = True):
    """Data-parallel replicas of ONE GTP shard (the full dp-cp group when GTP is off)."""
    g = get_group("dp_no_gtp", check_initialized=False)
    return g if g is not None else get_data_parallel_group(with_context_parallel=True)


def get_expert_gtp_weight_remat_group(check_initialized: bool = True):
    """Ranks that jointly hold one copy of every GTP-sharded EXPERT weight (``None`` when ``expert_gtp_remat_size == 1``)."""
    return get_group("egtp_remat", check_initialized=False)


def get_expert_gtp_weight_remat_world_size() -> int:
    return _ws("egtp_remat") if "egtp_remat" in _GROUPS else 1


def get_expert_data_parallel_group_without_gtp(check_initialized: bool = True):
    g = get_group("expt_dp_no_egtp", check_initialized=False)
    return g if g is not None else get_group("expt_dp", check_initialized=False)


def get_hierarchical_context_parallel_groups(check_initialized: bool = True):
    if check_initialized:
        assert _HIERARCHICAL_CP_GROUPS, "hierarchical context parallel groups are not initialised"
    return list(_HIERARCHICAL_CP_GROUPS)


def get_context_parallel_global_ranks():
    return get_group_ranks("cp")


def get_tensor_model_parallel_src_rank() -> int:
    """Global rank of the first member of this rank's TP group."""
    g = _GROUPS.get("tp")
    return dist.get_rank() if g is None else g.ranks[0]


def get_model_parallel_src_rank() -> int:
    return _grp("mp").ranks[0]


def get_data_parallel_src_rank(with_context_parallel: bool = False) -> int:
    return _grp("dp_cp" if with_context_parallel else "dp").ranks[0]


def get_expert_model_parallel_src_rank() -> int:
    return _grp("ep").ranks[0]


# ---- pipeline helpers --------------------------------------------------------


def get_pipeline_model_parallel_first_rank() -> int:
    return _PIPELINE_GLOBAL_RANKS[0]


def get_pipeline_model_parallel_last_rank() -> int:
    return _PIPELINE_GLOBAL_RANKS[-1]


def get_pipeline_model_parallel_next_rank() -> int:
    r = _rk("pp")
    return _PIPELINE_GLOBAL_RANKS[(r + 1) % len(_PIPELINE_GLOBAL_RANKS)]


def get_pipeline_model_parallel_prev_rank() -> int:
    r = _rk("pp")
    return _PIPELINE_GLOBAL_RANKS[(r - 1) % len(_PIPELINE_GLOBAL_RANKS)]


def get_virtual_pipeline_model_parallel_rank() -> Optional[int]:
    return _VIRTUAL_PP_RANK


def set_virtual_pipeline_model_parallel_rank(rank: Optional[int]) -> None:
    global _VIRTUAL_PP_RANK
    _VIRTUAL_PP_RANK = rank


def get_virtual_pipeline_model_parallel_world_size() -> Optional[int]:
    return _VIRTUAL_PP_WORLD_SIZE


def set_virtual_pipeline_model_parallel_world_size(ws: Optional[int]) -> None:
    global _VIRTUAL_PP_WORLD_SIZE
    _VIRTUAL_PP_WORLD_SIZE = ws


def is_pipeline_first_stage(ignore_virtual: bool = True, vp_stage: Optional[int] = None) -> bool:
    if not ignore_virtual and _VIRTUAL_PP_WORLD_SIZE is not None:
        stage = _VIRTUAL_PP_RANK if vp_stage is None else vp_stage
        if stage != 0:
            return False
    return _rk("pp") == 0


def is_pipeline_last_stage(ignore_virtual: bool = True, vp_stage: Optional[int] = None) -> bool:
    if not ignore_virtual and _VIRTUAL_PP_WORLD_SIZE is not None:
        stage = _VIRTUAL_PP_RANK if vp_stage is None else vp_stage
        if stage != _VIRTUAL_PP_WORLD_SIZE - 1:
            return False
    return _rk("pp") == _ws("pp") - 1


def is_rank_in_embedding_group(ignore_virtual: bool = True, vp_stage: Optional[int] = None) -> bool:
    if _EMBEDDING_GLOBAL_RANKS is None:
        return False
    rank = dist.get_rank()
    if ignore_virtual or _VIRTUAL_PP_WORLD_SIZE is None:
        return rank in _EMBEDDING_GLOBAL_RANKS
    if rank == _EMBEDDING_GLOBAL_RANKS[0]:
        return is_pipeline_first_stage(ignore_virtual=False, vp_stage=vp_stage)
    if rank == _EMBEDDING_GLOBAL_RANKS[-1]:
        return is_pipeline_last_stage(ignore_virtual=False, vp_stage=vp_stage)
    return True


def is_rank_in_position_embedding_group() -> bool:
    return _POSITION_EMBEDDING_GLOBAL_RANKS is not None and dist.get_rank() in _POSITION_EMBEDDING_GLOBAL_RANKS


def is_inside_encoder(rank=None) -> bool:
    return False


def is_inside_decoder(rank=None) -> bool:
    return True


# ---- memory buffer -----------------------------------------------------------


def _set_global_memory_buffer():
    global _GLOBAL_MEMORY_BUFFER
    from .utils import GlobalMemoryBuffer

    _GLOBAL_MEMORY_BUFFER = GlobalMemoryBuffer()


def get_global_memory_buffer():
    if _GLOBAL_MEMORY_BUFFER is None:
        _set_global_memory_buffer()
    return _GLOBAL_MEMORY_BUFFER


def destroy_global_memory_buffer():
    global _GLOBAL_MEMORY_BUFFER
    _GLOBAL_MEMORY_BUFFER = None


def get_all_ranks() -> str:
    """``tp_pp_dp_ep_etp`` style rank tag used in log/trace names."""
    return "_".join(str(_rk(n)) for n in ("tp", "cp", "ep", "dp", "pp") if n in _GROUPS)


def update_pg_timeout(timeout: timedelta, pg=None):
    """Best-effort timeout bump for NCCL groups (reference :203-229)."""
    try:
        from torch.distributed.distributed_c10d import _set_pg_timeout

        for g in ([pg] if pg is not None else [x.pg for x in _GROUPS.values()]):
            _set_pg_timeout(timeout, g)
    except Exception:  # pragma: no cover - backend dependent
        pass


def get_topology() -> Dict[str, int]:
    return dict(_TOPOLOGY)


# ---- round-2 additions: group builders usable outside ``initialize_model_parallel`` and the remaining accessors --------------


def create_hierarchical_groups(rank: int, ranks: Sequence[int], hierarchical_group_sizes: Sequence[int], create_gloo_process_groups: bool = False,
                               pg_options=None, timeout=None, group_desc: Optional[str] = None):
    """Split ``ranks`` into nested levels: level 0 = the innermost ``sizes[0]`` consecutive ranks (one NVLink domain), level 1 =
    ranks with the same level-0 position across ``sizes[1]`` domains, ...  Every rank must call this (all sub-groups are created
    everywhere, in the same order).  Returns (this rank's groups per level, their gloo twins or ``None``)
    (reference ``parallel_state.py:389``)."""
    assert int(np.prod(hierarchical_group_sizes)) == len(ranks), f"{list(hierarchical_group_sizes)} does not factor {len(ranks)} ranks"
    arr = np.array(list(ranks)).reshape(list(reversed(list(hierarchical_group_sizes))))
    nlev = len(hierarchical_group_sizes)
    mine, mine_gloo = [None] * nlev, [None] * nlev
    for lev in range(nlev):
        axis = nlev - 1 - lev
        for sub in np.moveaxis(arr, axis, -1).reshape(-1, arr.shape[axis]):
            sub = [int(r) for r in sub]
            pg = _new_group(sub, timeout=timeout, desc=f"{group_desc or 'HIERARCHICAL_GROUP'}_L{lev}", pg_options=pg_options)
            gl = _new_group(sub, backend="gloo", timeout=timeout, desc=f"{group_desc or 'HIERARCHICAL_GROUP'}_L{lev}_GLOO") if create_gloo_process_groups else None
            if rank in sub:
                mine[lev], mine_gloo[lev] = pg, gl
    return mine, (mine_gloo if create_gloo_process_groups else None)


def create_hybrid_dp_cp_groups(rank: int, ranks: Sequence[int], pg_options=None, timeout=None) -> Dict[int, object]:
    """Hybrid context parallelism (variable CP size per sample): one group per power-of-two size 2, 4, ... < len(ranks) over
    consecutive DP x CP ranks; a long sample borrows as many neighbours as it needs (reference ``parallel_state.py:440``).
    The full-size group is the ordinary dp_cp group and is not duplicated."""
    out: Dict[int, object] = {}
    n = len(ranks)
    size = 2
    while size < n:
        for i in range(0, n, size):
            sub = [int(r) for r in ranks[i:i + size]]
            pg = _new_group(sub, timeout=timeout, desc=f"HYBRID_DP_CP_GROUP_{size}", pg_options=pg_options)
            if rank in sub:
                assert size not in out, f"rank {rank} appears in two hybrid DPxCP groups of size {size}"
                out[size] = pg
        size *= 2
    _HYBRID_DP_CP_GROUPS.update(out)
    return out


def get_hybrid_data_context_parallel_groups(check_initialized: bool = True, group_size: Optional[int] = None):
    """The power-of-two DP x CP sub-group of ``group_size`` this rank belongs to (``None`` size: the whole dictionary); the
    full size resolves to the ordinary data x context parallel group."""
    if group_size is None:
        if check_initialized:
            assert _HYBRID_DP_CP_GROUPS, "hybrid DPxCP groups are not initialised (create_hybrid_dp_cp_groups)"
        return dict(_HYBRID_DP_CP_GROUPS)
    if group_size == get_data_parallel_world_size(with_context_parallel=True):
        return get_data_parallel_group(with_context_parallel=True)
    if check_initialized:
        assert group_size in _HYBRID_DP_CP_GROUPS, f"no hybrid DPxCP group of size {group_size}"
    return _HYBRID_DP_CP_GROUPS.get(group_size)


def create_all_gather_groups(names: Sequence[str] = ("dp", "dp_cp", "expt_dp"), pg_options=None, timeout=None) -> Dict[str, object]:
    """Second communicators over the same ranks as the data-parallel groups, used ONLY for parameter all-gathers, so that the
    next step's weight gathers and this step's gradient reduce-scatters do not serialise on one NCCL stream
    (reference ``parallel_state.py:create_all_gather_groups``; with the NVLink backend the same separation is two slots of
    the symmetric heap)."""
    rank = dist.get_rank()
    for name in names:
        g = _GROUPS.get(name)
        if g is None or name in _ALL_GATHER_GROUPS:
            continue
        mine = None
        for ranks in (g.all_rank_lists or [g.ranks]):         # every rank creates every replica's communicator, in the same order
            pg = _new_group(list(ranks), timeout=timeout, desc=f"{name.upper()}_ALL_GATHER_GROUP", pg_options=pg_options)
            if rank in ranks:
                mine = pg
        _ALL_GATHER_GROUPS[name] = mine
    return dict(_ALL_GATHER_GROUPS)


def get_all_gather_group(name: str = "dp_cp"):
    """The dedicated all-gather communicator of a data-parallel group, or the group itself when none was created."""
    return _ALL_GATHER_GROUPS.get(name) or get_group(name, check_initialized=False)


def overwrite_nccl_comm_cfgs(nccl_comm_cfgs: dict, pg_name: str, key_value_pair: tuple) -> None:
    """Set one NCCL option of one group in a configuration dictionary (as read by ``load_nccl_communicator_config``)."""
    k, v = key_value_pair
    nccl_comm_cfgs.setdefault(pg_name, {})[k] = v


def set_data_parallel_rank(rank) -> None:
    _OVERRIDES["dp.rk"] = rank


def get_gtp_weight_remat_global_ranks() -> List[int]:
    return list(_GROUPS["gtp_remat"].ranks) if "gtp_remat" in _GROUPS else [dist.get_rank() if dist.is_initialized() else 0]


def get_expert_gtp_weight_remat_rank() -> int:
    return _rk("egtp_remat") if "egtp_remat" in _GROUPS else 0


def get_expert_gtp_weight_remat_global_ranks() -> List[int]:
    return list(_GROUPS["egtp_remat"].ranks) if "egtp_remat" in _GROUPS else [dist.get_rank() if dist.is_initialized() else 0]
