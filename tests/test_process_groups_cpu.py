"""HyperCommGrid (views, shared dims, rank enumeration vs the RankGenerator) and ProcessGroupCollection helpers (from_grid, optimizer / DDP group
bundles, multi-module collection) — reference tests/unit_tests/test_hyper_comm_grid.py, test_process_groups_config.py."""
import os
from types import SimpleNamespace

import pytest
import torch

from dist_utils import run_distributed


def test_grid_enumeration_matches_rank_generator(monkeypatch):
    from megatron_b200.core.hyper_comm_grid import HyperCommGrid, enumerate_groups
    from megatron_b200.core.parallel_state import RankGenerator

    monkeypatch.setenv("WORLD_SIZE", "120")
    grid = HyperCommGrid([2, 3, 4, 5], ["tp", "cp", "pp", "dp"])
    gen = RankGenerator(tp=2, ep=1, dp=5, pp=4, cp=3, order="tp-cp-pp-dp")
    for dims, token in [("tp", "tp"), ("dp", "dp"), (["cp", "dp"], "dp-cp"), (["tp", "pp"], "tp-pp"), (["tp", "cp", "dp"], "tp-cp-dp")]:
        ours = sorted(sorted(g) for g in grid.get_rank_enum(dims))
        ref = sorted(sorted(g) for g in gen.get_ranks(token))
        assert ours == ref, (dims, ours[:2], ref[:2])
    # members ascend, groups ordered by first member; an offset shifts everything
    assert enumerate_groups([2, 2, 2], ["a", "b", "c"], ["a", "b"]) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert enumerate_groups([2, 2, 2], ["a", "b", "c"], ["c"], rank_offset=8) == [[8, 12], [9, 13], [10, 14], [11, 15]]
    assert grid.coords(2 * 3 * 4 * 3 + 2 * 3 * 1 + 2 * 2 + 1) == {"tp": 1, "cp": 2, "pp": 1, "dp": 3}


def test_grid_validation_and_views(monkeypatch):
    from megatron_b200.core.hyper_comm_grid import HyperCommGrid

    monkeypatch.setenv("WORLD_SIZE", "16")
    with pytest.raises(RuntimeError):
        HyperCommGrid([4, 8], ["tp", "dp"])
    with pytest.raises(RuntimeError):
        HyperCommGrid([2, 4], ["tp", "dp"], rank_offset=12)
    with pytest.raises(ValueError):
        HyperCommGrid([2, 4], ["tp"])
    grid = HyperCommGrid([2, 2, 2, 2], ["tp", "cp", "dp", "pp"])
    # expert factorisation of the same 16 ranks: etp 1 x ep 4 x edp 2 x pp 2; pp must (and does) coincide
    grid.register_view("expert", [1, 4, 2, 2], ["expt_tp", "ep", "expt_dp", "pp"], shared_dims=["pp"])
    assert grid.get_rank_enum("pp", view="expert") == grid.get_rank_enum("pp")
    assert grid.get_rank_enum("ep", view="expert")[0] == [0, 1, 2, 3]
    assert grid.get_rank_enum(["ep", "expt_dp"], view="expert")[1] == list(range(8, 16))
    assert grid._pg_key(grid._view("expert"), ["pp"])[0] == "pp"                         # shared → the base key
    assert grid._pg_key(grid._view("expert"), ["ep"])[0] == ("expert", "ep")
    assert grid._pg_key(grid._view(None), grid._view(None).canonical(["tp", "dp"]))[0] == "dp-tp"
    with pytest.raises(ValueError, match="already registered"):
        grid.register_view("expert", [16], ["x"])
    with pytest.raises(ValueError, match="size"):
        grid.register_view("bad", [3, 4], ["a", "b"])
    with pytest.raises(ValueError, match="different membership"):
        grid.register_view("bad", [2, 2, 2, 2], ["pp", "ep", "expt_dp", "expt_tp"], shared_dims=["pp"])     # pp fastest here, slowest in the base
    with pytest.raises(ValueError, match="not in the base view"):
        grid.register_view("bad", [16], ["ep"], shared_dims=["ep"])
    with pytest.raises(KeyError):
        grid.get_rank_enum("tp", view="nope")
    with pytest.raises(ValueError, match="is not in view"):
        grid.get_rank_enum("tp", view="expert")
    with pytest.raises(KeyError, match="create_pg first"):
        grid.get_pg("tp")


def _grid_worker(rank, world):
    import torch.distributed as dist

    from megatron_b200.core.hyper_comm_grid import HyperCommGrid
    from megatron_b200.core.process_groups_config import MultiModuleProcessGroupCollection, ProcessGroupCollection

    grid = HyperCommGrid([2, 2], ["tp", "dp"])
    grid.register_view("expert", [1, 4], ["expt_tp", "ep"])
    pgc = ProcessGroupCollection.from_grid(grid, expert_view="expert")
    assert pgc.tp.size() == 2 and pgc.dp.size() == 2 and pgc.ep.size() == 4 and pgc.expt_tp.size() == 1 and pgc.pp is None and pgc.cp is None
    assert pgc.dp_cp is pgc.dp and pgc.mp is pgc.tp and pgc.tp_dp_cp.size() == 4
    assert "tp(2)" in repr(pgc) and "ep(4)" in repr(pgc)
    with pytest.raises(KeyError):
        grid.create_pg("tp")                                   # already created by from_grid
    x = torch.tensor([float(rank)])
    dist.all_reduce(x, group=pgc.tp)
    assert x.item() == {0: 1.0, 1: 1.0, 2: 5.0, 3: 5.0}[rank]
    y = torch.tensor([float(rank)])
    dist.all_reduce(y, group=pgc.dp)
    assert y.item() == {0: 2.0, 2: 2.0, 1: 4.0, 3: 4.0}[rank]

    # DDP bundle: expt_dp missing → a self group is created; one optimizer instance → intra == full
    ddp_cfg = SimpleNamespace(num_distributed_optimizer_instances=1, use_distributed_optimizer=True)
    cfg = SimpleNamespace(context_parallel_size=1)
    with pytest.raises(ValueError, match="tp, pp and ep"):
        ProcessGroupCollection.setup_process_groups_for_ddp(pgc, cfg, ddp_cfg)
    pgc.pp = dist.new_group([rank], use_local_synchronization=True)
    d = ProcessGroupCollection.setup_process_groups_for_ddp(pgc, cfg, ddp_cfg)
    assert d["dp_group"] is pgc.dp and d["dp_cp_group"] is pgc.dp and d["intra_dp_cp_group"] is pgc.dp and d["expt_dp_group"].size() == 1
    assert d["inter_dist_opt_group"] is None and d["tp_group"] is pgc.tp and not ProcessGroupCollection.is_gtp_remat_active(d)
    with pytest.raises(ValueError, match="multiple optimizer instances"):
        ProcessGroupCollection.setup_process_groups_for_ddp(pgc, cfg, SimpleNamespace(num_distributed_optimizer_instances=2))
    with pytest.raises(ValueError, match="dp_cp process group is required"):
        ProcessGroupCollection(dp=pgc.dp)._resolve_data_groups(2, 1, True)
    # optimizer bundle
    pgc.expt_dp = d["expt_dp_group"]
    chunk = SimpleNamespace(ddp_config=ddp_cfg, config=cfg)
    with pytest.raises(ValueError, match="mp and expt_tp_pp"):
        ProcessGroupCollection.setup_process_groups_for_optimizer(ProcessGroupCollection(dp=pgc.dp, expt_dp=pgc.expt_dp), [chunk])
    pgc.expt_tp_pp = pgc.expt_tp
    o = ProcessGroupCollection.setup_process_groups_for_optimizer(pgc, [chunk])
    assert o["mp_group"] is pgc.tp and o["expt_tp_pp_group"] is pgc.expt_tp and o["intra_dist_opt_group"] is pgc.dp and o["intra_dp_cp_group_gloo"] is None

    # two modules on disjoint halves of the world; every rank builds both grids' groups, keeps its own module
    enc, llm = HyperCommGrid([2, 1], ["tp", "dp"], rank_offset=0), HyperCommGrid([1, 2], ["tp", "dp"], rank_offset=2)
    assert enc.is_current_rank_in_grid() == (rank < 2) and llm.is_current_rank_in_grid() == (rank >= 2)
    mm = MultiModuleProcessGroupCollection.from_grids({"encoder": enc, "llm": llm}, language_model_module_name="llm")
    assert list(mm.keys()) == (["encoder"] if rank < 2 else ["llm"]) and mm.has_language_model() == (rank >= 2) and len(mm) == 1
    if rank >= 2:
        assert mm.get_language_model_cp_size() == 1 and mm["llm"].dp.size() == 2 and "llm" in mm
    else:
        with pytest.raises(ValueError):
            mm.get_language_model_collection()
        with pytest.raises(KeyError):
            mm["llm"]
    with pytest.raises(ValueError):
        MultiModuleProcessGroupCollection({}, None)
    dist.barrier()
    for g in (grid, enc, llm):
        g.destroy()
        assert not g._pgs
    return True


def test_grid_groups_views_and_collections_gloo():
    assert all(run_distributed(_grid_worker, 4))


def _mpu_bundle_worker(rank, world):
    from megatron_b200.core import parallel_state as ps
    from megatron_b200.core.process_groups_config import ProcessGroupCollection

    ps.initialize_model_parallel(tensor_model_parallel_size=2)
    ddp_cfg = SimpleNamespace(num_distributed_optimizer_instances=1, use_distributed_optimizer=True)
    d = ProcessGroupCollection.setup_process_groups_for_ddp(None, SimpleNamespace(context_parallel_size=1), ddp_cfg)
    assert d["dp_group"].size() == 2 and d["tp_group"].size() == 2 and d["inter_dist_opt_group"] is None and d["intra_dist_opt_group"] is not None
    o = ProcessGroupCollection.setup_process_groups_for_optimizer(None, [SimpleNamespace(ddp_config=ddp_cfg)])
    assert o["mp_group"].size() == 2 and o["intra_dp_cp_group_gloo"] is not None and "tp_group" not in o
    full = ProcessGroupCollection.use_mpu_process_groups()
    assert full.intra_dist_opt is full.intra_dp_cp and full.expt_tp_pp is full.tp_ep_pp
    ps.destroy_model_parallel()
    return True


def test_bundles_from_parallel_state_gloo():
    assert all(run_distributed(_mpu_bundle_worker, 4))


def _extra_groups_worker(rank, world):
    import torch
    import torch.distributed as dist

    from megatron_b200.core import parallel_state as ps

    ps.initialize_model_parallel(tensor_model_parallel_size=1, context_parallel_size=2, hybrid_context_parallel=True)     # dp = 4, cp = 2 over 8 ranks
    hy = ps.get_hybrid_data_context_parallel_groups()
    assert sorted(hy) == [2, 4] and dist.get_world_size(hy[2]) == 2 and dist.get_world_size(hy[4]) == 4
    assert ps.get_hybrid_data_context_parallel_groups(group_size=8) is ps.get_data_parallel_group(with_context_parallel=True)
    t = torch.ones(1)
    dist.all_reduce(t, group=hy[4])
    assert float(t) == 4
    ag = ps.create_all_gather_groups(("dp_cp",))
    assert dist.get_process_group_ranks(ag["dp_cp"]) == dist.get_process_group_ranks(ps.get_data_parallel_group(with_context_parallel=True))
    assert ag["dp_cp"] is not ps.get_data_parallel_group(with_context_parallel=True) and ps.get_all_gather_group("dp_cp") is ag["dp_cp"]
    levels, gloo = ps.create_hierarchical_groups(rank, list(range(8)), [2, 4], group_desc="TEST")
    assert gloo is None and dist.get_process_group_ranks(levels[0]) == [rank // 2 * 2, rank // 2 * 2 + 1]
    assert dist.get_process_group_ranks(levels[1]) == list(range(rank % 2, 8, 2))
    cfg = {}
    ps.overwrite_nccl_comm_cfgs(cfg, "tp", ("max_ctas", 8))
    assert cfg == {"tp": {"max_ctas": 8}}
    ps.set_data_parallel_rank(3)
    assert ps.get_data_parallel_rank() == 3
    ps.set_data_parallel_rank(None)
    assert ps.get_gtp_weight_remat_global_ranks() == [rank] and ps.get_expert_gtp_weight_remat_rank() == 0
    assert ps.get_context_parallel_world_size() == 2 and ps.get_tensor_and_context_parallel_world_size() == 2
    return True


def test_hybrid_cp_all_gather_and_hierarchical_group_builders():
    from dist_utils import run_distributed

    assert all(run_distributed(_extra_groups_worker, 8))
