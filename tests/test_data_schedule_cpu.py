"""Sequence-packing schedules, multimodal mock data, object-storage readers."""
import os

import pytest
import torch

from dist_utils import run_distributed


def _lens(seed, n):
    g = torch.Generator().manual_seed(seed)
    short = torch.randint(16, 200, (n,), generator=g)
    long_ = torch.randint(600, 1000, (n,), generator=g)
    return torch.where(torch.rand(n, generator=g) < 0.2, long_, short).tolist()


def _samples(rank, n):
    out = []
    for i, L in enumerate(_lens(100 + rank, n)):
        tok = torch.full((L,), rank * 1000 + i, dtype=torch.long)
        out.append({"tokens": tok, "labels": tok + 1, "loss_mask": torch.ones(L), "ignored_2d": torch.zeros(2, 2)})
    return out


def _packing(rank, world):
    from megatron_b200.core.datasets.data_schedule import DpBalancedScheduler, NaiveSequentialScheduler, PackingSchedulerEnum, get_batch_on_this_rank_for_sequence_packing, wrap_data_iterator

    n = 12
    it = wrap_data_iterator(iter(_samples(rank, n) * 2), PackingSchedulerEnum.DP_BALANCED, max_seqlen_per_rank=1024, samples_per_rank=n, pad_to_multiple=8)
    first = get_batch_on_this_rank_for_sequence_packing(it)
    batches = [first] + [next(it) for _ in range(it.num_microbatches - 1)]
    seen, total = [], 0
    for b in batches:
        p = b["packed_seq_params"]
        T = b["tokens"].shape[1]
        assert T <= 1024 + 8 * len(b["seqlens"]) and T % 8 == 0 and b["labels"].shape == b["tokens"].shape and "ignored_2d" not in b
        assert int(p.cu_seqlens_q[-1]) == T and p.qkv_format == "thd"
        for j, L in enumerate(b["seqlens"]):
            a = int(p.cu_seqlens_q[j])
            seg = b["tokens"][0, a : a + L]
            assert (seg == seg[0]).all() and (b["labels"][0, a : a + L] == seg[0] + 1).all()          # a sample stays contiguous and whole
            assert b["loss_mask"][0, a : a + L].all() and not b["loss_mask"][0, a + L : int(p.cu_seqlens_q[j + 1])].any()
            assert (b["position_ids"][0, a : a + L] == torch.arange(L)).all()
            seen.append(int(seg[0]))
            total += L
    # every sample of the global batch is processed exactly once across the ranks
    import torch.distributed as dist

    allseen = [None] * world
    dist.all_gather_object(allseen, seen)
    flat = sorted(x for s in allseen for x in s)
    assert flat == sorted(r * 1000 + i for r in range(world) for i in range(n))
    counts = [None] * world
    dist.all_gather_object(counts, (it.num_microbatches, total))
    assert len({c[0] for c in counts}) == 1                                                   # same number of micro-batches everywhere
    # the balanced plan beats keeping samples where they were drawn
    lens = [(r * n + i, L) for r in range(world) for i, L in enumerate(_lens(100 + r, n))]
    bal, naive = DpBalancedScheduler(1024), NaiveSequentialScheduler(1024)
    bal.dp_size = naive.dp_size = world
    ib, inv = bal.imbalance(lens, bal.get_groups_and_subsamples(lens)), bal.imbalance(lens, naive.get_groups_and_subsamples(lens))
    assert ib < 1.1 and ib <= inv, (ib, inv)
    with pytest.raises(ValueError):
        bal.get_groups_and_subsamples([(0, 5000)])
    return True


def test_dp_balanced_sequence_packing_routes_and_packs():
    run_distributed(_packing, 2)


def _hybrid(rank, world):
    from megatron_b200.core.datasets.data_schedule import HybridCPDataLoaderWrapper

    def gen():
        i = 0
        while True:
            L = 3000 if (rank == 0 and i == 0) else 200 + 10 * i                              # one sample needs 4 ranks at 1024 tokens / rank
            yield {"tokens": torch.full((L,), rank * 100 + i, dtype=torch.long)}
            i += 1

    it = HybridCPDataLoaderWrapper(gen(), max_seqlen_per_rank=1024, samples_per_rank=3, dp_cp_group=None)
    b = next(it)
    batches = [b] + [next(it) for _ in range(it.num_microbatches - 1)]
    ids = [(int(x["tokens"][0, int(x["packed_seq_params"].cu_seqlens_q[j])]), tuple(x["cp_ranks"][j])) for x in batches for j in range(len(x["seqlens"]))]
    long_ = [r for v, r in ids if v == 0]
    assert long_ == [(0, 1, 2, 3)]                                                          # the 3000-token sample is on all four ranks
    import torch.distributed as dist

    every = [None] * world
    dist.all_gather_object(every, ids)
    where = {}
    for r, lst in enumerate(every):
        for v, ranks in lst:
            where.setdefault(v, []).append(r)
            assert r in ranks
    # every sample of the global batch is resident on exactly the ranks of its CP block (the scheduler may widen short samples too)
    assert sorted(where) == sorted(r * 100 + i for r in range(world) for i in range(3))
    cp = {v: ranks for lst in every for v, ranks in lst}
    assert all(sorted(where[v]) == list(cp[v]) for v in where)
    return True


def test_hybrid_cp_loader_replicates_long_samples_over_their_cp_block():
    run_distributed(_hybrid, 4)


def test_mock_multimodal_dataset_and_object_storage_reader(tmp_path):
    from megatron_b200.core.datasets import object_storage_utils as osu
    from megatron_b200.core.datasets.multimodal_dataset import MockMultimodalDataset, MultimodalDatasetConfig
    from megatron_b200.core.tokenizers import build_tokenizer

    tok = build_tokenizer("NullTokenizer", vocab_size=100)
    cfg = MultimodalDatasetConfig(random_seed=7, sequence_length=16, tokenizer=tok, image_h=8, image_w=12, split="1,0,0", reset_position_ids=False,
                                  reset_attention_mask=False, eod_mask_loss=False, preprocess_func=lambda s: {**s, "has_image": torch.tensor(True)})
    from megatron_b200.core.datasets.gpt_dataset import MockGPTLowLevelDataset
    from megatron_b200.core.datasets.utils import Split

    ds = MockMultimodalDataset(MockGPTLowLevelDataset(tok), None, torch.arange(10).numpy(), 10, Split.train, cfg)
    a, b = ds[3], ds[3]
    assert a["image"].shape == (3, 8, 12) and torch.equal(a["image"], b["image"]) and not torch.equal(a["image"], ds[4]["image"]) and bool(a["has_image"])
    assert a["tokens"].shape[0] == 16

    assert osu.is_object_storage_path("s3://b/k.idx") and not osu.is_object_storage_path("/data/k.idx")
    assert osu.parse_s3_path("s3://bucket/a/b.bin") == ("bucket", "a/b.bin")
    with pytest.raises(ValueError):
        osu.parse_s3_path("s3:///x")
    blob = bytes(range(256)) * 64

    class Fake:
        def __init__(self):
            self.downloads = 0

        def head_object(self, Bucket, Key):
            if Key != "d/x.idx":
                raise KeyError(Key)
            return {}

        def download_file(self, Bucket, Key, Filename):
            self.downloads += 1
            open(Filename, "wb").write(b"IDX")

        def get_object(self, Bucket, Key, Range):
            a, b = Range[len("bytes="):].split("-")
            return {"Body": blob[int(a) : int(b) + 1]}

    fake = Fake()
    cfg = osu.ObjectStorageConfig(path_to_idx_cache=str(tmp_path / "cache"), bin_chunk_nbytes=1024)
    assert osu.object_exists("s3://bk/d/x.idx", fake) and not osu.object_exists("s3://bk/d/y.idx", fake)
    p1 = osu.cache_index_file("s3://bk/d/x.idx", cfg, fake, rank=0)
    p2 = osu.cache_index_file("s3://bk/d/x.idx", cfg, fake, rank=0)
    assert p1 == p2 == os.path.join(str(tmp_path / "cache"), "bk", "d/x.idx") and fake.downloads == 1 and open(p1, "rb").read() == b"IDX"
    rd = osu.ObjectStorageBinReader("s3://bk/d/x.bin", cfg, fake)
    assert rd.read(10, 20) == blob[10:30] and rd.read(500, 100) == blob[500:600] and rd.requests == 1          # same 1 KiB chunk
    assert rd.read(1000, 100) == blob[1000:1100] and rd.requests == 2                                            # straddles → refetch
    assert rd.read(4096, 3000) == blob[4096:7096]
