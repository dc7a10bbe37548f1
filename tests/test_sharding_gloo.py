"""world_size-2 gloo test (CPU) of the multi-GPU ranking plumbing: shards cover the test set exactly
once and the all-reduced sums equal the single-process sums.  The per-shard "ranking" is done by the
CPU oracle here; on the GPU box kb2e_rank(first, count) takes its place (tests/test_gpu_rank.py checks
that a window equals a slice of the whole)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    from kb2e_b200 import kg, sharding
    from kb2e_oracle import Oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = kg.make_kg("tiny", seed=5)
    rng = np.random.default_rng(1)
    D = 8
    ent = rng.normal(0, 0.3, (g["nE"], D))
    rel = rng.normal(0, 0.3, (g["nR"], D))
    test = g["test"][:41]  # odd count: uneven shards
    lo, hi = sharding.shard_bounds(len(test), rank, world)
    o = Oracle()
    filt = np.concatenate([g["train"], g["valid"], test])  # the filter always holds the whole test set
    rlo, rhi, flo, fhi = o.rank(0, 0, ent, rel, None, test[lo:hi], filt)
    sums = np.array([rlo.sum(), flo.sum(), (rlo <= 10).sum(), (flo <= 10).sum()], dtype=np.int64)
    total = sharding.allreduce_sums(sums)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([[lo, hi], total]))
    dist.destroy_process_group()


def test_two_rank_sharded_ranking_sums(tmp_path, oracle):
    from kb2e_b200 import kg, sharding
    world, port = 2, 29000 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 41          # shards tile the test set
    assert np.array_equal(r0[2:], r1[2:])                          # every rank holds the global sums
    g = kg.make_kg("tiny", seed=5)
    rng = np.random.default_rng(1)
    ent = rng.normal(0, 0.3, (g["nE"], 8))
    rel = rng.normal(0, 0.3, (g["nR"], 8))
    test = g["test"][:41]
    rlo, rhi, flo, fhi = oracle.rank(0, 0, ent, rel, None, test, np.concatenate([g["train"], g["valid"], test]))
    want = np.array([rlo.sum(), flo.sum(), (rlo <= 10).sum(), (flo <= 10).sum()])
    assert np.array_equal(r0[2:], want)
    m = sharding.merge_metrics(r0[2:], 82)
    assert abs(m["raw_mean_rank"] - rlo.mean()) < 1e-12


def test_shard_bounds_properties():
    from kb2e_b200.sharding import shard_bounds
    for n in (0, 1, 7, 59071):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
