"""Host-side sharding logic over a world_size-2 gloo group on CPU: frame
blocks tile the trajectory, replicas partition by r mod G, statistics all-reduce to the
single-process result.  No GPU and no CUDA library call in this file."""
import os
import socket

import numpy as np
import pytest

from cmdlmc_b200 import parallel


def test_frame_blocks_tile_the_trajectory():
    for n in (0, 1, 7, 100, 100000, 12345):
        for world in (1, 2, 3, 4, 8):
            blocks = [parallel.frame_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.frame_block(10, 2, 2)


def test_replicas_partition():
    for n in (0, 1, 5, 1024):
        for world in (1, 2, 8):
            ids = [parallel.replica_ids(n, r, world) for r in range(world)]
            allids = np.sort(np.concatenate(ids)) if ids else np.zeros(0)
            np.testing.assert_array_equal(allids, np.arange(n))
            assert all((i % world == r).all() for r, i in enumerate(ids))


def test_single_process_paths():
    s = {"hist": np.arange(5, dtype=np.int64), "msd": np.array([1.5, 2.5])}
    out = parallel.allreduce_sum(s)
    np.testing.assert_array_equal(out["hist"], s["hist"])
    np.testing.assert_array_equal(out["msd"], s["msd"])
    rng = np.random.RandomState(0)
    rows = [np.column_stack([np.arange(4), np.arange(4) * 0.5, rng.rand(4, 4)]) for _ in range(6)]
    m = parallel.merge_observables(rows)
    data = np.stack([r[:, 2:6] for r in rows])
    np.testing.assert_allclose(m["mean"], data.mean(axis=0))
    np.testing.assert_allclose(m["sem"], data.std(axis=0, ddof=1) / np.sqrt(6))
    assert m["n"] == 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.rank_world() == (rank, world)
        # 1. frame-block statistics: every rank histograms its own frames
        n_frames, nbins = 1001, 50
        rng = np.random.RandomState(42)
        dists = rng.uniform(0, 5, size=(n_frames, 30))
        a, b = parallel.frame_block(n_frames, rank, world)
        hist = np.histogram(dists[a:b], bins=nbins, range=(0, 5))[0].astype(np.int64)
        rsum = np.array([dists[a:b].sum(), float(b - a)])
        tot = parallel.allreduce_sum({"hist": hist, "rsum": rsum})
        # 2. replica statistics: every rank merges the replicas it owns
        rrng = np.random.RandomState(7)
        all_rows = [np.column_stack([np.arange(5), np.arange(5) * 0.4, rrng.rand(5, 4)])
                    for _ in range(9)]
        mine = [all_rows[i] for i in parallel.replica_ids(9, rank, world)]
        merged = parallel.merge_observables(mine)
        q.put((rank, tot, merged, dists, all_rows))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_statistics_reduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tot, merged, dists, all_rows in results:
        want = np.histogram(dists, bins=50, range=(0, 5))[0]
        np.testing.assert_array_equal(tot["hist"], want)          # integers: exact
        np.testing.assert_allclose(tot["rsum"], [dists.sum(), 1001.0], rtol=1e-13)
        data = np.stack([r[:, 2:6] for r in all_rows])
        assert merged["n"] == 9
        np.testing.assert_allclose(merged["mean"], data.mean(axis=0), rtol=1e-13)
        np.testing.assert_allclose(merged["sem"], data.std(axis=0, ddof=1) / 3.0, rtol=1e-10)
