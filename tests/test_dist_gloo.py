"""world_size-2 gloo test of the N>1 host logic (sharding + the single density gather). CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coexpression_b200.dist import (gather_densities, pairs_of_blocks, reduce_histogram, shard_permutations,
                                    shard_row_blocks)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _fake_density(k):
    rng = np.random.default_rng(1000 + k)
    return rng.random(3 + (k * 7) % 11) * k


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = shard_permutations(counts, world)
    mine = shards[rank]
    out = gather_densities(mine, [_fake_density(k) for k in mine], len(counts))
    ok = all(np.array_equal(out[k], _fake_density(k)) for k in range(len(counts)))
    q.put((rank, ok, mine))
    dist.destroy_process_group()


def test_shard_is_a_balanced_partition():
    counts = [1690] + [600 + (37 * k) % 90 for k in range(100)]
    for world in (1, 2, 4, 8):
        shards = shard_permutations(counts, world)
        assert sorted(k for s in shards for k in s) == list(range(len(counts)))
        load = [sum(counts[k] ** 2 for k in s) for s in shards]
        assert max(load) <= min(load) + max(counts) ** 2
    assert shard_permutations(counts, 4) == shard_permutations(counts, 4)      # deterministic


def test_gather_world2_gloo():
    counts = [50, 7, 0, 1, 33, 12, 90, 3, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in got)
    assert sorted(k for _, _, mine in got for k in mine) == list(range(len(counts)))


def test_gather_single_process():
    out = gather_densities([0, 1, 2], [_fake_density(k) for k in range(3)], 3)
    assert all(np.array_equal(out[k], _fake_density(k)) for k in range(3))


def test_row_block_shards_partition_all_pairs():
    for n_rows, block in ((10, 3), (5000, 256), (200000, 1024)):
        total = n_rows * (n_rows - 1) // 2
        for world in (1, 2, 4, 8):
            shards = shard_row_blocks(n_rows, world, block)
            rows = sorted(r for s in shards for blk in s for r in range(*blk)) if n_rows <= 5000 else None
            if rows is not None:
                assert rows == list(range(n_rows))
            per = [pairs_of_blocks(n_rows, s) for s in shards]
            assert sum(per) == total
            if n_rows >= 5000 and n_rows // block >= 8 * world:
                assert max(per) <= 1.1 * min(per)                 # zig-zag deal: balanced pair counts


def _hist_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5 + rank)
    h = rng.integers(0, 1 << 40, size=64).astype(np.uint64)
    tot, nan = reduce_histogram(h, 10 + rank)
    q.put((rank, h.tolist(), tot.tolist(), nan))
    dist.destroy_process_group()


def test_histogram_reduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [a + b for a, b in zip(got[0][1], got[1][1])]
    assert got[0][2] == want and got[1][2] == want and got[0][3] == got[1][3] == 21


def test_score_layout_partitions_the_score_buffers():
    from coexpression_b200.dist import score_layout
    counts = [1690] + [600 + (37 * k) % 90 for k in range(40)] + [0, 1]
    for world in (1, 2, 4, 8):
        shards, owner, soff, sizes = score_layout(counts, world)
        assert sorted(k for s in shards for k in s) == list(range(len(counts)))
        for r, mine in enumerate(shards):
            assert mine == sorted(mine) and all(owner[k] == r for k in mine)
            acc = 0
            for k in mine:                                       # a rank's blocks follow each other in ascending k
                assert soff[k] == acc
                acc += counts[k] ** 2
            assert sizes[r] == acc
