"""N ranks == one rank, bit for bit: hit sets sharded with dist.shard_permutations, one NCCL gather of the density
vectors (dist.gather_densities) -- the pooling point of 2-collate-results.py:68-76.  Needs >= 2 GPUs on the box
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest

from conftest import has_cuda


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda() or _ngpu() < 2, reason="needs two CUDA devices")]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _inputs():
    from coexpression_b200 import synth
    t = synth.make_table(5000, 128, seed=41)
    hits, _, _ = synth.make_permutation_hits(t, 120, 12, window=8000, seed=42)
    ga, gb = synth.global_pair_indices(t.names, 40000, seed=43)
    return t, hits, ga, gb


def _densities(ctx, t, hits, ga, gb, opts):
    ctx.set_expression(t.X); ctx.prepare()
    ctx.set_features(t.chrom, t.start, t.end, t.names)
    g = ctx.pairs(ga, gb, ranked=True)
    ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
    res = ctx.network_batch(hits, opts)
    return [res.perm(k)["density"].copy() for k in range(len(hits))]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.dist import gather_densities, shard_permutations
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    t, hits, ga, gb = _inputs()
    mine = shard_permutations([len(h) for h in hits], world)[rank]
    with CoexContext(rank) as ctx:
        local = _densities(ctx, t, [hits[k] for k in mine], ga, gb, Options(anticorrelation=True))
    allden = gather_densities(mine, local, len(hits), device=dev)
    q.put((rank, mine, [d.tolist() for d in allden]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_one_gpu_densities_bit_for_bit():
    import torch.multiprocessing as mp
    from coexpression_b200.api import CoexContext, Options
    t, hits, ga, gb = _inputs()
    with CoexContext(0) as ctx:
        want = _densities(ctx, t, hits, ga, gb, Options(anticorrelation=True))
    ctx_mp = mp.get_context("spawn")
    q = ctx_mp.Queue()
    port = _free_port()
    procs = [ctx_mp.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(k for _, mine, _ in got for k in mine) == list(range(len(hits)))
    for _, _, allden in got:                                   # every rank holds every permutation's vector
        assert len(allden) == len(hits)
        for k in range(len(hits)):
            assert np.array_equal(np.array(allden[k]), want[k]), k


def _shared_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.dist import SharedJob, gather_densities
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    t, hits, ga, gb = _inputs()
    opts = Options(anticorrelation=True)
    with CoexContext(rank) as ctx:
        ctx.set_expression(t.X)
        job = SharedJob(ctx, None, dev)
        job.prepare()                                            # every rank ranks half of the rows, for both packed tables
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        g = ctx.pairs(ga, gb, ranked=True)
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        job.plan(hits)
        res = job.run(opts, want_scores=True)
        job.prepare()
        res2 = job.run(opts)                                     # a second run over the same (reused) peer buffers
        local = [res.perm(j)["density"].copy() for j in range(len(job.mine))]
        same = all(np.array_equal(res2.perm(j)["density"], local[j]) for j in range(len(job.mine)))
        allden = gather_densities(job.mine, local, len(hits), device=dev)
        job.barrier()
    q.put((rank, job.mine, [d.tolist() for d in allden], same, [res.score_matrix(j).tolist() for j in range(min(2, len(job.mine)))]))
    dist.barrier()
    dist.destroy_process_group()


def test_one_job_shared_by_two_gpus_rows_by_rank_scores_over_peer_memory():
    """include/coexpression_b200.h section 3: count stage sharded by table row, score rows stored into the owner's buffer over
    NVLink (CUDA IPC), network stage by permutation -- densities AND score matrices equal to the one-GPU run bit for bit."""
    import torch.multiprocessing as mp
    from coexpression_b200.api import CoexContext, Options
    t, hits, ga, gb = _inputs()
    with CoexContext(0) as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        g = ctx.pairs(ga, gb, ranked=True)
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        want = ctx.network_batch(hits, Options(anticorrelation=True), want_scores=True)
    ctx_mp = mp.get_context("spawn")
    q = ctx_mp.Queue()
    port = _free_port()
    procs = [ctx_mp.Process(target=_shared_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(k for _, mine, _, _, _ in got for k in mine) == list(range(len(hits)))
    for _, mine, allden, same, scores in got:
        assert same
        for k in range(len(hits)):
            assert np.array_equal(np.array(allden[k]), want.perm(k)["density"]), k
        for j, sc in enumerate(scores):
            assert np.array_equal(np.array(sc), want.score_matrix(mine[j])), mine[j]
