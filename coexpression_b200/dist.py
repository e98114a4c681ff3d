"""Multi-GPU plumbing: permutations are independent units (the reference runs them as separate
processes, runlocal.py:50-55, and only 2-collate-results.py:68-76 pools their density vectors),
so they are sharded across ranks with no data-path collective; the ONE exchange is the gather
of the per-permutation density vectors.  One process per GPU, torch.distributed (NCCL on
GPUs over NVLink/NVSwitch; gloo on CPU for the tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_permutations(hit_counts, world):
    """Deterministic longest-processing-time assignment by P^2 (score-matrix size ~ work).
    -> list over ranks of lists of permutation indices (each ascending)."""
    cost = np.asarray(hit_counts, dtype=np.float64) ** 2
    order = sorted(range(len(cost)), key=lambda k: (-cost[k], k))
    load = [0.0] * world
    shards = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda i: (load[i], i))
        shards[r].append(k)
        load[r] += cost[k]
    return [sorted(s) for s in shards]


def pack_densities(densities, device):
    """densities: list (local permutations) of 1-D float arrays -> (padded [K, Gmax] f64, lengths [K] i64)."""
    K = len(densities)
    gmax = max([len(d) for d in densities], default=0)
    buf = torch.zeros((K, max(gmax, 1)), dtype=torch.float64, device=device)
    lens = torch.zeros(K, dtype=torch.int64, device=device)
    for i, d in enumerate(densities):
        if len(d):
            buf[i, :len(d)] = torch.as_tensor(np.asarray(d, dtype=np.float64), device=device)
        lens[i] = len(d)
    return buf, lens


def gather_densities(local_perm_ids, densities, numperms_total, device=None):
    """All ranks call this; every rank gets the density vectors of ALL permutations, in
    permutation order.  One all_gather of sizes (tiny), then one all_gather of the padded block."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = device or torch.device("cpu")
    if world == 1:
        out = [None] * numperms_total
        for k, d in zip(local_perm_ids, densities):
            out[k] = np.asarray(d, dtype=np.float64)
        return out
    buf, lens = pack_densities(densities, device)
    meta = torch.tensor([buf.shape[0], buf.shape[1]], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    kmax = int(max(m[0].item() for m in metas))
    gmax = int(max(m[1].item() for m in metas))
    block = torch.zeros((kmax, gmax + 2), dtype=torch.float64, device=device)   # [:, 0] = perm id, [:, 1] = length
    block[:, 0] = -1
    if buf.shape[0]:
        block[:buf.shape[0], 0] = torch.as_tensor(np.asarray(local_perm_ids, dtype=np.float64), device=device)
        block[:buf.shape[0], 1] = lens.to(torch.float64)
        block[:buf.shape[0], 2:2 + buf.shape[1]] = buf
    blocks = [torch.zeros_like(block) for _ in range(world)]
    dist.all_gather(blocks, block)
    out = [None] * numperms_total
    for b in blocks:
        b = b.cpu().numpy()
        for row in b:
            k = int(row[0])
            if k >= 0:
                out[k] = row[2:2 + int(row[1])].copy()
    return out


# ------------------------------------------------------------------------------------------------
# C3: all unordered row pairs.  Row block b = rows [b * block, (b + 1) * block) against every LATER row; the blocks
# are independent (no exchange), their cost falls linearly with b, so they are dealt to the ranks in a zig-zag
# (0, 1, .., w-1, w-1, .., 1, 0, 0, 1, ..) that balances the pair counts; the ONE collective is the reduction of the
# per-rank histograms of r.
# ------------------------------------------------------------------------------------------------
def shard_row_blocks(n_rows, world, block=1024):
    """-> list over ranks of lists of (row0, row1) blocks."""
    blocks = [(r0, min(n_rows, r0 + block)) for r0 in range(0, n_rows, block)]
    shards = [[] for _ in range(world)]
    for b, blk in enumerate(blocks):
        lap, pos = divmod(b, world)
        shards[pos if lap % 2 == 0 else world - 1 - pos].append(blk)
    return shards


def pairs_of_blocks(n_rows, blocks):
    """Unordered pairs (i < j) whose first row lies in one of the blocks."""
    return int(sum((r1 - r0) * n_rows - (r0 + r1 - 1) * (r1 - r0) // 2 - (r1 - r0) for r0, r1 in blocks))


def reduce_histogram(hist, nan_count, device=None):
    """Sum of the per-rank histograms (uint64 counts carried as int64) and NaN-pair counts on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return np.asarray(hist, dtype=np.uint64), int(nan_count)
    device = device or torch.device("cpu")
    buf = torch.empty(len(hist) + 1, dtype=torch.int64, device=device)
    buf[:-1] = torch.as_tensor(np.asarray(hist, dtype=np.uint64).view(np.int64), device=device)
    buf[-1] = int(nan_count)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    return out[:-1].view(np.uint64).copy(), int(out[-1])
