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
# ONE job shared by the GPUs of a box: count stage sharded by table row, network stage by permutation, the edge scores
# stored straight into the owner's buffer over NVLink (include/coexpression_b200.h section 3).  The host side here only
# deals the permutations, exchanges the 64-byte IPC handles once and places one barrier between the two stages.
# ------------------------------------------------------------------------------------------------
def score_layout(hit_counts, world):
    """-> (shards, owner [K], score_off [K], n_doubles [world]): permutation k lives in the buffer of owner[k] at
    score_off[k] (doubles); a rank's blocks follow each other in ascending k."""
    shards = shard_permutations(hit_counts, world)
    K = len(hit_counts)
    owner = np.zeros(K, dtype=np.int32)
    soff = np.zeros(K, dtype=np.int64)
    sizes = np.zeros(world, dtype=np.int64)
    for r, mine in enumerate(shards):
        acc = 0
        for k in mine:
            owner[k] = r
            soff[k] = acc
            acc += int(hit_counts[k]) ** 2
        sizes[r] = acc
    return shards, owner, soff, sizes


class SharedJob:
    """All ranks construct it with an initialised NCCL process group and call plan() / run() with the SAME list of hit sets.
    The score buffers are allocated (and their IPC handles exchanged) once and again only when a larger list arrives."""

    def __init__(self, ctx, hit_sets, device):
        self.ctx, self.device = ctx, device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.cap = -1
        self.hit_sets = None
        if hit_sets is not None:
            self.plan(hit_sets)

    def plan(self, hit_sets):
        self.hit_sets = hit_sets
        counts = [len(h) for h in hit_sets]
        self.shards, self.owner, self.score_off, sizes = score_layout(counts, self.world)
        self.mine = self.shards[self.rank]
        need = int(sizes.max()) if len(sizes) else 0          # the same number on every rank: they (re)allocate together
        if need > self.cap:
            self.barrier()                                     # nobody still writes into the buffers about to be replaced
            handle = self.ctx.shard_scores_alloc(self.world, self.rank, need)
            self.cap = need
            if self.world > 1:
                t = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
                got = [torch.zeros_like(t) for _ in range(self.world)]
                dist.all_gather(got, t)
                for r, h in enumerate(got):
                    if r != self.rank:
                        self.ctx.shard_scores_open(r, bytes(h.cpu().tolist()))

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()

    def _exchange(self, payload):
        """all-gather of one byte string per rank -> list over ranks."""
        if self.world == 1:
            return [payload]
        t = torch.tensor(list(payload), dtype=torch.uint8, device=self.device)
        got = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(got, t)
        return [bytes(h.cpu().tolist()) for h in got]

    def prepare(self):
        """Rank transform + packing of the table the contexts hold (every rank: the same table, set_expression done): each
        rank ranks its block of rows and stores the results into every rank's packed table over NVLink.  Replaces ctx.prepare()."""
        shape = (self.ctx.N, self.ctx.n)
        if getattr(self, "_table_shape", None) != shape:
            self.barrier()
            handles = self._exchange(self.ctx.shard_table_export(self.world, self.rank))
            for r, h in enumerate(handles):
                if r != self.rank:
                    self.ctx.shard_table_open(r, h)
            self._table_shape = shape
        self.barrier()                       # nobody still reads the packed table of the previous run
        self.ctx.shard_prepare_rows()
        self.barrier()                       # every block of rows has landed everywhere
        self.ctx.shard_prepare_finish()

    def run(self, options, want_scores=False, hit_sets=None):
        """-> api.BatchResult of this rank's permutations (self.mine)."""
        if hit_sets is not None:
            self.plan(hit_sets)
        self.barrier()                       # nobody is still reading last run's scores
        self.ctx.shard_count(self.hit_sets, options, self.owner, self.score_off)
        self.barrier()                       # every rank's score rows have landed
        return self.ctx.shard_finish([self.hit_sets[k] for k in self.mine], options, want_scores=want_scores)


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
