"""Host-side handle over the batched C ABI: the table is made resident once and K permutations
are reduced to network densities per call (the reference re-reads, re-ranks and re-marshals the
whole table in each of its per-permutation processes, 1-make-network.py:70,100-122)."""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import CoexParams, check

STAGES = ["rank_pack", "rowstats", "gather", "corr_gemm", "count_score", "network"]    # (stage 6 of the library, the split count pipeline, is gone)


@dataclass
class Options:
    """The options of 0-prepare-coex.py that change hot-path arithmetic."""
    correlationmeasure: str = "Spearman"      # -cm
    anticorrelation: bool = False             # -a
    noiter: bool = False                      # -i
    useaverage: bool = False                  # -u
    nsp: float = 1.0                          # -s
    pval_to_join: float = 0.1                 # -j
    selectionthreshold: float = 0.0           # -st
    selectionthreshold_is_j: bool = False     # -m
    soft_separation: int = 100000             # -ss

    def to_params(self, N):
        if self.nsp < 1 and not (self.nsp <= 1.0 / N):
            # reference 1-make-network.py:200-201 uses `random` without importing it (quirk Q8)
            raise NameError("name 'random' is not defined")
        thr = self.selectionthreshold
        if self.selectionthreshold_is_j:                       # 1-make-network.py:95-96
            thr = -math.log(self.pval_to_join, 10)
        return CoexParams(
            measure=0 if self.correlationmeasure == "Spearman" else 1,
            anticorrelation=int(bool(self.anticorrelation)), noiter=int(bool(self.noiter)),
            useaverage=int(bool(self.useaverage)),
            use_specific_p=0 if self.nsp <= 1.0 / N else 1,    # 1-make-network.py:189-193
            reserved=0, pval_to_join=float(self.pval_to_join), selectionthreshold=float(thr),
            soft_separation=int(self.soft_separation))


@dataclass
class BatchResult:
    perm_off: np.ndarray        # [K+1]
    hit_rows: np.ndarray        # [H]
    n_groups: np.ndarray        # [K]
    group_of_hit: np.ndarray    # [H]
    group_label: np.ndarray     # [H] hit position of each group's label (first n_groups[k] slots)
    group_winner: np.ndarray    # [H]
    group_density: np.ndarray   # [H]
    hit_score: np.ndarray       # [H]
    scores: np.ndarray | None   # [sum P^2]
    counts: np.ndarray | None

    def perm(self, k):
        a, b = int(self.perm_off[k]), int(self.perm_off[k + 1])
        g = int(self.n_groups[k])
        return dict(hit_rows=self.hit_rows[a:b], group_of_hit=self.group_of_hit[a:b],
                    label=self.group_label[a:a + g], winner=self.group_winner[a:a + g],
                    density=self.group_density[a:a + g], hit_score=self.hit_score[a:b])

    def score_matrix(self, k):
        P = np.diff(self.perm_off)
        off = np.concatenate([[0], np.cumsum(P * P)])
        p = int(P[k])
        return self.scores[off[k]:off[k + 1]].reshape(p, p)

    def count_matrix(self, k):
        P = np.diff(self.perm_off)
        off = np.concatenate([[0], np.cumsum(P * P)])
        p = int(P[k])
        return self.counts[off[k]:off[k + 1]].reshape(p, p)


def _ptr(a):
    return None if a is None else a.ctypes.data


class CoexContext:
    """One GPU, one resident expression table."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        if self.lib.coex_device_count() <= 0:
            raise _lib.CoexError("no CUDA device: the coexpression hot path has no CPU fallback")
        h = ctypes.c_void_p()
        check(self.lib.coex_create(ctypes.byref(h), int(device)), "coex_create")
        self.h = h
        self.N = self.n = 0
        self._keep = []
        self._feature_chroms = None

    def close(self):
        if self.h:
            self.lib.coex_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- table -------------------------------------------------------------------------------
    def set_expression(self, X, pipelined=False):
        """X: float64 [N, n] numpy array (host) -- one bulk H2D copy (A3).
        pipelined=True: X is page-locked memory (e.g. the numpy view of a pinned torch tensor); the copy goes out in row chunks
        and prepare() ranks every chunk as it lands.  X is kept referenced here and must not be modified until a synchronising
        call (network_batch, sync) has returned."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.N, self.n = X.shape
        self._host_table = X if pipelined else None
        check(self.lib.coex_set_expression(self.h, X.ctypes.data, self.N, self.n, 2 if pipelined else 0), "coex_set_expression")

    def set_expression_device(self, dev_ptr, N, n, keepalive=None):
        """dev_ptr: device pointer to float64 [N, n] (e.g. torch tensor .data_ptr())."""
        self.N, self.n = int(N), int(n)
        self._keep.append(keepalive)
        check(self.lib.coex_set_expression(self.h, ctypes.c_void_p(dev_ptr), self.N, self.n, 1),
              "coex_set_expression")

    def prepare(self):
        check(self.lib.coex_prepare(self.h), "coex_prepare")

    def ranks(self):
        out = np.empty((self.N, self.n), dtype=np.float64)
        check(self.lib.coex_get_ranks(self.h, out.ctypes.data), "coex_get_ranks")
        return out

    def set_features(self, chrom, start, end, names):
        """chrom: per-row chromosome strings; names: row labels.  Strings are replaced by their
        ranks in sorted order, which preserves every comparison the reference makes on them
        (pandas sort on 'chromosome', coexfunctions.py:288; sorted(names), :314)."""
        chrom = np.asarray(chrom)
        uniq, chrom_id = np.unique(chrom, return_inverse=True)          # np.unique sorts
        self._feature_chroms = [str(u) for u in uniq]
        order = np.argsort(np.asarray(names), kind="stable")
        name_rank = np.empty(len(names), dtype=np.int32)
        name_rank[order] = np.arange(len(names), dtype=np.int32)
        self.set_features_raw(chrom_id.astype(np.int32), np.asarray(start, dtype=np.int64),
                              np.asarray(end, dtype=np.int64), name_rank)

    def permute_map(self, mode, snp_chrom, snp_addr, shifts, chromlist, chromrange, window, background=None,
                    capacity=None, picks=None, with_loci=False):
        """Row N1: the promoters of K permuted locus sets in one call (reference 0-prepare-coex.py:471-613 + windowBed).
        snp_chrom: chromosome names of the real loci, snp_addr their addresses; shifts[k] == 0 is the real data;
        chromlist / chromrange as built at 0-prepare-coex.py:319-330.  mode 'backcirc' needs `background` = the
        list of (chrom, address) sorted by (genome-wide address, key) (0-prepare-coex.py:582), which must contain
        every real locus.  mode 'background' (0-prepare-coex.py:564-578): `background` = (chrom, address) of every line
        of the background file in file order (column 0, column 2) and `picks[k]` = the M line numbers
        `random.sample(range(len(background)), M)` drew for set k (ignored where shifts[k] == 0).
        -> list of K int32 arrays of table rows (first-appearance order); with_loci: also a list of K int32 [P, 2] arrays
        holding, per hit, the first and the last locus (input order) that map to it."""
        if self._feature_chroms is None:
            raise _lib.CoexError("permute_map: call set_features first")
        cidx = {c: i for i, c in enumerate(chromlist)}
        fid = {c: i for i, c in enumerate(self._feature_chroms)}
        sc = np.ascontiguousarray([cidx[c] for c in snp_chrom], dtype=np.int32)
        sa = np.ascontiguousarray(snp_addr, dtype=np.int64)
        sh = np.ascontiguousarray(shifts, dtype=np.int64)
        cs = np.ascontiguousarray([chromrange[c][0] for c in chromlist], dtype=np.int64)
        ce = np.ascontiguousarray([chromrange[c][1] for c in chromlist], dtype=np.int64)
        cf = np.ascontiguousarray([fid.get(c, -1) for c in chromlist], dtype=np.int32)
        m = {"circular": 0, "backcirc": 1, "background": 2}[mode]
        sb = bc = ba = None
        nb = 0
        if m == 1:
            pos = {(c, int(a)): i for i, (c, a) in enumerate(background)}
            sb = np.ascontiguousarray([pos[(c, int(a))] for c, a in zip(snp_chrom, snp_addr)], dtype=np.int64)
            bc = np.ascontiguousarray([cidx[c] for c, _ in background], dtype=np.int32)
            ba = np.ascontiguousarray([a for _, a in background], dtype=np.int64)
            nb = len(background)
        K, M = len(sh), len(sc)
        if m == 2:
            sb = np.ascontiguousarray(picks, dtype=np.int64).reshape(K, M)
            bc = np.ascontiguousarray([cidx[c] for c, _ in background], dtype=np.int32)
            ba = np.ascontiguousarray([a for _, a in background], dtype=np.int64)
            nb = len(background)
        cap = int(capacity if capacity is not None else max(1024, 64 * K * M))
        off = np.zeros(K + 1, dtype=np.int64)
        rows = np.zeros(cap, dtype=np.int32)
        ptr = lambda x: x.ctypes.data if x is not None else None
        loci = np.zeros(2 * cap, dtype=np.int32) if with_loci else None
        check(self.lib.coex_permute_map(self.h, m, ptr(sc), ptr(sa), ptr(sb), M, ptr(sh), K, ptr(cs), ptr(ce), ptr(cf),
                                        len(chromlist), ptr(bc), ptr(ba), nb, int(window), ptr(off), ptr(rows), cap, 0, ptr(loci)),
              "coex_permute_map")
        out = [rows[off[k]:off[k + 1]].copy() for k in range(K)]
        if with_loci:
            return out, [loci[2 * off[k]:2 * off[k + 1]].reshape(-1, 2).copy() for k in range(K)]
        return out

    def set_features_raw(self, chrom_id, start, end, name_rank):
        chrom_id = np.ascontiguousarray(chrom_id, dtype=np.int32)
        start = np.ascontiguousarray(start, dtype=np.int64)
        end = np.ascontiguousarray(end, dtype=np.int64)
        name_rank = np.ascontiguousarray(name_rank, dtype=np.int32)
        check(self.lib.coex_set_features(self.h, chrom_id.ctypes.data, start.ctypes.data, end.ctypes.data,
                                         name_rank.ctypes.data, len(chrom_id)), "coex_set_features")

    def set_feature_order(self, bed_index):
        """bed_index[row] = line number of the row's feature in the feature BED (the B file of windowBed): `bedtools window`
        reports the features of one locus by (UCSC-bin level, bin, line), which fixes the order of a permutation's promoters.
        Default (after set_features): the BED lists the rows in table order."""
        b = np.ascontiguousarray(bed_index, dtype=np.int32)
        check(self.lib.coex_set_feature_order(self.h, b.ctypes.data, b.size), "coex_set_feature_order")

    def set_global_list(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        check(self.lib.coex_set_global_list(self.h, v.ctypes.data, v.size), "coex_set_global_list")

    # ---- A1 on the resident table ----------------------------------------------------------------
    def pairs(self, idxa, idxb, ranked):
        idxa = np.ascontiguousarray(idxa, dtype=np.int32)
        idxb = np.ascontiguousarray(idxb, dtype=np.int32)
        out = np.empty(idxa.size, dtype=np.float64)
        check(self.lib.coex_pairs(self.h, 1 if ranked else 0, idxa.ctypes.data, idxb.ctypes.data, idxa.size,
                                  out.ctypes.data, 0), "coex_pairs")
        return out

    def allpairs_accumulate(self, ranked, row0, row1, nbins=1 << 20, reset=False):
        """Adds the pairs (i < j) with i in [row0, row1) to the device-resident histogram; nothing is read back."""
        check(self.lib.coex_allpairs_accumulate(self.h, 1 if ranked else 0, row0, row1, nbins, 1 if reset else 0),
              "coex_allpairs_accumulate")

    def allpairs_read(self, nbins=1 << 20, device_ptr=None, host=True):
        """-> (hist uint64 [nbins], nan pairs) of everything accumulated; device_ptr (nbins + 1 int64 counters) gets a copy."""
        hist = np.zeros(nbins, dtype=np.uint64) if host else None
        nan = np.zeros(1, dtype=np.uint64)
        check(self.lib.coex_allpairs_read(self.h, hist.ctypes.data if host else None, nan.ctypes.data, device_ptr),
              "coex_allpairs_read")
        return hist, int(nan[0])

    def allpairs_histogram(self, ranked, row0=0, row1=None, nbins=1 << 20):
        row1 = self.N if row1 is None else row1
        hist = np.zeros(nbins, dtype=np.uint64)
        nan = np.zeros(1, dtype=np.uint64)
        check(self.lib.coex_allpairs_histogram(self.h, 1 if ranked else 0, row0, row1, nbins, hist.ctypes.data,
                                               nan.ctypes.data), "coex_allpairs_histogram")
        return hist, int(nan[0])

    def corr_block(self, which, row0, nrows):
        """which: 'spearman' | 'pearson' (tensor cores) | 'pearson_exact' -> float64 [nrows, N]."""
        out = np.empty((nrows, self.N), dtype=np.float64)
        code = {"spearman": 1, "pearson": 0, "pearson_exact": 2}[which]
        check(self.lib.coex_corr_block(self.h, code, int(row0), int(nrows), out.ctypes.data), "coex_corr_block")
        return out

    def set_pearson(self, digits=4, mode="tensor"):
        check(self.lib.coex_set_pearson(self.h, int(digits), {"tensor": 0, "exact": 1}[mode]), "coex_set_pearson")

    # ---- A4-A12 batched -----------------------------------------------------------------------------
    def network_batch(self, hit_sets, options: Options, want_scores=False, want_counts=False):
        """hit_sets: list of int arrays (table rows, map-file order), one per permutation."""
        K = len(hit_sets)
        perm_off = np.zeros(K + 1, dtype=np.int64)
        perm_off[1:] = np.cumsum([len(h) for h in hit_sets])
        H = int(perm_off[-1])
        hit_rows = (np.concatenate([np.asarray(h, dtype=np.int32) for h in hit_sets])
                    if H else np.zeros(0, dtype=np.int32))
        hit_rows = np.ascontiguousarray(hit_rows, dtype=np.int32)
        Hn = max(H, 1)
        n_groups = np.zeros(max(K, 1), dtype=np.int32)
        group_of_hit = np.full(Hn, -1, dtype=np.int32)
        group_label = np.full(Hn, -1, dtype=np.int32)
        group_winner = np.full(Hn, -1, dtype=np.int32)
        group_density = np.zeros(Hn, dtype=np.float64)
        hit_score = np.zeros(Hn, dtype=np.float64)
        SC = int(np.sum(np.diff(perm_off) ** 2))
        scores = np.zeros(max(SC, 1), dtype=np.float64) if want_scores else None
        counts = np.full(max(SC, 1), -1, dtype=np.int64) if want_counts else None
        prm = options.to_params(self.N)
        check(self.lib.coex_network_batch(self.h, ctypes.byref(prm), K, perm_off.ctypes.data, _ptr(hit_rows),
                                          n_groups.ctypes.data, group_of_hit.ctypes.data, group_label.ctypes.data,
                                          group_winner.ctypes.data, group_density.ctypes.data,
                                          hit_score.ctypes.data, _ptr(scores), _ptr(counts), 0),
              "coex_network_batch")
        return BatchResult(perm_off, hit_rows, n_groups[:K], group_of_hit[:H], group_label[:H], group_winner[:H],
                           group_density[:H], hit_score[:H], scores, counts)

    def network_batch_device(self, options: Options, K, perm_off_host, ptrs):
        """All buffers already on the device (ptrs: dict of device pointers); asynchronous on
        the context stream apart from the final stream sync the C entry point performs."""
        prm = options.to_params(self.N)
        perm_off_host = np.ascontiguousarray(perm_off_host, dtype=np.int64)
        g = lambda k: ctypes.c_void_p(ptrs[k]) if ptrs.get(k) else None
        check(self.lib.coex_network_batch(self.h, ctypes.byref(prm), int(K), perm_off_host.ctypes.data,
                                          g("hit_rows"), g("n_groups"), g("group_of_hit"), g("group_label"),
                                          g("group_winner"), g("group_density"), g("hit_score"), g("scores"),
                                          g("counts"), 1), "coex_network_batch")

    # ---- ONE job shared by several GPUs (include/coexpression_b200.h section 3; dist.SharedJob drives these) ----------
    def shard_table_export(self, world, rank):
        """After set_expression: -> the 9 x 64 bytes of CUDA IPC handles of this rank's packed-table buffers."""
        h = (ctypes.c_ubyte * (9 * 64))()
        check(self.lib.coex_shard_table_export(self.h, int(world), int(rank), h), "coex_shard_table_export")
        return bytes(h)

    def shard_table_open(self, peer_rank, handles):
        buf = (ctypes.c_ubyte * (9 * 64)).from_buffer_copy(bytes(handles))
        check(self.lib.coex_shard_table_open(self.h, int(peer_rank), buf), "coex_shard_table_open")

    def shard_prepare_rows(self):
        check(self.lib.coex_shard_prepare_rows(self.h), "coex_shard_prepare_rows")

    def shard_prepare_finish(self):
        check(self.lib.coex_shard_prepare_finish(self.h), "coex_shard_prepare_finish")

    def shard_scores_alloc(self, world, rank, n_doubles):
        """-> the 64-byte CUDA IPC handle of this rank's score buffer."""
        h = (ctypes.c_ubyte * 64)()
        check(self.lib.coex_shard_scores_alloc(self.h, int(world), int(rank), int(n_doubles), h), "coex_shard_scores_alloc")
        return bytes(h)

    def shard_scores_open(self, peer_rank, handle):
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(bytes(handle))
        check(self.lib.coex_shard_scores_open(self.h, int(peer_rank), buf), "coex_shard_scores_open")

    @staticmethod
    def _flatten(hit_sets):
        K = len(hit_sets)
        perm_off = np.zeros(K + 1, dtype=np.int64)
        perm_off[1:] = np.cumsum([len(h) for h in hit_sets])
        H = int(perm_off[-1])
        hit_rows = (np.concatenate([np.asarray(h, dtype=np.int32) for h in hit_sets]) if H else np.zeros(0, dtype=np.int32))
        return K, perm_off, np.ascontiguousarray(hit_rows, dtype=np.int32), H

    def shard_count(self, hit_sets_all, options: Options, perm_owner, perm_score_off):
        """Count stage of THIS rank's table rows over ALL permutations of the job; the score rows are stored into the buffers of
        the permutations' owners (peer memory)."""
        K, perm_off, hit_rows, _ = self._flatten(hit_sets_all)
        owner = np.ascontiguousarray(perm_owner, dtype=np.int32)
        soff = np.ascontiguousarray(perm_score_off, dtype=np.int64)
        prm = options.to_params(self.N)
        check(self.lib.coex_shard_count(self.h, ctypes.byref(prm), K, perm_off.ctypes.data, _ptr(hit_rows), owner.ctypes.data,
                                        soff.ctypes.data), "coex_shard_count")

    def shard_finish(self, hit_sets_mine, options: Options, want_scores=False):
        """Network stage of THIS rank's permutations (their score blocks are complete after the caller's barrier)."""
        K, perm_off, hit_rows, H = self._flatten(hit_sets_mine)
        Hn = max(H, 1)
        n_groups = np.zeros(max(K, 1), dtype=np.int32)
        group_of_hit = np.full(Hn, -1, dtype=np.int32)
        group_label = np.full(Hn, -1, dtype=np.int32)
        group_winner = np.full(Hn, -1, dtype=np.int32)
        group_density = np.zeros(Hn, dtype=np.float64)
        hit_score = np.zeros(Hn, dtype=np.float64)
        SC = int(np.sum(np.diff(perm_off) ** 2))
        scores = np.zeros(max(SC, 1), dtype=np.float64) if want_scores else None
        prm = options.to_params(self.N)
        check(self.lib.coex_shard_finish(self.h, ctypes.byref(prm), K, perm_off.ctypes.data, _ptr(hit_rows),
                                         n_groups.ctypes.data, group_of_hit.ctypes.data, group_label.ctypes.data,
                                         group_winner.ctypes.data, group_density.ctypes.data, hit_score.ctypes.data,
                                         _ptr(scores), 0), "coex_shard_finish")
        return BatchResult(perm_off, hit_rows, n_groups[:K], group_of_hit[:H], group_label[:H], group_winner[:H],
                           group_density[:H], hit_score[:H], scores, None)

    def read_scores(self, k, P, out=None):
        """individualscores of permutation k of the last reduced batch, as a [P, P] array (the batch itself only returns
        densities unless want_scores: the permuted sets' score matrices are never needed on the host).  out: a C-contiguous
        float64 array of P * P elements to fill (e.g. a view of pinned memory: a pageable destination copies ~10 x slower)."""
        if out is None:
            out = np.empty((int(P), int(P)), dtype=np.float64)
        assert out.dtype == np.float64 and out.size == int(P) * int(P) and out.flags["C_CONTIGUOUS"]
        check(self.lib.coex_read_scores(self.h, int(k), out.ctypes.data), "coex_read_scores")
        return out.reshape(int(P), int(P))

    def set_score_readback(self, k, out):
        """network_batch itself copies the score block of permutation k into `out` (float64, P_k^2 elements, ideally pinned)
        while the network stage runs; out=None switches it off."""
        self._rb = out
        check(self.lib.coex_set_score_readback(self.h, int(k), out.ctypes.data if out is not None else None),
              "coex_set_score_readback")

    # ---- A14: pooled null, empirical p, FDR (2-collate-results.py:67-76,142,161-163) on the device ----------------------
    def collate(self, pool, realmeasures, pool_device_ptr=None, n_pool=None):
        """pool: host array of every permuted density (or pool_device_ptr + n_pool: a device buffer, e.g. the gathered block);
        -> dict(p ascending, fdr_bh, fdr_by) as collate.collate_numbers computes them on the host."""
        real = np.ascontiguousarray(realmeasures, dtype=np.float64)
        G = real.size
        p, bh, by = np.zeros(max(G, 1)), np.zeros(max(G, 1)), np.zeros(max(G, 1))
        if pool_device_ptr is not None:
            src, npool, on_dev = ctypes.c_void_p(pool_device_ptr), int(n_pool), 1
        else:
            pool = np.ascontiguousarray(pool, dtype=np.float64)
            src, npool, on_dev = (pool.ctypes.data if pool.size else None), int(pool.size), 0
        check(self.lib.coex_collate(self.h, src, npool, on_dev, real.ctypes.data if G else None, G, p.ctypes.data, bh.ctypes.data,
                                    by.ctypes.data), "coex_collate")
        return {"p": p[:G].tolist(), "fdr_bh": bh[:G].tolist(), "fdr_by": by[:G].tolist()}

    # ---- introspection ------------------------------------------------------------------------------
    def stage_ms(self):
        out = {}
        for i, name in enumerate(STAGES):
            ms = ctypes.c_double()
            ln = ctypes.c_longlong()
            check(self.lib.coex_stage_ms(self.h, i, ctypes.byref(ms), ctypes.byref(ln)), "coex_stage_ms")
            out[name] = (ms.value, ln.value)
        return out

    def launch_count(self):
        return int(self.lib.coex_launch_count(self.h))

    def stream_ptr(self):
        return self.lib.coex_stream(self.h)

    def sync(self):
        check(self.lib.coex_sync(self.h), "coex_sync")

    def set_gemm_mode(self, mode):
        check(self.lib.coex_set_gemm_mode(self.h, {"tcgen05": 0, "simt": 1, "tcgen05_bn64": 64, "tcgen05_bn128": 128, "tcgen05_persistent": 2, "tcgen05_persistent128": 3, "tcgen05_cluster128": 5, "tcgen05_split": 6}.get(mode, mode)),
              "coex_set_gemm_mode")

    def set_count_mode(self, mode):
        check(self.lib.coex_set_count_mode(self.h, {"dedup": 0, "per_query": 1, "dedup_sorted": 2, "row_rank": 3, "dedup_fine": 4}.get(mode, mode)), "coex_set_count_mode")

    def set_workspace_mb(self, mb):
        check(self.lib.coex_set_workspace_mb(self.h, int(mb)), "coex_set_workspace_mb")

    def selftest_gemm(self, m_rows):
        bad = ctypes.c_longlong()
        first = (ctypes.c_longlong * 4)()
        check(self.lib.coex_selftest_gemm(self.h, int(m_rows), ctypes.byref(bad), first), "coex_selftest_gemm")
        return int(bad.value), [int(x) for x in first]
