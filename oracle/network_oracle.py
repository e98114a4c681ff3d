"""oracle/network_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement, in Python 3 + numpy, of the reference's network-density hot
path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package `coexpression_b200` never does.

What is restated (reference file:line):
  A1  getPearson ............ coexpression_v2.c:22-129  -> called through ctypes, either the
                              UNMODIFIED reference build oracle/_ref/coexpression_v2.so
                              (kind "reference") or our C port oracle/_ref/liboracle.so ("port")
  A2  rank transform ........ 1-make-network.py:108-111, coexfunctions.py:428-432
                              (scipy.stats.rankdata, method='average')
  A3  matrix marshal ........ 1-make-network.py:100-122 (row order = expression-file order)
  A4  pair list + Q1 skip ... 1-make-network.py:198-226
  A5  NaN->-99, sorted null . 1-make-network.py:239-258
  A6  get_correlation ....... coexfunctions.py:386-409
  A7  edge scores ........... 1-make-network.py:273-327
  A8  join_nearby ........... coexfunctions.py:277-316, merge_ranges :412-425
  A9  winners ............... 1-make-network.py:365-434
  A10 indjoined ............. 1-make-network.py:437-452
  A11 indmeasure ............ 1-make-network.py:456-462
  A12 iterative removal ..... 1-make-network.py:464-482
  A14 pooled null, p, FDR ... 2-collate-results.py:67-76,142,161-163;
                              coexfunctions.py:379-383 (empiricalp), :333-358 (fdr_correction)

Parity status
  * A1 is PINNED: the port is compared bit-for-bit with the unmodified reference C build
    (tests/test_oracle.py) and with tests/golden/getpearson_*.npz generated from it.
  * A2 is PINNED against scipy.stats.rankdata (the reference's own third-party call,
    scipy 0.19.0 in environment.yml:58; `method='average'` semantics are unchanged in the
    scipy 1.18 installed here).
  * A3-A13 are PINNED against the reference's own Python EXECUTED in the build container:
    the reference ships no tests or fixtures and its drivers are Python 2,
    but `coexfunctions.py` imports unmodified under Python 3 (matplotlib stubbed) and
    `1-make-network.py` runs after a mechanical, line-local 2 -> 3 rewrite of its text
    (tests/golden/run_reference_py3.py: print statements, iteritems, xrange, keys(), one tuple
    lambda, ConfigParser; Python 2.7's plain `sum` restored), driving the unmodified reference C
    build.  tests/golden/make_reference_golden.py committed its seven result objects for seven
    option sets x two hit sets; the *faithful* mode below reproduces every one of them BIT FOR
    BIT, dict key orders included, when given the two dict orders that run had
    (`group_order`, `indjoined_order="insertion"`; tests/test_reference_pin.py).
  * A8 additionally against a pandas/networkx transliteration (tests/test_oracle.py); A14's
    `empiricalp` / `fdr_correction` and `merge_ranges` against direct calls of the reference's
    functions (tests/golden/reference_functions.json).
  * the fixtures of tests/golden/make_golden.py (larger shapes) come from THIS restatement
    driving the reference C build.

Two modes:
  faithful   -- the hit-vs-hit statistic comes from scipy.stats.spearmanr / pearsonr exactly
                where the reference calls them (coexfunctions.py:390-399).
  consistent -- the statistic is the number getPearson returns for the same pair (so the
                statistic is bitwise equal to its twin inside the null list).  Parity of the
                GPU path is gated on this mode; faithful-vs-consistent differences are a
                documented property of the reference (1-ulp scipy-vs-C noise moving bisect
                indices by one), not of the replacement.

Ordering decisions (the reference iterates Python-2 dicts in hash order, which is not
reproducible; every order is made explicit here and the GPU path uses the same ones):
  rows ........ expression-table order (Q1)
  promoters ... first appearance in the map file (Q1)
  groups ...... ascending group label (string order); members ascending name (Q6)
  winner tie .. first maximum in ascending member-name order (Q6)
  pass 1 ...... each member's sum runs over the other hits in hit (map-file) order
  pass 2 ...... sequential over groups in group order, winners updated in place (Q6)
  sums ........ plain left-to-right double accumulation in group order (Python 2 `sum`;
                NOT Python >= 3.12's compensated `sum`)
"""
from __future__ import annotations

import ctypes
import math
import os
from bisect import bisect_left

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "coexpression_v2.so")
_PORT_SO = os.path.join(_HERE, "_ref", "liboracle.so")

_libs = {}


def available_kinds():
    kinds = []
    if os.path.exists(_REF_SO):
        kinds.append("reference")
    if os.path.exists(_PORT_SO):
        kinds.append("port")
    return kinds


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    if kind == "reference":
        lib = ctypes.CDLL(_REF_SO)
        fn = lib.getPearson
    elif kind == "port":
        lib = ctypes.CDLL(_PORT_SO)
        fn = lib.oracle_getPearson
    else:
        raise ValueError(kind)
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_int, ctypes.c_int]
    _libs[kind] = (lib, fn)
    return _libs[kind]


def default_kind():
    kinds = available_kinds()
    if not kinds:
        raise RuntimeError("oracle not built: run `make -C oracle` (see oracle/Makefile)")
    return kinds[0]


def num_threads(kind="port"):
    lib, _ = _load(kind if kind == "port" and os.path.exists(_PORT_SO) else "port")
    lib.oracle_num_threads.restype = ctypes.c_int
    return int(lib.oracle_num_threads())


def get_pearson(data, idxa, idxb, kind=None):
    """A1.  data: C-contiguous float64 [N, n]; idxa/idxb int32 [npairs] -> float64 [npairs].

    Calls the library exactly as 1-make-network.py:229-234 does (four buffers, two ints).
    """
    kind = kind or default_kind()
    _, fn = _load(kind)
    data = np.ascontiguousarray(data, dtype=np.float64)
    idxa = np.ascontiguousarray(idxa, dtype=np.int32)
    idxb = np.ascontiguousarray(idxb, dtype=np.int32)
    assert idxa.shape == idxb.shape and data.ndim == 2
    out = np.empty(idxa.shape[0], dtype=np.float64)
    if idxa.shape[0] == 0:
        return out
    fn(data.ctypes.data, out.ctypes.data, idxa.ctypes.data, idxb.ctypes.data,
       int(idxa.shape[0]), int(data.shape[1]))
    return out


# --------------------------------------------------------------------------------------
# A2  rank transform
# --------------------------------------------------------------------------------------
def rank_rows(X):
    """scipy.stats.rankdata per row, default method='average' (1-make-network.py:110-111)."""
    from scipy import stats
    X = np.asarray(X, dtype=np.float64)
    return np.vstack([stats.rankdata(X[i]) for i in range(X.shape[0])]) if X.shape[0] else X.copy()


def py2_sum(values):
    """Python 2 `sum` of floats: plain left-to-right double accumulation."""
    s = 0
    for v in values:
        s = s + v
    return s


# --------------------------------------------------------------------------------------
# A6  get_correlation
# --------------------------------------------------------------------------------------
class Correlator:
    """get_correlation (coexfunctions.py:386-409) in faithful or consistent mode."""

    def __init__(self, X, Xrank, mode="consistent", kind=None):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.Xrank = np.ascontiguousarray(Xrank, dtype=np.float64)
        self.mode = mode
        self.kind = kind
        # `sum(ed[pa]) > 0` guard, coexfunctions.py:389 (left-to-right double sum)
        self.rawsum_pos = np.array([py2_sum(row.tolist()) > 0 for row in self.X], dtype=bool)

    def many(self, ia, ib, cm="Spearman"):
        ia = np.asarray(ia, dtype=np.int32)
        ib = np.asarray(ib, dtype=np.int32)
        out = np.full(ia.shape[0], -99.0)
        ok = self.rawsum_pos[ia] & self.rawsum_pos[ib]
        if not ok.any():
            return out
        if self.mode == "consistent":
            src = self.Xrank if cm == "Spearman" else self.X
            r = get_pearson(src, ia[ok], ib[ok], self.kind)
        else:
            from scipy import stats
            r = np.empty(int(ok.sum()))
            for k, (a, b) in enumerate(zip(ia[ok], ib[ok])):
                try:
                    if cm == "Pearson":
                        r[k] = stats.pearsonr(self.X[a], self.X[b])[0]
                    else:
                        r[k] = stats.spearmanr(self.X[a], self.X[b])[0]
                except Exception:
                    r[k] = -99.0
        r = np.where(np.isnan(r), -99.0, r)          # coexfunctions.py:406-408
        out[ok] = r
        return out


# --------------------------------------------------------------------------------------
# A8  join_nearby / merge_ranges
# --------------------------------------------------------------------------------------
def merge_ranges(ranges, ss=100000):
    """coexfunctions.py:412-425.  `ranges` = [(start, end), ...] already sorted."""
    it = iter([(s, e + ss) for s, e in ranges])
    cur_start, cur_stop = next(it)
    for start, stop in it:
        if start > cur_stop:
            yield cur_start, cur_stop
            cur_start, cur_stop = start, stop
        else:
            cur_stop = max(cur_stop, stop)
    yield cur_start, cur_stop


def join_nearby(proms, addresses, globcors, corr, row_of, pvtj=0.1, cm="Spearman",
                softsep=100000, anticor=False):
    """coexfunctions.py:277-316 without pandas/networkx.

    proms: iterable of names; addresses: name -> (chrom, start, end); globcors: ascending
    list; corr: Correlator; row_of: name -> matrix row.  Returns {label: sorted members}
    with keys in ascending label order.
    """
    proms = list(proms)
    if not proms:
        return {}
    ordered = sorted(proms, key=lambda k: (addresses[k][0], addresses[k][1], addresses[k][2]))
    by_chrom = {}
    for k in ordered:
        by_chrom.setdefault(addresses[k][0], []).append(k)
    clusters = []
    for chrom, members in by_chrom.items():
        rng = [(addresses[k][1], addresses[k][2]) for k in members]
        for g0, g1 in merge_ranges(rng, softsep):
            clusters.append([k for k in members if g0 <= addresses[k][1] < g1])
    # all in-cluster pairs i<j, one correlation call
    pa, pb = [], []
    for ci, cl in enumerate(clusters):
        for i in range(len(cl)):
            for j in range(i + 1, len(cl)):
                pa.append(cl[i]); pb.append(cl[j])
    r = corr.many([row_of[a] for a in pa], [row_of[b] for b in pb], cm) if pa else []
    parent = {k: k for k in proms}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    L = len(globcors)
    for a, b, rv in zip(pa, pb, r):
        idx = bisect_left(globcors, rv)
        p1 = 1 - float(idx) / L
        if anticor:
            p2 = float(idx) / L
            p = min(p1, p2)
        else:
            p = p1
        if p < pvtj:
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[ra] = rb
    comps = {}
    for k in proms:
        comps.setdefault(find(k), []).append(k)
    ptj = {}
    for members in comps.values():
        new = sorted(members)
        ptj[new[0]] = new
    return {k: ptj[k] for k in sorted(ptj)}


# --------------------------------------------------------------------------------------
# A3-A12  one permutation (1-make-network.py:100-482)
# --------------------------------------------------------------------------------------
def edge_score(idx, length, anticor):
    """1-make-network.py:284-314 for a valid (non -99) statistic."""
    p1 = 1 - float(idx) / length
    if anticor:
        p2 = float(idx) / length
        p1 = min(p1, p2)
    try:
        return -math.log(p1, 10)
    except Exception:
        return 0


def make_network(names, X, promoters, addresses, globalcorrelationlist,
                 correlationmeasure="Spearman", anticorrelation=False, noiter=False,
                 pval_to_join=0.1, soft_separation=100000, nsp=1.0, useaverage=False,
                 selectionthreshold=0.0, selectionthreshold_is_j=False,
                 mode="consistent", kind=None, Xrank=None, return_counts=False, group_order=None,
                 indjoined_order="groups"):
    """One permutation.  names: row labels in table order; X: float64 [N, n] raw table;
    promoters: hit labels in map-file first-appearance order (all present in `names`).
    group_order: the order in which the groups are visited (quirk Q6: the second winner pass depends on it; the
    reference takes whatever order its `prom_to_join` dict has).  Default: ascending label, the order this repo fixes
    for both sides; tests/test_reference_pin.py passes the order an executed reference run used.
    indjoined_order: order of the addends of a density (`indjoined[g]` is summed in its dict order, :456-482).
    "groups" = the group order (what this repo fixes for both sides); "insertion" = the order in which the loops at
    :437-452 insert the keys into an insertion-ordered dict, i.e. what the reference does under Python 3.
    Returns the objects 1-make-network.py:523-532 dumps (without snp_map)."""
    names = list(names)
    X = np.ascontiguousarray(X, dtype=np.float64)
    N = len(names)
    row_of = {nm: i for i, nm in enumerate(names)}
    if Xrank is None:
        Xrank = rank_rows(X)
    data_c = Xrank if correlationmeasure == "Spearman" else X            # :118-121
    if selectionthreshold_is_j:                                           # :95-96
        selectionthreshold = -math.log(pval_to_join, 10)
    corr = Correlator(X, Xrank, mode, kind)
    promoters = list(promoters)
    P = len(promoters)
    hit_rows = np.array([row_of[p] for p in promoters], dtype=np.int32)

    use_specific_p = not (nsp <= 1.0 / N)                                 # :189-193
    null = None
    if use_specific_p:                                                    # :198-258
        if nsp < 1:
            raise NameError("name 'random' is not defined (reference quirk Q8)")
        cols = np.arange(N, dtype=np.int32)
        ia = np.repeat(hit_rows, N)
        ib = np.tile(cols, P)
        keep = np.repeat(np.arange(P), N) != np.tile(np.arange(N), P)     # Q1: i != j
        r = get_pearson(data_c, ia[keep], ib[keep], kind) if P else np.empty(0)
        r = np.where(np.isnan(r), -99.0, r)                               # :240-246
        null = []
        pos = 0
        for i in range(P):
            ln = N - 1 if i < N else N
            null.append(np.sort(r[pos:pos + ln]))                         # :252-258
            pos += ln

    # hit-vs-hit statistic, always Spearman on the raw table (Q2, :281)
    ii, jj = np.meshgrid(np.arange(P), np.arange(P), indexing="ij")
    off = ii != jj
    stat = np.full((P, P), np.nan)
    if P > 1:
        stat[off] = corr.many(hit_rows[ii[off]], hit_rows[jj[off]], "Spearman")

    individualscores = {}
    counts = np.full((P, P), -1, dtype=np.int64)
    for i in range(P):                                                    # :273-327
        for j in range(P):
            if i == j:
                continue
            c = stat[i, j]
            if use_specific_p:
                if c == -99:
                    score = -math.log(1, 10)
                else:
                    lst = null[i]
                    idx = int(np.searchsorted(lst, c, side="left"))       # bisect_left
                    counts[i, j] = idx
                    score = edge_score(idx, len(lst), anticorrelation)
            else:
                if c == -99:
                    score = -math.log(1, 10)
                else:
                    idx = bisect_left(globalcorrelationlist, c)
                    counts[i, j] = idx
                    score = edge_score(idx, len(globalcorrelationlist), anticorrelation)
            individualscores.setdefault(promoters[i], {})[promoters[j]] = score
            if not nsp:                                                   # :321-326
                individualscores.setdefault(promoters[j], {})[promoters[i]] = score

    all_proms = []                                                        # :333-338
    for prom in individualscores:
        all_proms.append(prom)
        for other in individualscores[prom]:
            all_proms.append(other)
    all_proms = sorted(set(all_proms))

    prom_to_join = join_nearby(all_proms, addresses, globalcorrelationlist, corr, row_of,
                               pval_to_join, correlationmeasure, soft_separation,
                               anticorrelation)                          # :347-348
    if group_order is not None:
        assert sorted(group_order) == sorted(prom_to_join), "group_order must list exactly the group labels"
        prom_to_join = {k: prom_to_join[k] for k in group_order}
    joining_instructions = {}
    for label in prom_to_join:
        for prom in prom_to_join[label]:
            joining_instructions[prom] = label

    winners = {}
    allpromoterscores = {}
    for label in prom_to_join:                                            # pass 1 :365-375
        best, best_score = None, None
        for prom in prom_to_join[label]:
            if joining_instructions[prom] != label:
                continue
            # iteration order of individualscores[prom] = insertion order = hit order j
            s = py2_sum([individualscores[prom][o] for o in individualscores[prom]
                         if individualscores[prom][o] >= selectionthreshold
                         and joining_instructions[o] != label])
            if best is None or s > best_score:
                best, best_score = prom, s
        if best is not None:
            winners[label] = best
    for label in prom_to_join:                                            # pass 2 :378-434
        best, best_score = None, None
        for prom in prom_to_join[label]:
            if joining_instructions[prom] != label:
                continue
            this = [individualscores[prom][o] for o in winners.values()
                    if prom != o and joining_instructions[o] != label]
            s = py2_sum(this)
            allpromoterscores[prom] = s
            if best is None or s > best_score:
                best, best_score = prom, s
        if best is not None:
            winners[label] = best

    indjoined = {}                                                        # :437-452
    if indjoined_order == "insertion":
        for prom in individualscores:
            label = joining_instructions[prom]
            chosen = winners[label]
            if label not in indjoined:
                indjoined[label] = {}
            for other_prom in individualscores[prom]:
                other = joining_instructions[other_prom]
                other_chosen = winners[other]
                if chosen != other_chosen:
                    indjoined[label][other] = individualscores[chosen][other_chosen]
    else:
      for label in prom_to_join:
          if winners.get(label) is None:
              continue
          indjoined[label] = {}
          for other in prom_to_join:
              if other == label:
                  continue
              chosen, other_chosen = winners[label], winners[other]
              indjoined[label][other] = individualscores[chosen][other_chosen]

    indmeasure = {}                                                       # :456-462
    for g in indjoined:
        s = py2_sum([indjoined[g][h] for h in indjoined[g] if indjoined[g][h] >= selectionthreshold])
        indmeasure[g] = s / len(indjoined) if useaverage else s

    if not noiter:                                                        # :464-482
        whittled = {}
        cut = set()
        lastscore, lastresult = -1, None
        for g, score in sorted(indmeasure.items(), key=lambda kv: (kv[1], kv[0]), reverse=True):
            if score == lastscore:
                whittled[g] = lastresult
            else:
                s = py2_sum([indjoined[g][h] for h in indjoined[g]
                             if indjoined[g][h] >= selectionthreshold and h not in cut])
                whittled[g] = s / len(indjoined) if useaverage else s
            cut.add(g)
            lastscore = score
            lastresult = whittled[g]
        indmeasure = whittled

    out = {
        "indmeasure": indmeasure,
        "individualscores": individualscores,
        "prom_to_join": prom_to_join,
        "indjoined": indjoined,
        "joining_instructions": joining_instructions,
        "winners": winners,
        "allpromoterscores": allpromoterscores,
    }
    if return_counts:
        out["_counts"] = counts
        out["_stat"] = stat
    return out


def query_counts(X, Xrank, hit_rows, i, kind=None, anticorrelation=False):
    """ONE hit of one permutation: row i of the bisect-index matrix and of `individualscores`
    (1-make-network.py:208-258 and :273-327 restricted to hit position i, Spearman, nsp = 1,
    consistent mode).  make_network above does exactly this for every i; this entry exists so that
    a SAMPLE of queries can be checked at shapes where a whole permutation is too slow on the CPU
    (bench.py, full FANTOM5 shape).  X: raw table (only the rows of the hits are read: raw-sum
    guard of coexfunctions.py:389), Xrank: ranked table.  -> (counts [P] int64, -1 where the
    statistic is -99 or j == i; scores [P] float64, 0 at j == i)."""
    hit_rows = np.asarray(hit_rows, dtype=np.int32)
    N, P = Xrank.shape[0], len(hit_rows)
    u = int(hit_rows[i])
    cols = np.arange(N, dtype=np.int32)
    keep = cols != i                                                      # Q1: the INDEX i, not the row u
    r = get_pearson(Xrank, np.full(int(keep.sum()), u, dtype=np.int32), cols[keep], kind)
    lst = np.sort(np.where(np.isnan(r), -99.0, r))                        # :240-258
    pos = np.array([py2_sum(X[int(h)].tolist()) > 0 for h in hit_rows], dtype=bool)
    counts = np.full(P, -1, dtype=np.int64)
    scores = np.zeros(P)
    others = np.array([j for j in range(P) if j != i], dtype=np.int64)
    t = get_pearson(Xrank, np.full(len(others), u, dtype=np.int32), hit_rows[others], kind) if len(others) else np.empty(0)
    for j, c in zip(others, t):
        c = -99.0 if (np.isnan(c) or not (pos[i] and pos[j])) else float(c)
        if c == -99:
            scores[j] = -math.log(1, 10)
        else:
            idx = int(np.searchsorted(lst, c, side="left"))
            counts[j] = idx
            scores[j] = edge_score(idx, len(lst), anticorrelation)
    return counts, scores


# --------------------------------------------------------------------------------------
# A14  pooled null, empirical p, FDR (2-collate-results.py:67-76,142,161-163)
# --------------------------------------------------------------------------------------
def empiricalp(value, collection_sorted):
    """coexfunctions.py:379-383 (collection must already be ascending)."""
    if len(collection_sorted) == 0:
        return 1
    return 1 - float(bisect_left(collection_sorted, value)) / len(collection_sorted)


def fdr_correction(pvals, alpha=0.05, method="indep"):
    """coexfunctions.py:333-358 restated (Benjamini-Hochberg 'indep', -Yekutieli 'negcorr')."""
    p = np.asarray(pvals, dtype=np.float64).ravel()
    m = p.size
    if m == 0:
        return np.zeros(0, dtype=bool), np.zeros(0)
    order = np.argsort(p)
    ps = p[order]
    frac = np.arange(1, m + 1) / float(m)
    if method in ("i", "indep", "p", "poscorr"):
        pass
    elif method in ("n", "negcorr"):
        frac = frac / py2_sum((1.0 / np.arange(1, m + 1)).tolist())
    else:
        raise ValueError("Method should be 'indep' and 'negcorr'")
    reject = ps < frac * alpha
    if reject.any():
        reject[:np.nonzero(reject)[0].max()] = True
    adj = np.minimum.accumulate((ps / frac)[::-1])[::-1]
    adj[adj > 1.0] = 1.0
    inv = np.argsort(order)
    return reject[inv], adj[inv]


def collate(real_indmeasure, permuted_indmeasures):
    """real_indmeasure: {group: density}; permuted_indmeasures: list of such dicts (k != 0).
    Returns (sorted empirical p list, BY-adjusted, BH-adjusted) as 2-collate-results.py:161-163."""
    pool = []
    for d in permuted_indmeasures:
        pool += list(d.values())
    pool.sort()
    real = list(real_indmeasure.values())
    ps = sorted(empiricalp(x, pool) for x in real)
    by = list(fdr_correction(ps, 0.05, "negcorr")[1])
    bh = list(fdr_correction(ps)[1])
    return ps, by, bh
