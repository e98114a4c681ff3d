"""Seeded synthetic inputs shaped like the reference's data files.

Nothing here is arithmetic of the hot path; it only manufactures inputs:
  * an expression table  (format of `read_expression_file`, reference coexfunctions.py:201-233):
    N region labels `chr<c>:<start>..<end>,<strand>` x n samples, zero-inflated heavy-tailed
    non-negative values with low-rank "tissue programme" structure, a few all-zero rows
    (NaN / -99 path, coexpression_v2.c:89-92) and a few duplicated rows (exact r = 1 ties);
  * a feature BED        (format of `readfeaturecoordinates`, coexfunctions.py:235-274);
  * windowBed-style map files for the real loci and for circularly shifted loci
    (consumer: 1-make-network.py:131-164; producer restated from 0-prepare-coex.py:471-613
    only as far as needed to get realistic hit sets);
  * random cross-chromosome pair indices for the global correlation list
    (0-prepare-coex.py:672-687).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# hg19-like chromosome lengths (Mb), enough for realistic spacing
_CHROM_MB = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103,
             90, 81, 78, 59, 63, 48, 51, 155, 59]
_CHROM_NAMES = [str(i) for i in range(1, 23)] + ["X", "Y"]

SHAPES = {
    # name: (N rows, n samples).  "minimal" is an ASSUMPTION: the reference's
    # pure_f5p_condense_ptt_minimal_av.expression is an external file whose shape is not
    # recorded anywhere in the reference.
    "tiny": (600, 48),
    "small": (3000, 96),
    "minimal": (20000, 256),
    "full": (200000, 1829),
}


@dataclass
class Table:
    names: list                      # row labels, table order
    X: np.ndarray                    # float64 [N, n]
    chrom: np.ndarray                # str per row, e.g. "chr7"
    start: np.ndarray                # int64
    end: np.ndarray                  # int64
    chrom_len: dict = field(default_factory=dict)

    @property
    def N(self):
        return self.X.shape[0]

    @property
    def n(self):
        return self.X.shape[1]

    def addresses(self):
        return {nm: (c, int(s), int(e)) for nm, c, s, e in zip(self.names, self.chrom, self.start, self.end)}


def make_table(N, n, seed=0, zero_row_frac=0.005, dup_rows=4, n_programmes=8, dtype=np.float64):
    rng = np.random.default_rng(seed)
    # ---- genomic layout: genes with a few alternative promoters each --------------------
    # chromosome lengths scale with N so that the feature density stays FANTOM5-like
    # (~200k regions on a 3.1 Gb genome) and shifted loci still land on features.
    lens = np.maximum((np.array(_CHROM_MB, dtype=np.float64) * 1e6 * (N / 200000.0)).astype(np.int64),
                      200_000)
    genes = max(1, int(N / 2.2))
    g_chrom = rng.choice(len(lens), size=genes, p=lens / lens.sum())
    margin = np.minimum(100_000, lens // 10)
    g_pos = (rng.random(genes) * (lens[g_chrom] - 2 * margin[g_chrom])).astype(np.int64) + margin[g_chrom]
    per_gene = 1 + rng.poisson(1.2, size=genes)
    gi = np.repeat(np.arange(genes), per_gene)
    while gi.size < N:                       # top up with singletons
        gi = np.concatenate([gi, rng.integers(0, genes, size=N - gi.size)])
    gi = gi[rng.permutation(gi.size)[:N]]
    c = g_chrom[gi]
    s = g_pos[gi] + rng.integers(-5000, 5000, size=N)
    w = rng.integers(8, 40, size=N)
    strand = np.where(rng.random(N) < 0.5, "+", "-")
    # unique (chrom, start): nudge duplicates
    key = c.astype(np.int64) * (1 << 40) + s
    _, first = np.unique(key, return_index=True)
    dup = np.ones(N, dtype=bool); dup[first] = False
    s[dup] += np.arange(1, dup.sum() + 1) * 50 + 60
    order = np.lexsort((s, c))
    c, s, w, strand, gi = c[order], s[order], w[order], strand[order], gi[order]
    e = s + w
    chrom = np.array(["chr" + _CHROM_NAMES[k] for k in c])
    names = [f"{ch}:{a}..{b},{st}" for ch, a, b, st in zip(chrom, s, e, strand)]
    # ---- expression: programmes shared by promoters of one gene --------------------------
    F = rng.lognormal(0.0, 1.0, size=(n_programmes, n))
    load_gene = rng.gamma(0.35, 1.0, size=(genes, n_programmes)) * (rng.random((genes, n_programmes)) < 0.4)
    L = load_gene[gi] * rng.lognormal(0.0, 0.3, size=(N, n_programmes))
    mean = L @ F + 0.05
    X = rng.gamma(0.6, 1.0, size=(N, n)) * mean / 0.6
    X[X < 0.35] = 0.0                         # zero inflation -> long tie runs
    X = np.round(X, 6)
    nz = max(1, int(round(zero_row_frac * N))) if zero_row_frac > 0 else 0
    if nz:
        X[rng.choice(N, size=nz, replace=False)] = 0.0
    for _ in range(dup_rows):
        a, b = rng.choice(N, size=2, replace=False)
        X[b] = X[a]
    return Table(names=names, X=np.ascontiguousarray(X, dtype=dtype), chrom=chrom,
                 start=s.astype(np.int64), end=e.astype(np.int64),
                 chrom_len={"chr" + nm: int(l) for nm, l in zip(_CHROM_NAMES, lens)})


# ------------------------------------------------------------------------------------------
# loci -> promoters (synthetic windowBed -w) and circular permutation of the loci
# ------------------------------------------------------------------------------------------
class Mapper:
    """Overlap join of SNP positions (+-window) with the feature table, in genome-wide
    coordinates (the reference's `gwad`, coexfunctions.py:104-130)."""

    def __init__(self, table: Table):
        self.t = table
        chroms = list(table.chrom_len)
        off = np.concatenate([[0], np.cumsum([table.chrom_len[c] for c in chroms])])
        self.chrom_off = {c: int(o) for c, o in zip(chroms, off[:-1])}
        self.chrom_ids = {c: i for i, c in enumerate(chroms)}
        self.chrom_list = chroms
        self.chrom_starts = off[:-1]
        self.genome_len = int(off[-1])
        self.f_chrom = np.array([self.chrom_ids[c] for c in table.chrom])
        self.f_gs = np.array([self.chrom_off[c] for c in table.chrom], dtype=np.int64) + table.start
        self.f_ge = self.f_gs + (table.end - table.start)
        self.order = np.argsort(self.f_gs, kind="stable")
        self.sorted_gs = self.f_gs[self.order]
        self.maxw = int((table.end - table.start).max())

    def gwad_to_chrom(self, g):
        ci = np.searchsorted(self.chrom_starts, g, side="right") - 1
        return ci, g - self.chrom_starts[ci]

    def map_loci(self, gwad, window):
        """-> list of (snp_index, feature_row) in (snp order, feature table order)."""
        gwad = np.asarray(gwad, dtype=np.int64)
        ci, _ = self.gwad_to_chrom(gwad)
        lo = np.searchsorted(self.sorted_gs, gwad - window - self.maxw, side="left")
        hi = np.searchsorted(self.sorted_gs, gwad + 1 + window, side="left")
        out = []
        for k in range(gwad.size):
            if hi[k] <= lo[k]:
                continue
            rows = self.order[lo[k]:hi[k]]
            ok = (self.f_chrom[rows] == ci[k]) & (self.f_ge[rows] > gwad[k] - window) & \
                 (self.f_gs[rows] < gwad[k] + 1 + window)
            for r in np.sort(rows[ok]):
                out.append((k, int(r)))
        return out


def make_loci(table: Table, n_loci, seed=1, near_feature_frac=0.8):
    """GWAS-like loci: most sit near a feature (so they map), some are intergenic."""
    rng = np.random.default_rng(seed)
    m = Mapper(table)
    near = rng.random(n_loci) < near_feature_frac
    g = np.empty(n_loci, dtype=np.int64)
    pick = rng.integers(0, table.N, size=n_loci)
    g[near] = m.f_gs[pick[near]] + rng.integers(-3000, 3000, size=int(near.sum()))
    g[~near] = rng.integers(0, m.genome_len, size=int((~near).sum()))
    return np.clip(g, 0, m.genome_len - 2)


def circular_shifts(n_loci_gwad, genome_len, numperms, seed=2):
    """k = 0 is the real data; k >= 1 adds one random genome-wide offset to every locus and
    wraps (the 'circular' mode of 0-prepare-coex.py:471-613)."""
    rng = np.random.default_rng(seed)
    sets = [np.asarray(n_loci_gwad, dtype=np.int64)]
    for _ in range(numperms):
        shift = int(rng.integers(1, genome_len))
        sets.append((sets[0] + shift) % genome_len)
    return sets


def hit_rows_from_pairs(pairs):
    """1-make-network.py:131-177: promoters = first-appearance order of the last column."""
    seen, rows = set(), []
    for _, r in pairs:
        if r not in seen:
            seen.add(r)
            rows.append(r)
    return np.array(rows, dtype=np.int32)


def make_permutation_hits(table: Table, n_loci, numperms, window=2000, seed=3):
    """-> list over k = 0..numperms of int32 hit-row arrays (map-file order)."""
    m = Mapper(table)
    loci = make_loci(table, n_loci, seed)
    return [hit_rows_from_pairs(m.map_loci(g, window))
            for g in circular_shifts(loci, m.genome_len, numperms, seed + 1)], loci, m


def write_map_file(path, table: Table, mapper: Mapper, gwad, window, prefix="rs"):
    """windowBed -a snps -b features output: A cols (chrom start end snp) + B cols; the
    consumer uses col 3 and the last col only (1-make-network.py:137-147)."""
    ci, pos = mapper.gwad_to_chrom(np.asarray(gwad, dtype=np.int64))
    with open(path, "w") as o:
        for k, r in mapper.map_loci(gwad, window):
            ch = mapper.chrom_list[ci[k]].replace("chr", "")
            o.write(f"{ch}\t{pos[k]}\t{pos[k] + 1}\t{prefix}{k}\t{table.chrom[r]}\t{table.start[r]}"
                    f"\t{table.end[r]}\t{table.names[r]}\n")


def write_expression_file(path, table: Table):
    """`sample <tab> s0 <tab> s1 ...` header, one row per region (coexfunctions.py:204-209)."""
    with open(path, "w") as o:
        o.write("sample\t" + "\t".join(f"s{j}" for j in range(table.n)) + "\n")
        for nm, row in zip(table.names, table.X):
            o.write(nm + "\t" + "\t".join(repr(float(v)) for v in row) + "\n")


def write_feature_bed(path, table: Table):
    with open(path, "w") as o:
        for nm, c, s, e in zip(table.names, table.chrom, table.start, table.end):
            o.write(f"{c}\t{s}\t{e}\t{nm}\n")


def global_pair_indices(names, npairs, seed=4):
    """0-prepare-coex.py:672-687: random pairs, kept only if the first five characters of
    the two labels differ ("different chromosome")."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, len(names), size=npairs)
    b = rng.integers(0, len(names), size=npairs)
    pre = np.array([nm[:5] for nm in names])
    keep = pre[a] != pre[b]
    return a[keep].astype(np.int32), b[keep].astype(np.int32)
