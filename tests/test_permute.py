"""Row N1: permutation generation + locus -> promoter mapping.
CPU: the restated reference arithmetic against hand-computed answers.  GPU: coex_permute_map against the restatement."""
import numpy as np
import pytest

from oracle import permute_oracle as po


CHROMLIST = ["chr1", "chr2", "chr3"]
LENGTHS = {"chr1": 1000, "chr2": 500, "chr3": 250}


def test_gwad_and_address_known_answers():
    cr = po.chromrange_from_lengths(CHROMLIST, LENGTHS)
    assert cr == {"chr1": [0, 1000], "chr2": [1000, 1500], "chr3": [1500, 1750]}
    assert po.get_gwad("chr2", 7, cr) == 1007
    assert po.get_gwad("chr2", 500, cr) is None            # not < end: the reference returns None (coexfunctions.py:109-112)
    assert po.get_gwad("chr2", -3, cr) == 997              # negative addresses pass
    assert po.get_gwad("chrQ", 1, cr) is None
    assert po.get_address(1007, cr, CHROMLIST) == ("chr2", 7)
    assert po.get_address(1000, cr, CHROMLIST) == ("chr1", 1000)   # start < gwad <= end: a boundary belongs to the LEFT chromosome
    assert po.get_address(1750 + 5, cr, CHROMLIST) == ("chr1", 5)  # wrap: while gwad > genome_max
    assert po.get_address(2 * 1750, cr, CHROMLIST) == ("chr3", 250)
    assert po.get_address(0, cr, CHROMLIST) == ("chr1", 0)         # nothing matches: the reference's fallback


def test_listitem_rounds_halves_away_from_zero():
    assert po.round_half_away(0.5) == 1.0 and po.round_half_away(-0.5) == -1.0 and po.round_half_away(2.5) == 3.0
    assert po.round_half_away(0.49999999999999994) == 0.0
    n = 10
    assert [po.listitem_index(n, i) for i in range(0, 12)] == [0, 1, 2, 3, 4, -5, -4, -3, -2, -1, 0, 1]
    for idx in (3, 17, 123456789, 10 ** 10 + 7):
        assert po.listitem_index(n, idx) % n == idx % n


def test_circular_and_window_known_answers():
    cr = po.chromrange_from_lengths(CHROMLIST, LENGTHS)
    snps = [("chr1", 990), ("chr2", 10), ("chr3", 249)]
    assert po.circular_positions(snps, 0, cr, CHROMLIST) == [("chr1", 990), ("chr2", 10), ("chr3", 249)]
    assert po.circular_positions(snps, 20, cr, CHROMLIST) == [("chr2", 10), ("chr2", 30), ("chr1", 19)]
    fc, fs, fe = ["chr2", "chr2", "chr1", "chr2"], [0, 40, 15, 26], [8, 60, 18, 27]
    pairs = po.window_bed([("chr2", 10), ("chr2", 30), ("chr1", 19), None], fc, fs, fe, 2)
    # chr2:10 +-2 = [8, 13): feature 0 ends at 8 (not > 8) -> no hit; chr2:30 = [28, 33): none; chr1:19 = [17, 22): feature 2
    assert pairs == [(2, 2)]
    pairs = po.window_bed([("chr2", 10), ("chr2", 30), ("chr1", 19)], fc, fs, fe, 10)
    assert pairs == [(0, 0), (1, 1), (1, 3), (2, 2)]
    assert po.hit_rows(pairs + [(3, 0)]) == [0, 1, 3, 2]


def _synthetic_inputs(seed, N=4000, n=8, M=120):
    from coexpression_b200 import synth
    t = synth.make_table(N, n, seed=seed)
    chromlist = list(t.chrom_len)
    cr = po.chromrange_from_lengths(chromlist, t.chrom_len)
    rng = np.random.default_rng(seed + 1)
    pick = rng.integers(0, N, size=M)
    snps = []
    for r in pick:
        c = t.chrom[r]
        a = int(np.clip(t.start[r] + rng.integers(-4000, 4000), 1, t.chrom_len[c] - 2))
        snps.append((c, a))
    snps = list(dict.fromkeys(snps))
    return t, chromlist, cr, snps, rng


@pytest.mark.gpu
@pytest.mark.parametrize("window", [0, 2000, 50000])
def test_gpu_circular_matches_restatement(window):
    from coexpression_b200.api import CoexContext
    t, chromlist, cr, snps, rng = _synthetic_inputs(5)
    gmax = cr[chromlist[-1]][1]
    shifts = [0] + [int(rng.random() * 10000000000) for _ in range(6)] + [gmax, 1, gmax - 1]
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        got = ctx.permute_map("circular", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, window)
    for k, s in enumerate(shifts):
        want = po.hit_rows(po.window_bed(po.circular_positions(snps, s, cr, chromlist), t.chrom, t.start, t.end, window))
        assert got[k].tolist() == want, (k, s)
    assert len(got[0]) > 0 or window == 0


@pytest.mark.gpu
def test_gpu_backcirc_matches_restatement():
    from coexpression_b200.api import CoexContext
    t, chromlist, cr, snps, rng = _synthetic_inputs(9)
    # background = random positions plus the real loci (0-prepare-coex.py:520-528), sorted by (gwad, key)
    backdict = {}
    for _ in range(3000):
        c = chromlist[int(rng.integers(0, len(chromlist)))]
        a = int(rng.integers(1, t.chrom_len[c] - 1))
        backdict["%s_%s" % (c, a)] = po.get_gwad(c, a, cr)
    for c, a in snps:
        backdict["%s_%s" % (c, a)] = po.get_gwad(c, a, cr)
    sb = po.sorted_background(backdict)
    background = [(k.split("_")[0], int(k.split("_")[1])) for k in sb]
    shifts = [0] + [int(rng.random() * 10000000000) for _ in range(8)] + [len(sb) // 2, len(sb) // 2 + 1, len(sb)]
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        got = ctx.permute_map("backcirc", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, 3000,
                              background=background)
    for k, s in enumerate(shifts):
        pos = po.circular_positions(snps, 0, cr, chromlist) if s == 0 else po.backcirc_positions(snps, s, sb)
        want = po.hit_rows(po.window_bed(pos, t.chrom, t.start, t.end, 3000))
        assert got[k].tolist() == want, (k, s)


def test_background_positions_known_answers():
    # 0-prepare-coex.py:568-572: chrom = column 0 without chr, address = column 2 (the END column), one locus per drawn line
    backlines = [["chr1", "99", "100", "x", "rsA"], ["2", "500", "501", "x", "rsB"], ["chrX", "7", "8", "x", "rsC"]]
    assert po.background_positions(backlines, [2, 0]) == [("chrX", 8), ("chr1", 100)]
    assert po.background_positions(backlines, [1]) == [("chr2", 501)]


@pytest.mark.gpu
def test_gpu_background_matches_restatement():
    import random
    from coexpression_b200.api import CoexContext
    t, chromlist, cr, snps, rng = _synthetic_inputs(11)
    # a background file: 4000 lines `chrom start end id snp`, some of them next to features, duplicates allowed
    backlines = []
    for _ in range(4000):
        c = chromlist[int(rng.integers(0, len(chromlist)))]
        a = int(rng.integers(1, t.chrom_len[c] - 1))
        backlines.append([c.replace("chr", ""), str(a - 1), str(a), "bg", "rs%d" % a])
    backlines += backlines[:50]
    pyrand = random.Random(17)
    K = 9
    shifts = [0] + list(range(1, K))                       # 0-prepare-coex.py:486-489: shift = i + 1, 0 = real data
    picks = [pyrand.sample(range(len(backlines)), len(snps)) for _ in range(K)]
    background = [("chr" + l[0], int(l[2])) for l in backlines]
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        got = ctx.permute_map("background", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, 20000,
                              background=background, picks=picks)
    n_hits = 0
    for k, s in enumerate(shifts):
        pos = po.circular_positions(snps, 0, cr, chromlist) if s == 0 else po.background_positions(backlines, picks[k])
        want = po.hit_rows(po.window_bed(pos, t.chrom, t.start, t.end, 20000))
        assert got[k].tolist() == want, (k, s)
        n_hits += len(want)
    assert n_hits > K


@pytest.mark.gpu
def test_gpu_mapped_sets_feed_the_network_batch():
    """End to end on the device path: permute_map -> network_batch gives what host-made hit lists give."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    t, chromlist, cr, snps, rng = _synthetic_inputs(21, N=3000, n=64, M=150)
    shifts = [0] + [int(rng.random() * 10000000000) for _ in range(5)]
    ga, gb = synth.global_pair_indices(t.names, 20000, seed=3)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        g = ctx.pairs(ga, gb, ranked=True)
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        hits = ctx.permute_map("circular", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, 5000)
        want = [np.array(po.hit_rows(po.window_bed(po.circular_positions(snps, s, cr, chromlist), t.chrom, t.start,
                                                   t.end, 5000)), dtype=np.int32) for s in shifts]
        a = ctx.network_batch(hits, Options())
        b = ctx.network_batch(want, Options())
    assert np.array_equal(a.group_density, b.group_density) and np.array_equal(a.n_groups, b.n_groups)
    assert sum(len(h) for h in hits) > 50


def test_window_bed_reports_in_bedtools_bin_order():
    """`bedtools window` walks the UCSC bins from the finest level up: a feature straddling a 16 kb boundary sits in a
    coarser bin and is reported AFTER the finer-bin features behind it, and features of one bin come in B-file order."""
    assert po.bedtools_bin(0, 10) == (0, 0) and po.bedtools_bin(16383, 16384) == (0, 0)
    assert po.bedtools_bin(16380, 16390) == (1, 0)            # straddles the first 16 kb boundary -> 128 kb level
    assert po.bedtools_bin(131070, 131080) == (2, 0)          # straddles a 128 kb boundary -> 1 Mb level
    assert po.bedtools_bin(16384, 16400) == (0, 1)
    fc = ["chr1"] * 5
    fs = [16000, 16380, 16390, 16500, 40000]
    fe = [16010, 16390, 16400, 16510, 40010]
    pairs = po.window_bed([("chr1", 16395)], fc, fs, fe, 1000)
    # level 0: bin 0 -> row 0, bin 1 -> rows 2, 3; level 1: row 1 (straddler)
    assert pairs == [(0, 0), (0, 2), (0, 3), (0, 1)]
    # B-file order inside a bin: the BED lists row 3 before row 2
    pairs = po.window_bed([("chr1", 16395)], fc, fs, fe, 1000, bed_index=[0, 1, 3, 2, 4])
    assert pairs == [(0, 0), (0, 3), (0, 2), (0, 1)]
    # a wide window reaches the next bin of level 0 before any coarser level
    pairs = po.window_bed([("chr1", 16395)], fc, fs, fe, 30000)
    assert pairs == [(0, 0), (0, 2), (0, 3), (0, 4), (0, 1)]


@pytest.mark.gpu
def test_gpu_window_order_follows_the_bed_file():
    """coex_set_feature_order: the B file lists the features in another order than the table; promoters of a locus come
    out by (UCSC-bin level, bin, B-file line) as `bedtools window` reports them.  Long features (some straddling 16 kb
    bin boundaries) make the coarser levels matter."""
    from coexpression_b200.api import CoexContext
    t, chromlist, cr, snps, rng = _synthetic_inputs(21, N=6000)
    start = t.start.copy()
    end = t.end.copy()
    wide = rng.random(t.N) < 0.3
    end[wide] = start[wide] + rng.integers(2000, 40000, size=int(wide.sum()))
    bed_index = rng.permutation(t.N).astype(np.int32)
    levels = {po.bedtools_bin(int(s), int(e))[0] for s, e in zip(start, end)}
    assert len(levels) >= 3
    shifts = [0] + [int(rng.random() * 10000000000) for _ in range(5)]
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, start, end, t.names)
        got_default = ctx.permute_map("circular", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, 20000)
        ctx.set_feature_order(bed_index)
        got = ctx.permute_map("circular", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, 20000)
    differs = False
    for k, s in enumerate(shifts):
        pos = po.circular_positions(snps, s, cr, chromlist)
        assert got_default[k].tolist() == po.hit_rows(po.window_bed(pos, t.chrom, start, end, 20000)), k
        assert got[k].tolist() == po.hit_rows(po.window_bed(pos, t.chrom, start, end, 20000, bed_index=bed_index)), k
        differs |= got[k].tolist() != got_default[k].tolist()
        assert sorted(got[k].tolist()) == sorted(got_default[k].tolist())
    assert differs
