"""Parity of the CUDA path (through the C ABI) with the oracle.  Needs a B200: -m gpu."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def _same_doubles(a, b):
    """bit-exact where defined, NaN exactly where the reference has NaN"""
    a = np.asarray(a); b = np.asarray(b)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(_bits(a[~np.isnan(a)]), _bits(b[~np.isnan(b)]))


@pytest.fixture(scope="module")
def lib():
    from coexpression_b200 import _lib
    return _lib.load()


# ---------------------------------------------------------------------------------------------
# A1: the legacy drop-in symbol, called exactly as 1-make-network.py:229-234 does
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["getpearson_raw", "getpearson_ranked", "getpearson_onepair"])
def test_legacy_getpearson_golden(lib, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    X = np.ascontiguousarray(g["X"]); ia = np.ascontiguousarray(g["ia"]); ib = np.ascontiguousarray(g["ib"])
    n_pairs = ia.size
    data_c = (ctypes.c_double * X.size)(*X.ravel())
    idxa_c = (ctypes.c_int * n_pairs)(*ia.tolist())
    idxb_c = (ctypes.c_int * n_pairs)(*ib.tolist())
    result_c = (ctypes.c_double * n_pairs)()
    raw = ctypes.cdll.LoadLibrary(lib._name)          # untyped, as the reference calls it
    raw.getPearson(ctypes.byref(data_c), ctypes.byref(result_c), ctypes.byref(idxa_c), ctypes.byref(idxb_c),
                   ctypes.c_int(n_pairs), ctypes.c_int(X.shape[1]))
    assert _same_doubles(np.array(result_c[:]), g["r"])


def test_legacy_getpearson_zero_pairs_and_random(lib):
    from oracle import network_oracle as no
    rng = np.random.default_rng(5)
    X = rng.lognormal(0, 2, size=(700, 1829)) * (rng.random((700, 1829)) < 0.5)
    X[11] = 0.0
    ia = rng.integers(0, 700, 30000).astype(np.int32); ib = rng.integers(0, 700, 30000).astype(np.int32)
    out = np.full(ia.size, 7.0)
    lib.getPearson(X.ctypes.data, out.ctypes.data, ia.ctypes.data, ib.ctypes.data, 0, 1829)
    assert np.all(out == 7.0)                           # nPromPairs == 0 touches nothing
    lib.getPearson(X.ctypes.data, out.ctypes.data, ia.ctypes.data, ib.ctypes.data, ia.size, 1829)
    assert _same_doubles(out, no.get_pearson(X, ia, ib))
    Y = X - 0.3                                         # mixed signs
    lib.getPearson(Y.ctypes.data, out.ctypes.data, ia.ctypes.data, ib.ctypes.data, ia.size, 1829)
    assert _same_doubles(out, no.get_pearson(Y, ia, ib))


# ---------------------------------------------------------------------------------------------
# A2: ranks bit-exact with scipy.stats.rankdata
# ---------------------------------------------------------------------------------------------
def test_ranks_golden_and_random():
    from coexpression_b200.api import CoexContext
    from oracle import network_oracle as no
    g = np.load(os.path.join(GOLDEN, "ranks_small.npz"))
    with CoexContext() as ctx:
        ctx.set_expression(g["X"]); ctx.prepare()
        assert np.array_equal(ctx.ranks(), g["ranks"])
        rng = np.random.default_rng(3)
        for n in (1, 2, 31, 256, 1000, 1829):
            X = np.round(rng.gamma(0.4, 3.0, size=(97, n)), 1)      # heavy ties
            X[0] = 0.0
            X[1] = np.arange(n)
            ctx.set_expression(X); ctx.prepare()
            assert np.array_equal(ctx.ranks(), no.rank_rows(X)), n


def test_pairs_on_resident_table():
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    from oracle import network_oracle as no
    t = synth.make_table(900, 333, seed=9)
    R = no.rank_rows(t.X)
    rng = np.random.default_rng(1)
    ia = rng.integers(0, 900, 20000).astype(np.int32); ib = rng.integers(0, 900, 20000).astype(np.int32)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        assert _same_doubles(ctx.pairs(ia, ib, ranked=True), no.get_pearson(R, ia, ib))
        assert _same_doubles(ctx.pairs(ia, ib, ranked=False), no.get_pearson(t.X, ia, ib))


# ---------------------------------------------------------------------------------------------
# K3: tcgen05 GEMM == dp4a GEMM (exact integers), several K extents incl. the FANTOM5 one
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,n", [(600, 48), (1000, 300), (2500, 1829), (700, 1860)])
def test_tcgen05_gemm_matches_dp4a(N, n):
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    t = synth.make_table(N, n, seed=N + n)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        for mode in ("tcgen05_bn128", "tcgen05_bn64", "tcgen05_persistent", "tcgen05_persistent128", "tcgen05_cluster128", "tcgen05_split"):
            ctx.set_gemm_mode(mode)
            bad, first = ctx.selftest_gemm(min(N, 512))
            assert bad == 0, (mode, first)


# ---------------------------------------------------------------------------------------------
# A4-A12: golden network cases (oracle consistent mode driving the reference C build)
# ---------------------------------------------------------------------------------------------
NETWORK_CASES = ["spearman_default", "spearman_anticor", "spearman_noiter_j001", "spearman_useavg_m",
                 "spearman_global_nsp0", "pearson_default", "pearson_anticor_j05"]


def _run_case(name, gemm_mode, count_mode="dedup"):
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    with open(os.path.join(GOLDEN, f"network_{name}.json")) as f:
        case = json.load(f)
    kw = case["kwargs"]
    cm = kw.get("correlationmeasure", "Spearman")
    inp = np.load(os.path.join(GOLDEN, f"network_inputs_{case['shape']}_{cm}.npz"))
    names = [str(s) for s in inp["names"]]
    opts = Options(**kw)
    with CoexContext() as ctx:
        ctx.set_gemm_mode(gemm_mode)
        ctx.set_count_mode(count_mode)
        ctx.set_expression(inp["X"]); ctx.prepare()
        ctx.set_features(inp["chrom"], inp["start"], inp["end"], names)
        ctx.set_global_list(inp["glist"])
        res = ctx.network_batch([np.array(p["hit_rows"], dtype=np.int32) for p in case["perms"]], opts,
                                want_scores=True, want_counts=True)
    for k, perm in enumerate(case["perms"]):
        got = perm_objects(res, k, names)
        P = len(perm["hit_rows"])
        assert np.array_equal(res.count_matrix(k), np.array(perm["counts"]).reshape(P, P)), (name, k)
        assert got["prom_to_join"] == perm["prom_to_join"], (name, k)            # bit-exact groups
        assert got["winners"] == perm["winners"], (name, k)
        assert set(got["indmeasure"]) == set(perm["indmeasure"])
        for g, v in perm["indmeasure"].items():
            assert abs(got["indmeasure"][g] - v) <= 1e-6, (name, k, g)           # north_star tolerance
        for p_, v in perm["allpromoterscores"].items():
            assert abs(got["allpromoterscores"][p_] - v) <= 1e-6


@pytest.mark.parametrize("name", NETWORK_CASES)
def test_network_golden_tcgen05(name):
    _run_case(name, "tcgen05")


@pytest.mark.parametrize("name", ["spearman_default", "spearman_anticor"])
def test_network_golden_dp4a(name):
    _run_case(name, "simt")


@pytest.mark.parametrize("name", ["spearman_default", "spearman_anticor", "spearman_useavg_m"])
@pytest.mark.parametrize("count_mode", ["per_query", "dedup_sorted", "dedup_fine", "row_rank"])
def test_network_golden_other_count_kernels(name, count_mode):
    _run_case(name, "tcgen05", count_mode)


def test_network_against_live_oracle_small_shape():
    """Bigger than the golden cases: 3000 x 96 table, 4 hit sets incl. an empty and a single-hit one."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    N, n = synth.SHAPES["small"]
    t = synth.make_table(N, n, seed=31)
    hits, loci, m = synth.make_permutation_hits(t, 150, 3, window=6000, seed=8)
    hits = hits + [np.zeros(0, dtype=np.int32), hits[0][:1]]
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 50000, seed=2)
    g = no.get_pearson(R, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    for opts in (Options(), Options(anticorrelation=True, pval_to_join=0.01)):
        with CoexContext() as ctx:
            ctx.set_expression(t.X); ctx.prepare()
            ctx.set_features(t.chrom, t.start, t.end, t.names)
            ctx.set_global_list(glist)
            # the global list itself through the resident-table pair kernel (call site #2)
            g_gpu = ctx.pairs(ga, gb, ranked=True)
            assert _same_doubles(g_gpu, g)
            res = ctx.network_batch(hits, opts, want_scores=True, want_counts=True)
        for k, h in enumerate(hits):
            want = no.make_network(t.names, t.X, [t.names[r] for r in h], t.addresses(), glist.tolist(),
                                   anticorrelation=opts.anticorrelation, pval_to_join=opts.pval_to_join,
                                   Xrank=R, return_counts=True)
            got = perm_objects(res, k, t.names)
            assert got["prom_to_join"] == want["prom_to_join"]
            assert got["winners"] == want["winners"]
            if len(h) > 1:
                assert np.array_equal(res.count_matrix(k), want["_counts"])
            assert set(got["indmeasure"]) == set(want["indmeasure"])
            for grp, v in want["indmeasure"].items():
                assert abs(got["indmeasure"][grp] - v) <= 1e-6
            for a in want["individualscores"]:
                for b, v in want["individualscores"][a].items():
                    assert abs(got["individualscores"][a][b] - v) <= 1e-6


@pytest.mark.parametrize("N,n", [(700, 64), (1500, 300), (900, 1829)])
def test_corr_block_spearman_exact_and_pearson_tensor(N, n):
    """Spearman block: bit-identical to the reference.  Pearson on the tensor cores (z-scores in
    int8 digits, exact integer accumulation): within north_star's 1e-6 absolute of the reference's
    doubles.  In-order double Pearson: bit-identical."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    from oracle import network_oracle as no
    t = synth.make_table(N, n, seed=N + 3 * n)
    X = t.X.copy()
    X[130] = 2.5                      # constant row -> NaN
    X[131] = X[140] * 1000.0 + 1.0    # heavy-tailed scale
    R = no.rank_rows(X)
    rows = np.arange(128, 128 + 160)
    ia = np.repeat(rows, N).astype(np.int32); ib = np.tile(np.arange(N), rows.size).astype(np.int32)
    want_s = no.get_pearson(R, ia, ib).reshape(rows.size, N)
    want_p = no.get_pearson(X, ia, ib).reshape(rows.size, N)
    with CoexContext() as ctx:
        ctx.set_expression(X); ctx.prepare()
        assert _same_doubles(ctx.corr_block("spearman", 128, rows.size), want_s)
        assert _same_doubles(ctx.corr_block("pearson_exact", 128, rows.size), want_p)
        errs = {}
        for digits in (3, 4):
            ctx.set_pearson(digits, "tensor")
            got = ctx.corr_block("pearson", 128, rows.size)
            assert np.array_equal(np.isnan(got), np.isnan(want_p))
            errs[digits] = np.nanmax(np.abs(got - want_p))
            print(f"pearson tensor digits={digits} N={N} n={n}: max |r - reference| = {errs[digits]:.3g}")
        # north_star tolerance for correlations: 1e-6 absolute -> met by the default 4 digits with two
        # orders of margin; 3 digits (the faster option) stay within 5e-6
        assert errs[4] <= 2e-8 and errs[3] <= 5e-6, errs


def test_allpairs_histogram_small():
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    from oracle import network_oracle as no
    t = synth.make_table(500, 64, seed=12)
    R = no.rank_rows(t.X)
    iu = np.triu_indices(500, 1)
    nb = 4096
    for ranked, data in ((True, R), (False, t.X)):
        r = no.get_pearson(data, iu[0].astype(np.int32), iu[1].astype(np.int32))
        ok = ~np.isnan(r)
        want = np.bincount(np.clip(((r[ok] + 1.0) * 0.5 * nb).astype(np.int64), 0, nb - 1), minlength=nb)
        with CoexContext() as ctx:
            ctx.set_expression(t.X); ctx.prepare()
            if not ranked:
                ctx.set_pearson(3, "exact")
            hist, nan = ctx.allpairs_histogram(ranked, nbins=nb)
            assert nan == int((~ok).sum())
            assert np.array_equal(hist.astype(np.int64), want)
            if not ranked:                      # tensor-core Pearson: values within 1e-6 -> bins may move at edges only
                ctx.set_pearson(4, "tensor")
                hist_t, nan_t = ctx.allpairs_histogram(False, nbins=nb)
                assert nan_t == nan and int(hist_t.sum()) == int(want.sum())
                near_edge = np.abs(((r[ok] + 1.0) * 0.5 * nb) - np.round((r[ok] + 1.0) * 0.5 * nb)) < 1e-6 * nb
                assert np.abs(hist_t.astype(np.int64) - want).sum() <= 2 * int(near_edge.sum())


def test_allpairs_accumulate_equals_one_shot():
    """Device-resident accumulation over row blocks (what the multi-GPU C3 driver does) == the one-shot histogram,
    for the exact Spearman path (upper-triangle GEMM strips) and the Pearson tensor path."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    t = synth.make_table(2300, 200, seed=31)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        for ranked in (True, False):
            want, want_nan = ctx.allpairs_histogram(ranked, 0, t.N, 4096)
            blocks = [(0, 512), (512, 1408), (1408, t.N)]
            for i, (r0, r1) in enumerate(blocks):
                ctx.allpairs_accumulate(ranked, r0, r1, 4096, reset=(i == 0))
            got, got_nan = ctx.allpairs_read(4096)
            assert np.array_equal(got, want) and got_nan == want_nan
            assert int(got.sum()) + got_nan == t.N * (t.N - 1) // 2


def test_file_level_entry_point_matches_oracle(tmp_path):
    """The drop-in for `1-make-network.py -wd ... --all`: same inputs on disk (settings.txt, expression table,
    feature BED, correlation list, windowBed map files), same eight JSON objects per permutation."""
    from coexpression_b200 import synth
    from coexpression_b200.make_network import run_batch
    from oracle import network_oracle as no
    t = synth.make_table(1200, 60, seed=44)
    wd = tmp_path / "work"
    (wd / "permutation_store").mkdir(parents=True)
    expr = tmp_path / "t.expression"; bed = tmp_path / "f5ep.bed"; clist = tmp_path / "t.expression.spearmanlist"
    synth.write_expression_file(expr, t); synth.write_feature_bed(bed, t)
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 30000, seed=3)
    g = no.get_pearson(R, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    with open(clist, "w") as o:
        for v in glist:
            o.write("%s\n" % (int(v) if v == -99 else repr(float(v))))
    m = synth.Mapper(t)
    loci = synth.make_loci(t, 70, seed=9)
    sets = synth.circular_shifts(loci, m.genome_len, 2, seed=10)
    for k, gw in enumerate(sets):
        synth.write_map_file(wd / "permutation_store" / f"f5ep.{k}.bed", t, m, gw, 7000, prefix="rs" if k == 0 else "rand_rs")
    settings = {"verbose": "False", "feature_coordinates": str(bed), "correlationfile": str(clist), "expression_file": str(expr),
                "supplementary_label": "x", "correlationmeasure": "Spearman", "soft_separation": 100000, "nsp": 1.0,
                "seedfile": "none", "pval_to_join": 0.1, "useaverage": "False", "anticorrelation": "True", "noiter": "False",
                "selectionthreshold": 0, "selectionthreshold_is_j": "False", "config_file": "app.cfg"}
    json.dump(settings, open(wd / "settings.txt", "w"))
    written = run_batch(str(wd))
    assert len(written) == 3 * 8
    store = wd / "perm_results_store"
    for k in range(3):
        snp_map = json.load(open(store / f"f5ep.snp_map.{k}"))
        promoters = list(snp_map)
        want = no.make_network(t.names, t.X, promoters, t.addresses(), glist.tolist(), anticorrelation=True, Xrank=R)
        got = {name: json.load(open(store / f"f5ep.{name}.{k}")) for name in
               ("indmeasure", "individualscores", "prom_to_join", "indjoined", "joining_instructions", "winners",
                "allpromoterscores")}
        assert got["prom_to_join"] == want["prom_to_join"] and got["winners"] == want["winners"]
        assert got["joining_instructions"] == want["joining_instructions"]
        assert set(got["indmeasure"]) == set(want["indmeasure"]) and len(promoters) > 5
        for grp, v in want["indmeasure"].items():
            assert abs(got["indmeasure"][grp] - v) <= 1e-6
        for a_ in want["indjoined"]:
            for b_, v in want["indjoined"][a_].items():
                assert abs(got["indjoined"][a_][b_] - v) <= 1e-6
        for a_ in want["individualscores"]:
            for b_, v in want["individualscores"][a_].items():
                assert abs(got["individualscores"][a_][b_] - v) <= 1e-6


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_randomised_stress_against_oracle(seed):
    """Many small random configurations in one batch: duplicated rows inside and across groups, all-zero and
    constant rows among the hits, two-hit and one-hit sets, dense clusters (every hit within the soft
    separation), options drawn at random.  Everything the reference defines must match the oracle."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    rng = np.random.default_rng(100 + seed)
    N, n = int(rng.integers(150, 500)), int(rng.integers(5, 70))
    t = synth.make_table(N, n, seed=200 + seed, zero_row_frac=0.03, dup_rows=12)
    X = t.X.copy()
    X[7] = 4.0                                        # constant, non-zero: zero variance -> NaN -> -99
    X[8] = np.round(X[8], 0)                          # heavy ties
    start, end, chrom = t.start.copy(), t.end.copy(), t.chrom.copy()
    dense = rng.choice(N, size=40, replace=False)     # one dense cluster: 40 regions within 60 kb on one chromosome
    chrom[dense] = "chr3"; start[dense] = 5_000_000 + np.sort(rng.integers(0, 60_000, size=40)); end[dense] = start[dense] + 20
    names = [f"{c}:{s}..{e},+" for c, s, e in zip(chrom, start, end)]
    addresses = {nm: (c, int(s), int(e)) for nm, c, s, e in zip(names, chrom, start, end)}
    R = no.rank_rows(X)
    ga = rng.integers(0, N, 8000).astype(np.int32); gb = rng.integers(0, N, 8000).astype(np.int32)
    g = no.get_pearson(R, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    hit_sets = []
    for k in range(7):
        p = int(rng.integers(2, 60))
        h = rng.choice(N, size=p, replace=False)
        if k == 0:
            h = np.concatenate([dense[:25], [7, 8], h[:10]])
        if k == 1:
            h = h[:2]
        if k == 2:
            h = h[:1]
        hit_sets.append(np.unique(h)[rng.permutation(len(np.unique(h)))].astype(np.int32))
    opts = Options(anticorrelation=bool(rng.integers(0, 2)), noiter=bool(rng.integers(0, 2)),
                   useaverage=bool(rng.integers(0, 2)), pval_to_join=float(rng.choice([0.01, 0.1, 0.5, 1.5])),
                   selectionthreshold_is_j=bool(rng.integers(0, 2)))
    with CoexContext() as ctx:
        ctx.set_expression(X); ctx.prepare()
        ctx.set_features(chrom, start, end, names)
        ctx.set_global_list(glist)
        res = ctx.network_batch(hit_sets, opts, want_scores=True, want_counts=True)
    for k, h in enumerate(hit_sets):
        want = no.make_network(names, X, [names[r] for r in h], addresses, glist.tolist(), Xrank=R, return_counts=True,
                               anticorrelation=opts.anticorrelation, noiter=opts.noiter, useaverage=opts.useaverage,
                               pval_to_join=opts.pval_to_join, selectionthreshold_is_j=opts.selectionthreshold_is_j)
        got = perm_objects(res, k, names)
        assert got["prom_to_join"] == want["prom_to_join"], (seed, k)
        assert got["winners"] == want["winners"], (seed, k)
        if len(h) > 1:
            assert np.array_equal(res.count_matrix(k), want["_counts"]), (seed, k)
        assert set(got["indmeasure"]) == set(want["indmeasure"])
        for grp, v in want["indmeasure"].items():
            assert abs(got["indmeasure"][grp] - v) <= 1e-6, (seed, k, grp)
        for p_, v in want["allpromoterscores"].items():
            assert abs(got["allpromoterscores"][p_] - v) <= 1e-6


def test_c1_anchor_minimal_shape_three_permutations():
    """BASELINE config C1 (`-n 3`, Spearman, circular, minimal shape 20000 x 256): the parity anchor.
    Four map files (real data + 3 permutations) against the oracle driving the reference C build."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    N, n = synth.SHAPES["minimal"]
    t = synth.make_table(N, n, seed=7)
    hits, _, _ = synth.make_permutation_hits(t, 60, 3, window=6000, seed=70)
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 100000, seed=71)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        assert np.array_equal(ctx.ranks(), R)
        g = ctx.pairs(ga, gb, ranked=True)
        assert _same_doubles(g, no.get_pearson(R, ga, gb))
        glist = np.sort(np.where(np.isnan(g), -99.0, g))
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        ctx.set_global_list(glist)
        res = ctx.network_batch(hits, Options(), want_scores=True, want_counts=True)
    addresses = t.addresses()
    for k, h in enumerate(hits):
        want = no.make_network(t.names, t.X, [t.names[r] for r in h], addresses, glist.tolist(), Xrank=R, return_counts=True)
        got = perm_objects(res, k, t.names)
        assert got["prom_to_join"] == want["prom_to_join"] and got["winners"] == want["winners"]
        if len(h) > 1:
            assert np.array_equal(res.count_matrix(k), want["_counts"])
        for grp, v in want["indmeasure"].items():
            assert abs(got["indmeasure"][grp] - v) <= 1e-6


def test_c5_like_many_hits_j001_iterative_removal():
    """BASELINE config C5 flavour: one hit set of ~2500 regions (wide window -> large clusters, statistics of a row
    split over several work items), -j 0.01, iterative removal on, plus the same set under -a."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    t = synth.make_table(6000, 80, seed=55)
    hits, _, _ = synth.make_permutation_hits(t, 900, 1, window=40000, seed=56)
    hits = [hits[0][:2500], hits[1][:800]]
    assert len(hits[0]) > 2000
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 60000, seed=57)
    g = no.get_pearson(R, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    addresses = t.addresses()
    for anticor in (False, True):
        opts = Options(pval_to_join=0.01, anticorrelation=anticor)
        with CoexContext() as ctx:
            ctx.set_expression(t.X); ctx.prepare()
            ctx.set_features(t.chrom, t.start, t.end, t.names)
            ctx.set_global_list(glist)
            res = ctx.network_batch(hits, opts, want_scores=False, want_counts=True)
        for k, h in enumerate(hits):
            want = no.make_network(t.names, t.X, [t.names[r] for r in h], addresses, glist.tolist(), Xrank=R,
                                   return_counts=True, pval_to_join=0.01, anticorrelation=anticor)
            got = perm_objects(res, k, t.names, with_scores=False)
            assert got["prom_to_join"] == want["prom_to_join"] and got["winners"] == want["winners"]
            assert np.array_equal(res.count_matrix(k), want["_counts"])
            assert max(len(v) for v in got["prom_to_join"].values()) >= 3
            for grp, v in want["indmeasure"].items():
                assert abs(got["indmeasure"][grp] - v) <= 1e-6


def test_properties_at_scale_and_cross_implementation():
    """Size-independent properties on a 40000 x 512 table (too large for the Python oracle): rank sums, symmetry and
    unit diagonal of the correlation block, tcgen05 == dp4a, and the two independent count kernels
    (de-duplicated bucket kernel vs per-query binary search) agreeing on every bisect index and density."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    N, n = 40000, 512
    t = synth.make_table(N, n, seed=91)
    hits, _, _ = synth.make_permutation_hits(t, 300, 11, window=9000, seed=92)
    ga, gb = synth.global_pair_indices(t.names, 100000, seed=93)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        R = ctx.ranks()
        assert np.all(R.sum(axis=1) == n * (n + 1) / 2.0)                      # exact half-integer sums
        assert R.min() >= 1.0 and R.max() <= n
        bad, first = ctx.selftest_gemm(1024)
        assert bad == 0, first
        blk = ctx.corr_block("spearman", 1024, 256)
        sub = blk[:, 1024:1280]
        assert np.array_equal(np.isnan(sub), np.isnan(sub.T))
        ok = ~np.isnan(sub)
        assert np.array_equal(sub[ok], sub.T[ok])                              # bitwise symmetric
        d = np.diag(sub)
        assert np.all(np.isnan(d) | (np.abs(d - 1.0) < 1e-15))
        assert np.nanmax(np.abs(blk)) <= 1.0 + 1e-15
        g = ctx.pairs(ga, gb, ranked=True)
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        out = {}
        for mode in ("dedup", "per_query", "dedup_sorted", "dedup_fine", "row_rank"):
            ctx.set_count_mode(mode)
            out[mode] = ctx.network_batch(hits, Options(anticorrelation=True), want_scores=True, want_counts=True)
        # the bucket-table kernel with several passes of its counters over the row, with a small table (long chains of
        # statistics per bucket, many statistics next to a bucket edge) and with every third work item handed to the
        # sorted-statistics kernel through the re-do list
        import os
        for tag, env in (("fine_chunked", {"COEX_FINE_CHUNK": "8192"}), ("fine_small_table", {"COEX_FINE_NB": "1024"}),
                         ("fine_redo_list", {"COEX_FINE_DECLINE_EVERY": "3"})):
            os.environ.update(env)
            try:
                ctx.set_count_mode("dedup_fine")
                out[tag] = ctx.network_batch(hits, Options(anticorrelation=True), want_scores=True, want_counts=True)
            finally:
                for k in env:
                    del os.environ[k]
        a = out["dedup"]
        for other in ("per_query", "dedup_sorted", "dedup_fine", "row_rank", "fine_chunked", "fine_small_table", "fine_redo_list"):
            b = out[other]
            assert np.array_equal(a.counts, b.counts), other
            assert np.array_equal(a.scores, b.scores), other
            assert np.array_equal(a.n_groups, b.n_groups) and np.array_equal(a.group_winner, b.group_winner)
            assert np.array_equal(a.group_density, b.group_density), other
        cnt = a.counts[a.counts >= 0]
        assert cnt.min() >= 0 and cnt.max() <= N - 1
        # idx is monotone in the statistic within a query row: higher score <=> larger idx (without -a it is strict)
        assert np.all(a.scores >= 0)


def test_global_correlation_list_row_n2(tmp_path):
    """Row N2 (0-prepare-coex.py:621-716): the list file produced through the resident-table pair kernel holds
    exactly the values the reference C library gives for the same random pairs, ascending, -99 for NaN."""
    import random
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    from coexpression_b200.coexfunctions import make_correlation_list, read_correlation_list
    from oracle import network_oracle as no
    t = synth.make_table(2000, 50, seed=17)
    R = no.rank_rows(t.X)
    path = str(tmp_path / "t.expression.spearmanlist")
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        got = make_correlation_list(ctx, t.names, path, "Spearman", gpl_len=5000, seed=123)
        # a file that already holds gpl_len entries is kept as is (0-prepare-coex.py:629-637: no sub-sampling, new_to_add == 0)
        again = make_correlation_list(ctx, t.names, path, "Spearman", gpl_len=len(got), seed=5)
        # a longer file is sub-sampled to gpl_len (random.sample, :630) and nothing is added
        fewer = make_correlation_list(ctx, t.names, path, "Spearman", gpl_len=1000, seed=5)
    rnd = random.Random(123)
    ia, ib = [], []
    for _ in range(int(5000 * 1.1)):
        a, b = rnd.randint(0, 1999), rnd.randint(0, 1999)
        if t.names[a][:5] != t.names[b][:5]:
            ia.append(a); ib.append(b)
    want = no.get_pearson(R, np.array(ia, dtype=np.int32), np.array(ib, dtype=np.int32))
    want = np.sort(np.where(np.isnan(want), -99.0, want))
    assert np.array_equal(got, want) and np.array_equal(read_correlation_list(path), want)
    assert np.array_equal(again, want)
    assert len(fewer) == 1000 and np.all(np.isin(fewer, want))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_pearson_null_tensor_filter_against_oracle(seed):
    """`-cm Pearson` (Pearson null, Spearman statistic -- quirk Q2) through the de-duplicated tensor-core path: the tensor
    r is only a filter, so every bisect index must equal the oracle's (reference C arithmetic), including duplicated
    rows (exact ties between null values), all-zero / constant rows and a dense cluster."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    rng = np.random.default_rng(300 + seed)
    N, n = int(rng.integers(400, 900)), int(rng.integers(20, 140))
    t = synth.make_table(N, n, seed=400 + seed, zero_row_frac=0.02, dup_rows=10)
    X = t.X.copy()
    X[5] = 2.5                                         # constant row: Pearson and Spearman undefined
    R = no.rank_rows(X)
    ga = rng.integers(0, N, 6000).astype(np.int32); gb = rng.integers(0, N, 6000).astype(np.int32)
    g = no.get_pearson(X, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    hit_sets = []
    for k in range(6):
        h = rng.choice(N, size=int(rng.integers(2, 80)), replace=False)
        if k == 0:
            h = np.concatenate([[5], h])
        hit_sets.append(np.unique(h)[rng.permutation(len(np.unique(h)))].astype(np.int32))
    anticor = bool(seed % 2)
    opts = Options(correlationmeasure="Pearson", anticorrelation=anticor, pval_to_join=0.3)
    with CoexContext() as ctx:
        ctx.set_expression(X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        ctx.set_global_list(glist)
        res = ctx.network_batch(hit_sets, opts, want_scores=True, want_counts=True)
        ctx.set_count_mode("per_query")                # the bit-exact in-order strip
        ref = ctx.network_batch(hit_sets, opts, want_scores=True, want_counts=True)
    assert np.array_equal(res.counts, ref.counts) and np.array_equal(res.scores, ref.scores)
    assert np.array_equal(res.group_density, ref.group_density)
    for k, h in enumerate(hit_sets):
        want = no.make_network(t.names, X, [t.names[r] for r in h], t.addresses(), glist.tolist(), Xrank=R, return_counts=True,
                               correlationmeasure="Pearson", anticorrelation=anticor, pval_to_join=0.3)
        got = perm_objects(res, k, t.names)
        assert np.array_equal(res.count_matrix(k), want["_counts"]), (seed, k)
        assert got["prom_to_join"] == want["prom_to_join"] and got["winners"] == want["winners"]
        for grp, v in want["indmeasure"].items():
            assert abs(got["indmeasure"][grp] - v) <= 1e-6


@pytest.mark.gpu
def test_c4_like_backcirc_anticorrelation_device_mapping():
    """Config C4 in miniature: hit sets drawn on the DEVICE by the backcirc rule from a background file, reduced with -a;
    the hit lists equal the restated reference arithmetic and the densities equal the oracle's."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    from oracle import permute_oracle as po
    t = synth.make_table(2500, 48, seed=61)
    chromlist = list(t.chrom_len)
    cr = po.chromrange_from_lengths(chromlist, t.chrom_len)
    rng = np.random.default_rng(62)
    snps = []
    for r in rng.integers(0, t.N, size=90):
        c = t.chrom[r]
        snps.append((c, int(np.clip(t.start[r] + rng.integers(-3000, 3000), 1, t.chrom_len[c] - 2))))
    snps = list(dict.fromkeys(snps))
    backdict = {"%s_%s" % (c, a): po.get_gwad(c, a, cr) for c, a in snps}
    for r in rng.integers(0, t.N, size=4000):           # background variants near features, so that permuted sets map
        c = t.chrom[r]
        a = int(np.clip(t.start[r] + rng.integers(-20000, 20000), 1, t.chrom_len[c] - 2))
        backdict["%s_%s" % (c, a)] = po.get_gwad(c, a, cr)
    sb = po.sorted_background(backdict)
    background = [(k.split("_")[0], int(k.split("_")[1])) for k in sb]
    shifts = [0] + [int(rng.random() * 10000000000) for _ in range(4)]
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 15000, seed=63)
    g = no.get_pearson(R, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    opts = Options(anticorrelation=True)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        ctx.set_global_list(glist)
        hits = ctx.permute_map("backcirc", [c for c, _ in snps], [a for _, a in snps], shifts, chromlist, cr, 5000,
                               background=background)
        res = ctx.network_batch(hits, opts, want_counts=True)
    for k, s in enumerate(shifts):
        pos = po.circular_positions(snps, 0, cr, chromlist) if s == 0 else po.backcirc_positions(snps, s, sb)
        want_rows = po.hit_rows(po.window_bed(pos, t.chrom, t.start, t.end, 5000))
        assert hits[k].tolist() == want_rows and len(want_rows) > 5, (k, len(want_rows))
        want = no.make_network(t.names, t.X, [t.names[r] for r in want_rows], t.addresses(), glist.tolist(), Xrank=R,
                               return_counts=True, anticorrelation=True)
        got = perm_objects(res, k, t.names)
        assert np.array_equal(res.count_matrix(k), want["_counts"])
        assert got["prom_to_join"] == want["prom_to_join"] and got["winners"] == want["winners"]
        for grp, v in want["indmeasure"].items():
            assert abs(got["indmeasure"][grp] - v) <= 1e-6


@pytest.mark.gpu
def test_large_hit_set_shared_memory_limits():
    """One hit set of 4500 regions (config C5 territory: wide -w windows): the kernels whose shared memory scales with the
    hit count (statistics capacity of the count kernels, sort / scan buffers of join_nearby, removal) at a size beyond
    4096, checked by the three independent count kernels agreeing on every index, score, group and density."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    t = synth.make_table(9000, 24, seed=71)
    rng = np.random.default_rng(72)
    hits = [rng.choice(t.N, size=4500, replace=False).astype(np.int32), rng.choice(t.N, size=300, replace=False).astype(np.int32),
            rng.choice(t.N, size=2, replace=False).astype(np.int32)]
    ga, gb = synth.global_pair_indices(t.names, 30000, seed=73)
    out = {}
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        g = ctx.pairs(ga, gb, ranked=True)
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        for mode in ("dedup", "per_query", "dedup_sorted", "row_rank"):      # (the bucket-table kernel takes at most 4096 hits per set)
            ctx.set_count_mode(mode)
            out[mode] = ctx.network_batch(hits, Options(pval_to_join=0.01), want_scores=True, want_counts=True)
    a = out["dedup"]
    assert int(a.n_groups[0]) > 100
    for other in ("per_query", "dedup_sorted", "row_rank"):
        b = out[other]
        assert np.array_equal(a.counts, b.counts), other
        assert np.array_equal(a.scores, b.scores), other
        assert np.array_equal(a.n_groups, b.n_groups) and np.array_equal(a.group_of_hit, b.group_of_hit)
        assert np.array_equal(a.group_winner, b.group_winner) and np.array_equal(a.group_density, b.group_density)


# ---------------------------------------------------------------------------------------------
# Round 2: the timed configurations themselves against the oracle
# ---------------------------------------------------------------------------------------------
def _oracle_vs_gpu(t, hits, opts, glist, R, ctx, **okw):
    from coexpression_b200.results import perm_objects
    from oracle import network_oracle as no
    res = ctx.network_batch(hits, opts, want_scores=True, want_counts=True)
    addresses = t.addresses()
    worst = 0.0
    for k, h in enumerate(hits):
        want = no.make_network(t.names, t.X, [t.names[r] for r in h], addresses, glist.tolist(), Xrank=R, return_counts=True,
                               anticorrelation=opts.anticorrelation, pval_to_join=opts.pval_to_join,
                               correlationmeasure=opts.correlationmeasure, **okw)
        got = perm_objects(res, k, t.names)
        assert got["prom_to_join"] == want["prom_to_join"], k
        assert got["winners"] == want["winners"], k
        if len(h) > 1:
            assert np.array_equal(res.count_matrix(k), want["_counts"]), k
        assert set(got["indmeasure"]) == set(want["indmeasure"])
        for grp, v in want["indmeasure"].items():
            worst = max(worst, abs(got["indmeasure"][grp] - v))
        for p_, v in want["allpromoterscores"].items():
            worst = max(worst, abs(got["allpromoterscores"][p_] - v))
        if opts.correlationmeasure == "Spearman" and opts.nsp == 1.0 and len(h) > 1:
            # the edge scores come from a table the host builds with the C library's log, as -math.log(p, 10) does: BIT-identical
            S = res.score_matrix(k)
            nm = [t.names[r] for r in h]
            W = np.array([[0.0 if i == j else want["individualscores"][nm[i]][nm[j]] for j in range(len(h))] for i in range(len(h))])
            assert np.array_equal(S, W), k
    assert worst <= 1e-6, worst                                # north_star tolerance
    return res


def test_network_at_1829_samples_default_gemm_against_oracle():
    """The FANTOM5 row length through the kernels the full-shape run uses (n >= 1024 -> gemm_i8_tc_persistent128_kernel,
    count_dedup float filter at sigma = 1 / sqrt(1829)): 3000 x 1829 table, 3 hit sets, every bisect index vs the oracle."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from oracle import network_oracle as no
    t = synth.make_table(3000, 1829, seed=1829)
    hits, _, _ = synth.make_permutation_hits(t, 120, 2, window=9000, seed=18)
    assert min(len(h) for h in hits) > 20
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 60000, seed=29)
    g = no.get_pearson(R, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    for opts in (Options(), Options(anticorrelation=True)):
        with CoexContext() as ctx:
            ctx.set_expression(t.X); ctx.prepare()
            assert np.array_equal(ctx.ranks(), R)
            ctx.set_features(t.chrom, t.start, t.end, t.names)
            ctx.set_global_list(glist)
            _oracle_vs_gpu(t, hits, opts, glist, R, ctx)


def test_bench_c2_inputs_two_hit_sets_against_oracle():
    """bench.py's own default workload (20000 x 256, seed 1234, 500 loci, -w 10000): the real set and one permuted set
    against the oracle, alone and inside the 101-set batch the benchmark times (de-duplication regime)."""
    import argparse
    import bench
    from coexpression_b200.api import CoexContext, Options
    from oracle import network_oracle as no
    args = argparse.Namespace(shape="minimal", rows=0, samples=0, loci=500, perms=100, window=10000, seed=1234)
    t, hits, ga, gb = bench.workload(args, 0)
    assert len(hits) == 101 and len(hits[0]) > 1000
    R = no.rank_rows(t.X)
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        g = ctx.pairs(ga, gb, ranked=True)
        assert _same_doubles(g, no.get_pearson(R, ga, gb))
        glist = np.sort(np.where(np.isnan(g), -99.0, g))
        ctx.set_global_list(glist)
        small = _oracle_vs_gpu(t, [hits[0], hits[1]], Options(), glist, R, ctx)
        big = ctx.network_batch(hits, Options(), want_scores=False, want_counts=True)
    for k in (0, 1):
        assert np.array_equal(big.count_matrix(k), small.count_matrix(k))
        assert np.array_equal(big.perm(k)["density"], small.perm(k)["density"])       # bit for bit, whatever the batch
        assert np.array_equal(big.perm(k)["winner"], small.perm(k)["winner"])


def test_pearson_null_more_than_1860_samples_takes_the_exact_strip():
    """with -cm Pearson the de-duplicated path reads its (Spearman, quirk Q2) statistics from the int32 strip,
    which overflows for rows of more than 1860 samples; such tables must take the 64-bit statistic kernel."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200._lib import CoexError
    from oracle import network_oracle as no
    t = synth.make_table(700, 2100, seed=2100, zero_row_frac=0.01)
    # strongly correlated rows: |S_ab| exceeds 2^31 for |r| > 0.7 at n = 2100
    t.X[10] = t.X[11] * 3.0 + 0.01
    hits = [np.array([10, 11, 40, 41, 300, 301, 302, 650], dtype=np.int32), np.arange(5, 60, 3, dtype=np.int32)]
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 20000, seed=3)
    g = no.get_pearson(t.X, ga, gb)
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    with CoexContext() as ctx:
        ctx.set_expression(t.X); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        ctx.set_global_list(glist)
        _oracle_vs_gpu(t, hits, Options(correlationmeasure="Pearson"), glist, R, ctx)
        with pytest.raises(CoexError):                          # Spearman null: refused loudly, never wrong silently
            ctx.network_batch(hits, Options())


def test_shared_job_entry_points_on_one_rank_equal_the_plain_call():
    """coex_shard_count + coex_shard_finish with world = 1 (score rows through the per-permutation pointer table, network stage
    on the job's own buffer) against coex_network_batch: every output bit for bit, scores included."""
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.dist import SharedJob
    t = synth.make_table(4000, 96, seed=61)
    hits, _, _ = synth.make_permutation_hits(t, 150, 6, window=7000, seed=62)
    hits = hits + [np.zeros(0, dtype=np.int32), hits[0][:1]]
    ga, gb = synth.global_pair_indices(t.names, 40000, seed=63)
    for opts in (Options(), Options(anticorrelation=True, noiter=True), Options(correlationmeasure="Pearson")):
        with CoexContext() as ctx:
            ctx.set_expression(t.X); ctx.prepare()
            ctx.set_features(t.chrom, t.start, t.end, t.names)
            g = ctx.pairs(ga, gb, ranked=opts.correlationmeasure == "Spearman")
            ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
            want = ctx.network_batch(hits, opts, want_scores=True)
            for k in (0, 3):                                     # coex_read_scores: one block of the last batch
                assert np.array_equal(ctx.read_scores(k, len(hits[k])), want.score_matrix(k))
            ranks = ctx.ranks()
            job = SharedJob(ctx, hits, None)
            assert job.mine == list(range(len(hits)))
            got = job.run(opts, want_scores=True)
            job.prepare()                                        # the rank transform through the shared-job entry points
            assert np.array_equal(ctx.ranks(), ranks)
            again = job.run(opts, want_scores=True)              # the buffers are reused
        for res in (got, again):
            for name in ("n_groups", "group_of_hit", "group_label", "group_winner", "group_density", "hit_score", "scores"):
                assert np.array_equal(getattr(res, name), getattr(want, name)), name


def test_collate_on_device_equals_the_reference_arithmetic():
    """Row A14 (2-collate-results.py:67-76,142,161-163): pooled null, empirical p, BH / BY on the device, bit for bit against
    the host mirror of the reference's functions (which tests/test_reference_pin.py pins to the executed reference)."""
    from coexpression_b200.api import CoexContext
    from coexpression_b200.collate import collate_numbers
    rng = np.random.default_rng(14)
    with CoexContext() as ctx:
        for n_pool, G in ((0, 5), (1, 1), (1000, 37), (250000, 612), (5000, 1500)):
            pool = np.round(rng.gamma(2.0, 30.0, size=n_pool), 2)              # ties inside the pool
            real = np.round(rng.gamma(2.0, 35.0, size=G), 2)
            if G > 3 and n_pool:
                real[0] = real[1]                                               # equal real densities
                real[2] = pool[0]                                               # a real density equal to pool values
                real[3] = pool.max() + 1.0                                      # above everything: p = 0 -> FDR 0
            want = collate_numbers(pool.tolist(), {"indmeasure": {str(i): float(v) for i, v in enumerate(real)},
                                                   "indjoined": {str(i): {} for i in range(G)}})
            got = ctx.collate(pool, real)
            assert got["p"] == want["p"], (n_pool, G)
            assert got["fdr_bh"] == [float(x) for x in want["fdr_bh"]], (n_pool, G)
            assert got["fdr_by"] == [float(x) for x in want["fdr_by"]], (n_pool, G)


def test_pipelined_upload_equals_plain_upload():
    """coex_set_expression(on_device = 2): the table goes out in row chunks from page-locked memory and coex_prepare ranks every
    chunk as it lands; ranks, row statistics (through the legacy pair kernel) and a network result equal the plain upload's."""
    import torch
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext, Options
    t = synth.make_table(3001, 97, seed=5)                   # (rows not divisible by the eight chunks)
    hits, _, _ = synth.make_permutation_hits(t, 60, 3, window=20000, seed=6)
    ga, gb = synth.global_pair_indices(t.names, 20000, seed=7)
    X_pin = torch.from_numpy(t.X).pin_memory()
    out = []
    for pipelined in (False, True, True):                    # (twice: a second pipelined upload into the same context)
        ctx = CoexContext() if not out or not pipelined else ctx
        ctx.set_expression(X_pin.numpy(), pipelined=pipelined); ctx.prepare()
        ctx.set_features(t.chrom, t.start, t.end, t.names)
        g = ctx.pairs(ga, gb, ranked=False)
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        res = ctx.network_batch(hits, Options(), want_scores=True)
        out.append((ctx.ranks(), g, res.scores.copy(), res.group_density.copy()))
    ctx.close()
    for o in out[1:]:
        for a, b in zip(out[0], o):
            assert np.array_equal(a, b, equal_nan=True)
