"""Pins the oracle (test infrastructure) before anything trusts it: CPU only."""
import json
import math
import os

import numpy as np
import pytest

from coexpression_b200 import synth
from oracle import network_oracle as no

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


@pytest.mark.parametrize("name", ["getpearson_raw", "getpearson_ranked", "getpearson_onepair"])
@pytest.mark.parametrize("kind", ["port", "reference"])
def test_getpearson_matches_reference_golden(name, kind):
    if kind not in no.available_kinds():
        pytest.skip(f"{kind} build absent")
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    r = no.get_pearson(g["X"], g["ia"], g["ib"], kind)
    assert np.array_equal(np.isnan(r), np.isnan(g["r"]))
    assert np.array_equal(_bits(r), _bits(g["r"]))            # bit-exact incl. NaN payload sign


def test_port_bit_exact_with_reference_build_random():
    if "reference" not in no.available_kinds():
        pytest.skip("reference build absent (GPU box): golden vectors cover it")
    rng = np.random.default_rng(7)
    X = rng.lognormal(0, 2, size=(257, 129)) * (rng.random((257, 129)) < 0.6)
    X[3] = 0.0
    ia = rng.integers(0, 257, 20000).astype(np.int32)
    ib = rng.integers(0, 257, 20000).astype(np.int32)
    for data in (X, no.rank_rows(X), X - 1.0):
        a = no.get_pearson(data, ia, ib, "reference")
        b = no.get_pearson(data, ia, ib, "port")
        assert np.array_equal(_bits(a), _bits(b))


def test_getpearson_zero_pairs():
    X = np.ones((3, 4))
    assert no.get_pearson(X, np.zeros(0, np.int32), np.zeros(0, np.int32), "port").shape == (0,)


def test_rank_rows_golden():
    g = np.load(os.path.join(GOLDEN, "ranks_small.npz"))
    R = no.rank_rows(g["X"])
    assert np.array_equal(R, g["ranks"])
    assert np.array_equal(R[0], np.full(R.shape[1], (R.shape[1] + 1) / 2.0))       # all tied
    assert np.array_equal(R[1], R[0])                                              # -0.0 ties 0.0
    assert np.all(R.sum(axis=1) == R.shape[1] * (R.shape[1] + 1) / 2.0)            # exact row sums


def test_spearman_sums_are_exact_integers():
    """The design relies on it: for rank rows the reference's xy/xx/yy sums are exact, so
    r = (S_ab/4) / (sqrt(S_aa/4) * sqrt(S_bb/4)) reproduces getPearson bit for bit."""
    g = np.load(os.path.join(GOLDEN, "getpearson_ranked.npz"))
    R, ia, ib = g["X"], g["ia"], g["ib"]
    n = R.shape[1]
    U = (2 * R - (n + 1)).astype(np.int64)
    assert np.array_equal(U, 2 * R - (n + 1))
    S = np.einsum("ij,ij->i", U[ia], U[ib])
    Saa = (U * U).sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = (S * 0.25) / (np.sqrt(Saa[ia] * 0.25) * np.sqrt(Saa[ib] * 0.25))
    assert np.array_equal(_bits(np.where(np.isnan(r), np.nan, r)), _bits(np.where(np.isnan(g["r"]), np.nan, g["r"])))


def _load_case(name):
    with open(os.path.join(GOLDEN, f"network_{name}.json")) as f:
        case = json.load(f)
    cm = case["kwargs"].get("correlationmeasure", "Spearman")
    inp = np.load(os.path.join(GOLDEN, f"network_inputs_{case['shape']}_{cm}.npz"))
    return case, inp


NETWORK_CASES = ["spearman_default", "spearman_anticor", "spearman_noiter_j001", "spearman_useavg_m",
                 "spearman_global_nsp0", "pearson_default", "pearson_anticor_j05"]


@pytest.mark.parametrize("name", NETWORK_CASES)
def test_network_oracle_reproduces_golden(name):
    case, inp = _load_case(name)
    names = [str(s) for s in inp["names"]]
    addresses = {nm: (str(c), int(s), int(e)) for nm, c, s, e in zip(names, inp["chrom"], inp["start"], inp["end"])}
    R = no.rank_rows(inp["X"])
    glist = inp["glist"].tolist()
    for perm in case["perms"]:
        res = no.make_network(names, inp["X"], [names[r] for r in perm["hit_rows"]], addresses, glist,
                              mode="consistent", Xrank=R, return_counts=True, **case["kwargs"])
        assert res["prom_to_join"] == perm["prom_to_join"]
        assert res["winners"] == perm["winners"]
        assert np.array_equal(res["_counts"], np.array(perm["counts"]).reshape(res["_counts"].shape))
        assert set(res["indmeasure"]) == set(perm["indmeasure"])
        for g, v in perm["indmeasure"].items():
            assert res["indmeasure"][g] == v


def _join_nearby_pandas_networkx(proms, addresses, globcors, r_of_pair, pvtj, softsep, anticor):
    """Transliteration of the reference's join_nearby control flow on pandas + networkx
    (coexfunctions.py:277-316) used only to pin the plain-Python restatement."""
    import networkx as nx
    import pandas as pd
    from bisect import bisect_left
    ptj = {}
    a = pd.DataFrame({k: addresses[k] for k in proms}, index=["chromosome", "start", "end"]).T
    a = a.astype({"start": "int64", "end": "int64"})
    a = a.sort_values(["chromosome", "start", "end"])
    for chrom in pd.unique(a.chromosome):
        sub = a[a.chromosome == chrom][["start", "end"]].values
        for g0, g1 in no.merge_ranges([tuple(x) for x in sub], softsep):
            members = list(a[(a.chromosome == chrom) & (a.start >= g0) & (a.start < g1)].index)
            gph = nx.Graph()
            gph.add_nodes_from(members)
            for i in range(len(members)):
                for j in range(i + 1, len(members)):
                    r = r_of_pair(members[i], members[j])
                    idx = bisect_left(globcors, r)
                    p1 = 1 - float(idx) / len(globcors)
                    p = min(p1, float(idx) / len(globcors)) if anticor else p1
                    if p < pvtj:
                        gph.add_edge(members[i], members[j])
            for comp in nx.connected_components(gph):
                new = sorted(comp)
                ptj[new[0]] = new
    return ptj


@pytest.mark.parametrize("anticor,pvtj", [(False, 0.1), (True, 0.3), (False, 1.5)])
def test_join_nearby_against_pandas_networkx(anticor, pvtj):
    case, inp = _load_case("spearman_default")
    names = [str(s) for s in inp["names"]]
    addresses = {nm: (str(c), int(s), int(e)) for nm, c, s, e in zip(names, inp["chrom"], inp["start"], inp["end"])}
    R = no.rank_rows(inp["X"])
    corr = no.Correlator(inp["X"], R, "consistent")
    row_of = {nm: i for i, nm in enumerate(names)}
    glist = inp["glist"].tolist()
    proms = [names[r] for r in case["perms"][0]["hit_rows"]]
    mine = no.join_nearby(proms, addresses, glist, corr, row_of, pvtj, "Spearman", 100000, anticor)
    theirs = _join_nearby_pandas_networkx(
        proms, addresses, glist, lambda a, b: float(corr.many([row_of[a]], [row_of[b]])[0]), pvtj, 100000, anticor)
    assert mine == {k: theirs[k] for k in sorted(theirs)}
    assert any(len(v) > 1 for v in mine.values())


def test_edge_score_quirks():
    # Q4: p == 0 scores 0, not infinity (1-make-network.py:311-314)
    assert no.edge_score(100, 100, False) == 0
    # -a: r at the very bottom of the list -> min(p1, 0) = 0 -> 0
    assert no.edge_score(0, 100, True) == 0
    assert no.edge_score(0, 100, False) == -math.log(1, 10)
    assert no.edge_score(50, 100, False) == -math.log(0.5, 10)


def test_q5_zero_rows_give_minus99_and_p1():
    t = synth.make_table(200, 24, seed=3)
    X = t.X.copy()
    X[10] = 0.0
    hits = [10, 11, 12, 50]
    R = no.rank_rows(X)
    res = no.make_network(t.names, X, [t.names[r] for r in hits], t.addresses(), [-99.0, 0.0, 0.5],
                          Xrank=R, return_counts=True)
    assert np.all(res["_stat"][0, 1:] == -99) and np.all(res["_stat"][1:, 0] == -99)
    for other in hits[1:]:
        assert res["individualscores"][t.names[10]][t.names[other]] == 0


def test_q7_iterative_removal_tie_reuse():
    """Two groups with bit-identical densities: the second inherits the first's whittled value."""
    t = synth.make_table(200, 24, seed=4)
    X = t.X.copy()
    X[21] = X[20]                       # duplicate rows, far apart on the genome -> separate groups
    hits = [20, 21, 60, 100, 140]
    res_i = no.make_network(t.names, X, [t.names[r] for r in hits], t.addresses(), [0.0], noiter=True)
    res = no.make_network(t.names, X, [t.names[r] for r in hits], t.addresses(), [0.0], noiter=False)
    assert set(res["indmeasure"]) == set(res_i["indmeasure"])
    order = sorted(res_i["indmeasure"].items(), key=lambda kv: (kv[1], kv[0]), reverse=True)
    for (g0, s0), (g1, s1) in zip(order, order[1:]):
        if s0 == s1:
            assert res["indmeasure"][g1] == res["indmeasure"][g0]
    # whittled densities never exceed the un-whittled ones
    for g in res["indmeasure"]:
        assert res["indmeasure"][g] <= res_i["indmeasure"][g] + 1e-12


def test_fdr_and_empiricalp():
    p = [0.01, 0.04, 0.03, 0.20]
    rej, adj = no.fdr_correction(p)
    assert np.allclose(adj, [0.04, 0.05333333333333334, 0.05333333333333334, 0.2])
    assert list(rej) == [True, False, False, False]
    _, adj_by = no.fdr_correction(p, 0.05, "negcorr")
    cm = 1 + 1 / 2 + 1 / 3 + 1 / 4
    assert np.allclose(adj_by, np.minimum(1.0, np.array([0.04, 0.05333333333333334, 0.05333333333333334, 0.2]) * cm))
    assert no.empiricalp(0.5, []) == 1
    assert no.empiricalp(2.0, [1.0, 2.0, 2.0, 3.0]) == 0.75
    ps, by, bh = no.collate({"a": 3.0, "b": 0.5}, [{"x": 1.0, "y": 2.0}, {"z": 0.1}])
    assert ps == [0.0, 1 - 1 / 3]


def test_faithful_vs_consistent_is_a_reference_property():
    """Q3: scipy-vs-C ulp-level noise moves bisect indices (by one, or across a run of tied
    null entries); counts can only differ where the two statistics differ bitwise."""
    case, inp = _load_case("spearman_default")
    names = [str(s) for s in inp["names"]]
    addresses = {nm: (str(c), int(s), int(e)) for nm, c, s, e in zip(names, inp["chrom"], inp["start"], inp["end"])}
    R = no.rank_rows(inp["X"])
    proms = [names[r] for r in case["perms"][1]["hit_rows"]]
    a = no.make_network(names, inp["X"], proms, addresses, inp["glist"].tolist(), mode="consistent", Xrank=R,
                        return_counts=True)
    b = no.make_network(names, inp["X"], proms, addresses, inp["glist"].tolist(), mode="faithful", Xrank=R,
                        return_counts=True)
    assert np.abs(a["_stat"] - b["_stat"])[a["_stat"] > -99].max() < 1e-12
    moved = a["_counts"] != b["_counts"]
    assert not np.any(moved & (a["_stat"] == b["_stat"]))
    assert 0 < moved.mean() < 0.6


def test_query_counts_is_one_row_of_make_network():
    """oracle.query_counts (the sampled check of bench.py at the full FANTOM5 shape) == row i of make_network's
    bisect-index matrix and of individualscores, including -99 statistics (all-zero hit row) and quirk Q1."""
    from coexpression_b200 import synth
    t = synth.make_table(600, 48, seed=3)
    hits, _, _ = synth.make_permutation_hits(t, 40, 1, window=8000, seed=4)
    zero = int(np.flatnonzero(t.X.sum(axis=1) == 0)[0])
    h = np.concatenate([hits[1], [zero]]).astype(np.int32)
    R = no.rank_rows(t.X)
    ga, gb = synth.global_pair_indices(t.names, 5000, seed=2)
    g = no.get_pearson(R, ga, gb)
    gl = np.sort(np.where(np.isnan(g), -99.0, g)).tolist()
    for anticor in (False, True):
        w = no.make_network(t.names, t.X, [t.names[r] for r in h], t.addresses(), gl, Xrank=R, return_counts=True,
                            anticorrelation=anticor)
        for i in range(len(h)):
            c, s = no.query_counts(t.X, R, h, i, anticorrelation=anticor)
            assert np.array_equal(c, w["_counts"][i])
            want = np.array([0.0 if j == i else w["individualscores"][t.names[h[i]]][t.names[h[j]]] for j in range(len(h))])
            assert np.array_equal(s, want)
