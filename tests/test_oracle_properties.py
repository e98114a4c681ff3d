"""Property tests (hypothesis) of the restated arithmetic: CPU only, small cases."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import network_oracle as no


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(min_value=-5, max_value=5, allow_nan=False, width=32), min_size=2, max_size=40))
def test_rank_rows_matches_definition(values):
    x = np.array([values], dtype=np.float64)
    r = no.rank_rows(x)[0]
    for i, v in enumerate(values):
        below = sum(1 for w in values if w < v)
        equal = sum(1 for w in values if w == v)
        assert r[i] == below + (equal + 1) / 2.0
    assert r.sum() == len(values) * (len(values) + 1) / 2.0


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 30), st.integers(3, 12), st.integers(0, 10 ** 6))
def test_getpearson_port_is_symmetric_and_bounded(rows, n, seed):
    rng = np.random.default_rng(seed)
    X = np.round(rng.gamma(0.5, 2.0, size=(rows, n)), 1)
    ia = rng.integers(0, rows, 50).astype(np.int32); ib = rng.integers(0, rows, 50).astype(np.int32)
    a = no.get_pearson(X, ia, ib, "port"); b = no.get_pearson(X, ib, ia, "port")
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    assert np.array_equal(a[ok], b[ok])                       # bitwise symmetric: the products commute
    assert np.all(np.abs(a[ok]) <= 1.0 + 1e-12)
    same = ia == ib
    assert np.all(np.isnan(a[same]) | (np.abs(a[same] - 1.0) < 1e-12))


@settings(max_examples=80, deadline=None)
@given(st.integers(1, 500), st.integers(0, 500), st.booleans())
def test_edge_score_rules(length, idx, anticor):
    idx = min(idx, length)
    s = no.edge_score(idx, length, anticor)
    p1 = 1 - float(idx) / length
    if anticor:
        p1 = min(p1, float(idx) / length)
    if p1 == 0:
        assert s == 0                                          # quirk Q4
    else:
        assert s == -math.log(p1, 10) and s >= 0


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(min_value=0, max_value=1, allow_nan=False), min_size=1, max_size=30))
def test_fdr_correction_properties(pvals):
    for method in ("indep", "negcorr"):
        rej, adj = no.fdr_correction(pvals, 0.05, method)
        assert np.all(adj >= np.asarray(pvals) - 1e-15) and np.all(adj <= 1.0)
        order = np.argsort(pvals, kind="stable")
        assert np.all(np.diff(adj[order]) >= -1e-15)          # monotone in p
    _, bh = no.fdr_correction(pvals, 0.05, "indep")
    _, by = no.fdr_correction(pvals, 0.05, "negcorr")
    assert np.all(by >= bh - 1e-15)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 10 ** 6), st.integers(0, 500)), min_size=1, max_size=25), st.integers(1, 200000))
def test_merge_ranges_partitions_sorted_starts(raw, ss):
    rng = sorted((s, s + w) for s, w in raw)
    merged = list(no.merge_ranges(rng, ss))
    assert all(a <= b for a, b in merged)
    assert all(merged[i][1] < merged[i + 1][0] for i in range(len(merged) - 1))       # gaps between clusters
    for s, e in rng:                                                               # every range in exactly one
        assert sum(1 for a, b in merged if a <= s < b) == 1
