"""CPU-only: the host-side writer of a permutation's two large result objects (coex_hostio.so,
include/coexpression_b200_io.h) produces the text of the reference's json.dump calls (1-make-network.py:523-540,
Python 3) byte for byte, and make_network.write_result_objects writes the same eight files as the dict route."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

from coexpression_b200 import hostio
from coexpression_b200.api import BatchResult
from coexpression_b200.make_network import permutation_number, result_path, write_result_objects
from coexpression_b200.results import perm_objects

HEADER = os.path.join(ROOT, "include", "coexpression_b200_io.h")


def test_io_header_symbols_exported_and_bound():
    hostio.build()
    hostio.load()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    names = set(re.findall(r"\b(coex_io_[a-z_0-9]+)\s*\(", src))
    exported = subprocess.check_output(["nm", "-D", "--defined-only", hostio.SO_PATH], text=True)
    exported = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert names == exported == set(hostio.SIGNATURES) and len(names) == 3


def test_numbers_are_written_as_python_repr():
    rng = np.random.default_rng(11)
    v = rng.integers(0, 2 ** 64, size=300000, dtype=np.uint64).view(np.float64)
    v = v[np.isfinite(v)]
    assert hostio.format_doubles(v) == [repr(x) for x in v.tolist()]
    scores = -np.log10(rng.random(200000))                       # what the files hold: -log10(p)
    edge = [0.0, -0.0, 1.0, -1.0, 1e15, 1e16, 9999999999999998.0, 123456789012345680.0, 1e-4, 1e-5, 0.00012345,
            5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 0.1, 0.30000000000000004, 1 / 3, 1e22, 1e23, 1e100, 1.5e-100]
    pw = np.concatenate([10.0 ** np.arange(-30, 31), -(10.0 ** np.arange(-30, 31)), 2.0 ** np.arange(-60, 61)])
    for vals in (scores, np.array(edge), pw, np.nextafter(pw, np.inf), np.nextafter(pw, -np.inf)):
        assert hostio.format_doubles(vals) == [repr(x) for x in vals.tolist()]
    assert hostio.format_doubles([float("nan"), float("inf"), -float("inf")]) == ["NaN", "Infinity", "-Infinity"]
    assert json.dumps([float("nan"), float("inf"), -float("inf")]) == "[NaN, Infinity, -Infinity]"
    assert hostio.format_doubles([]) == []


@pytest.mark.parametrize("R", [0, 1, 2, 37])
def test_matrix_text_equals_json_dump(tmp_path, R):
    rng = np.random.default_rng(R)
    big = -np.log10(rng.random((R + 3, R + 5)))
    S = big[1:R + 1, 2:R + 2]                                    # a view: row stride > R
    if R > 1:
        S[0, 1] = -0.0
        S[1, 0] = 0.0
    labels = ["chr%d:%d..%d,%s" % (i % 22 + 1, 1000 * i, 1000 * i + 300, "+-"[i % 2]) for i in range(R)]
    if R > 2:
        labels[2] = 'p1@GENE"quoted"\\back\u00e9\tTab'
    for skip in (True, False):
        want = json.dumps({labels[i]: {labels[j]: float(S[i, j]) for j in range(R) if not (skip and j == i)} for i in range(R)})
        path = tmp_path / ("m%d" % skip)
        n = hostio.write_json_matrix(str(path), S, labels, skip_diag=skip)
        text = path.read_bytes()
        assert text == want.encode("ascii") and n == len(text)
    if R > 1:
        with pytest.raises(ValueError):
            hostio.write_json_matrix(str(tmp_path / "dup"), S, [labels[0]] * R)
        with pytest.raises(OSError):
            hostio.write_json_matrix(str(tmp_path / "no_such_dir" / "x"), S, labels)
        path = tmp_path / "T"
        hostio.write_json_matrix(str(path), S.T, labels)           # column-major view: copied, same text as the dict route
        assert json.loads(path.read_text()) == {labels[i]: {labels[j]: float(S[j, i]) for j in range(R) if j != i} for i in range(R)}


def _fake_batch(seed, sizes):
    """A BatchResult as coex_network_batch fills it (api.BatchResult), made up on the host."""
    rng = np.random.default_rng(seed)
    perm_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    H = int(perm_off[-1])
    hit_rows = np.concatenate([rng.permutation(500)[:P] for P in sizes] + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    n_groups = np.zeros(len(sizes), dtype=np.int32)
    group_of_hit = np.zeros(H, dtype=np.int32)
    label = np.zeros(H, dtype=np.int32)
    winner = np.zeros(H, dtype=np.int32)
    density = rng.random(H) * 40
    for k, P in enumerate(sizes):
        a = int(perm_off[k])
        if P < 2:
            continue
        G = max(1, P - P // 3)
        g_of = np.concatenate([np.arange(G), rng.integers(0, G, size=P - G)])
        rng.shuffle(g_of)
        n_groups[k] = G
        group_of_hit[a:a + P] = g_of
        for g in range(G):
            members = np.nonzero(g_of == g)[0]
            label[a + g] = members[0]
            winner[a + g] = members[-1]
    scores = -np.log10(rng.random(int(np.sum(np.asarray(sizes, dtype=np.int64) ** 2))))
    return BatchResult(perm_off=perm_off, hit_rows=hit_rows, n_groups=n_groups, group_of_hit=group_of_hit, group_label=label,
                       group_winner=winner, group_density=density, hit_score=rng.random(H) * 9, scores=scores, counts=None)


def test_result_files_equal_the_dict_route(tmp_path):
    names = ["chr%d:%d..%d,+" % (i % 22 + 1, 700 * i, 700 * i + 250) for i in range(500)]
    sizes = [23, 1, 0, 2, 40]
    res = _fake_batch(5, sizes)
    store = tmp_path / "perm_results_store"
    store.mkdir()
    for k, P in enumerate(sizes):
        mapfile = str(tmp_path / ("f5ep.%d.bed" % k))
        snp_map = {names[r]: {"snps": "rs%d" % r} for r in res.perm(k)["hit_rows"]}
        written = write_result_objects(str(store), mapfile, res, k, names, snp_map, True)
        want = perm_objects(res, k, names, with_scores=True)
        want["snp_map"] = snp_map
        order = ["indmeasure", "prom_to_join", "joining_instructions", "winners", "allpromoterscores", "individualscores",
                 "indjoined", "snp_map"]
        assert written == [result_path(str(store), mapfile, obj) for obj in order]
        for obj in order:
            assert open(result_path(str(store), mapfile, obj)).read() == json.dumps(want[obj]), (k, obj)
        if P >= 2:
            assert len(want["individualscores"]) == P and len(want["indjoined"]) == int(res.n_groups[k])
    # without the score block: six objects, no P x P files
    lean = write_result_objects(str(tmp_path), str(tmp_path / "x.7.bed"), res, 0, names, {}, False)
    assert len(lean) == 6 and not any("individualscores" in p or "indjoined" in p for p in lean)


def test_permutation_number_of_a_map_file():
    assert permutation_number("/a/b/f5ep300_100.0.bed") == 0
    assert permutation_number("permutation_store/f5ep.17.bed") == 17
    assert permutation_number("whatever.bed") == -1


def test_session_run_chunks_and_lean_mode_on_a_stand_in_context(tmp_path):
    """Session.run (make_network.py) without a GPU: the context is a stand-in that hands back made-up results.  Checked:
    which calls ask for the score blocks, that a byte budget splits the map files, and which files a lean run writes."""
    from coexpression_b200.make_network import Session

    names = ["chr%d:%d..%d,+" % (i % 22 + 1, 700 * i, 700 * i + 250) for i in range(500)]

    class Table:
        pass

    class Ctx:
        def __init__(self):
            self.calls = []

        def network_batch(self, hit_lists, options, want_scores=True):
            self.calls.append(([len(h) for h in hit_lists], want_scores))
            res = _fake_batch(len(self.calls), [len(h) for h in hit_lists])
            res.hit_rows = np.concatenate([np.asarray(h, dtype=np.int32) for h in hit_lists])
            if not want_scores:
                res.scores = None
            return res

    sess = Session.__new__(Session)
    sess.table = Table()
    sess.table.names = names
    sess.options = None
    sess.ctx = Ctx()
    rng = np.random.default_rng(3)
    hits = {("f5ep.%d.bed" % k): rng.permutation(500)[:P] for k, P in enumerate([30, 12, 20, 25])}
    sess.hits_of = lambda mf: ({names[r]: {"p": -9999, "snps": "rs%d" % r} for r in hits[os.path.split(mf)[-1]]},
                               hits[os.path.split(mf)[-1]].astype(np.int32))
    mapfiles = [str(tmp_path / nm) for nm in hits]

    full = sess.run(mapfiles, str(tmp_path / "full"))
    assert sess.ctx.calls == [([30, 12, 20, 25], True)] and len(full) == 4 * 8

    sess.ctx.calls.clear()
    sess.run(mapfiles, str(tmp_path / "chunks"), max_scores_bytes=8 * (30 * 30 + 12 * 12))
    assert sess.ctx.calls == [([30, 12], True), ([20, 25], True)]
    sess.ctx.calls.clear()
    sess.run(mapfiles, str(tmp_path / "one"), max_scores_bytes=1)                # a file larger than the budget still runs, alone
    assert sess.ctx.calls == [([30], True), ([12], True), ([20], True), ([25], True)]

    sess.ctx.calls.clear()
    lean = sess.run(mapfiles, str(tmp_path / "lean"), lean=True)
    assert sess.ctx.calls == [([30], True), ([12, 20, 25], False)]
    got = sorted(os.listdir(tmp_path / "lean"))
    assert len(lean) == len(got) == 8 + 3 * 6
    assert [f for f in got if "individualscores" in f or "indjoined" in f] == ["f5ep.individualscores.0", "f5ep.indjoined.0"]
    assert sum(".indmeasure." in f for f in got) == 4                      # what 2-collate pools (2-collate-results.py:67-76)

    sess.ctx.calls.clear()
    none = sess.run(mapfiles, str(tmp_path / "none"), full_dump=False)
    assert sess.ctx.calls == [([30, 12, 20, 25], False)] and len(none) == 4 * 6
