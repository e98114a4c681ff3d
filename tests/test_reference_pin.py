"""The oracle pinned against the reference's OWN Python executed in the build container.

tests/golden/reference_run_*.json.gz hold inputs and the seven objects the reference's `1-make-network.py` (its text,
mechanically rewritten 2 -> 3, driving its unmodified `coexfunctions.py` and its unmodified C file) dumped for them;
tests/golden/reference_functions.json holds known answers of `coexfunctions` helpers called directly
(tests/golden/run_reference_py3.py, make_reference_golden.py).  Nothing here reads /root/reference.

The *faithful* oracle mode (hit-vs-hit statistic from scipy, as the reference computes it) must reproduce every object
bit for bit, dict orders included, when it is given the two dict orders the executed run had (quirk Q6: the reference's
results depend on them; Python 2 leaves them to the hash function).  The *consistent* mode the GPU path is gated on
differs from it only where SURVEY.md section 8c says it must (1-ulp scipy-vs-C flips of a bisect index).
"""
import glob
import gzip
import json
import math
import os

import numpy as np
import pytest

from oracle import network_oracle as no
from oracle import permute_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
RUNS = sorted(glob.glob(os.path.join(HERE, "golden", "reference_run_*.json.gz")))
OBJECTS = ("prom_to_join", "joining_instructions", "winners", "individualscores", "indjoined", "allpromoterscores", "indmeasure")


def _b(v):
    return v if isinstance(v, bool) else str(v).lower() == "true"


def _oracle_kwargs(settings):
    return dict(correlationmeasure=settings["correlationmeasure"], anticorrelation=_b(settings["anticorrelation"]),
                noiter=_b(settings["noiter"]), pval_to_join=float(settings["pval_to_join"]),
                soft_separation=float(settings["soft_separation"]), nsp=float(settings["nsp"]),
                useaverage=_b(settings["useaverage"]), selectionthreshold=float(settings["selectionthreshold"]),
                selectionthreshold_is_j=_b(settings["selectionthreshold_is_j"]))


def test_fixtures_present():
    assert len(RUNS) >= 7


@pytest.mark.parametrize("path", RUNS, ids=[os.path.basename(p)[14:-8] for p in RUNS])
def test_faithful_oracle_reproduces_executed_reference_bit_for_bit(path):
    with gzip.open(path, "rt") as f:
        fx = json.load(f)
    X = np.asarray(fx["X"], dtype=np.float64)
    addresses = {k: tuple(v) for k, v in fx["addresses"].items()}
    kw = _oracle_kwargs(fx["settings"])
    for run in fx["runs"]:
        promoters = list(run["snp_map"])                 # map-file first-appearance order (1-make-network.py:131-177)
        assert len(promoters) > 20
        got = no.make_network(fx["names"], X, promoters, addresses, fx["glist"], mode="faithful",
                              group_order=list(run["prom_to_join"]), indjoined_order="insertion", **kw)
        for obj in OBJECTS:
            assert got[obj] == run[obj], obj             # every float bit-identical (json round-trips doubles exactly)
            assert list(got[obj]) == list(run[obj]), obj + " key order"


@pytest.mark.parametrize("path", RUNS, ids=[os.path.basename(p)[14:-8] for p in RUNS])
def test_consistent_oracle_differs_only_by_documented_index_flips(path):
    """The mode the GPU path is compared with: same groups; every edge score either identical to the executed
    reference's or one bisect step (one null value, or one run of tied null values) away from it: the statistic is
    the C number instead of scipy's."""
    with gzip.open(path, "rt") as f:
        fx = json.load(f)
    X = np.asarray(fx["X"], dtype=np.float64)
    addresses = {k: tuple(v) for k, v in fx["addresses"].items()}
    kw = _oracle_kwargs(fx["settings"])
    run = fx["runs"][0]
    promoters = list(run["snp_map"])
    got = no.make_network(fx["names"], X, promoters, addresses, fx["glist"], mode="consistent", **kw)
    assert got["prom_to_join"] == {k: run["prom_to_join"][k] for k in sorted(run["prom_to_join"])}
    L = (len(fx["names"]) - 1) if float(fx["settings"]["nsp"]) > 0 else len(fx["glist"])
    n_diff = n_all = 0
    for a in run["individualscores"]:
        for b, v in run["individualscores"][a].items():
            w = got["individualscores"][a][b]
            n_all += 1
            if w != v:
                n_diff += 1
                # the bisect index moved over one null value, or over one run of TIED null values (duplicated rows):
                # p changes by a whole number of 1/L
                # (a score of exactly 0 is p = 1 or quirk Q4: p = 0 -> log raises -> 0)
                cands = [abs(10.0 ** -w - pv) * L for pv in ([1.0, 0.0] if v == 0 else [10.0 ** -v])]
                cands += [abs(pw - 10.0 ** -v) * L for pw in ([1.0, 0.0] if w == 0 else [])]
                assert any(abs(st - round(st)) < 1e-6 and 1 <= round(st) <= 64 for st in cands) or abs(w - v) < 1e-12, (a, b, v, w)
    assert n_all > 400 and n_diff < 0.6 * n_all


def test_file_readers_against_reference_functions(tmp_path):
    """read_expression_file / readfeaturecoordinates (coexfunctions.py:201-274) on the same text the reference's own
    functions parsed when the fixture was made: NA column dropped, duplicated label keeps its first row, BED header
    lines skipped, label = last column, {chrom: {start: greatest end}}."""
    from coexpression_b200.coexfunctions import read_expression_file, readfeaturecoordinates
    with open(os.path.join(HERE, "golden", "reference_functions.json")) as f:
        rd = json.load(f)["readers"]
    ep, bp = tmp_path / "e.expression", tmp_path / "f.bed"
    ep.write_text(rd["expr_text"]); bp.write_text(rd["bed_text"])
    table, header = read_expression_file(str(ep))
    assert list(table.names) == rd["names"] and [str(h) for h in header] == rd["header"]
    assert np.array_equal(np.asarray(table.values, dtype=np.float64), np.asarray(rd["values"], dtype=np.float64))
    assert np.array_equal(np.signbit(np.asarray(table.values, dtype=np.float64)), np.signbit(np.asarray(rd["values"])))
    pa, fd, sl = readfeaturecoordinates(str(bp))
    assert pa == rd["pa"] and sl == rd["sl"]
    assert {c: {str(k): v for k, v in d.items()} for c, d in fd.items()} == rd["fd"]


def test_reference_helper_functions_known_answers():
    with open(os.path.join(HERE, "golden", "reference_functions.json")) as f:
        fx = json.load(f)
    cr = {k: tuple(v) for k, v in fx["chromrange"].items()}
    cl = fx["chromlist"]
    for chrom, address, want in fx["get_gwad"]:
        assert po.get_gwad(chrom, address, cr) == want
    for gwad, want in fx["get_address"]:
        assert list(po.get_address(gwad, cr, cl)) == want
    for n_items, index, want in fx["listitem"]:
        d = po.listitem_index(n_items, index)           # may be negative: Python indexes from the end
        assert list(range(n_items))[d] == want
    assert len(fx["listitem"]) > 200
    for ranges, ss, want in fx["merge_ranges"]:
        assert [[int(a), int(b)] for a, b in no.merge_ranges([tuple(r) for r in ranges], ss)] == want
    for v, coll, want in fx["empiricalp"]:
        got = no.empiricalp(v, sorted(coll)) if coll else 1
        assert got == want
    for p, method, rej, corr in fx["fdr_correction"]:
        r, c = no.fdr_correction(p, 0.05, method)
        assert [bool(x) for x in r] == rej
        assert np.array_equal(np.asarray(c, dtype=np.float64), np.asarray(corr, dtype=np.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("path", RUNS, ids=[os.path.basename(p)[14:-8] for p in RUNS])
def test_gpu_path_against_executed_reference(path):
    """The CUDA path (through the C ABI) against what the reference's own Python produced: identical groups; every
    edge score identical or one documented bisect flip away (the GPU statistic is the number getPearson gives, not
    scipy's); densities equal to the consistent oracle's within 1e-6 (the executed reference's own densities differ from
    those only through the flips, checked on the CPU side above)."""
    from coexpression_b200.api import CoexContext, Options
    from coexpression_b200.results import perm_objects
    with gzip.open(path, "rt") as f:
        fx = json.load(f)
    X = np.asarray(fx["X"], dtype=np.float64)
    names = fx["names"]
    row_of = {nm: i for i, nm in enumerate(names)}
    addresses = {k: tuple(v) for k, v in fx["addresses"].items()}
    kw = _oracle_kwargs(fx["settings"])
    opt = Options(correlationmeasure=kw["correlationmeasure"], anticorrelation=kw["anticorrelation"], noiter=kw["noiter"],
                  useaverage=kw["useaverage"], nsp=kw["nsp"], pval_to_join=kw["pval_to_join"],
                  selectionthreshold=kw["selectionthreshold"], selectionthreshold_is_j=kw["selectionthreshold_is_j"],
                  soft_separation=int(kw["soft_separation"]))
    hits = [np.asarray([row_of[p] for p in run["snp_map"]], dtype=np.int32) for run in fx["runs"]]
    with CoexContext() as ctx:
        ctx.set_expression(X); ctx.prepare()
        ctx.set_features([addresses[nm][0] for nm in names], [addresses[nm][1] for nm in names],
                         [addresses[nm][2] for nm in names], names)
        ctx.set_global_list(np.asarray(fx["glist"], dtype=np.float64))
        res = ctx.network_batch(hits, opt, want_scores=True)
    L = (len(names) - 1) if kw["nsp"] > 0 else len(fx["glist"])
    for k, run in enumerate(fx["runs"]):
        got = perm_objects(res, k, names)
        assert got["prom_to_join"] == {g: run["prom_to_join"][g] for g in sorted(run["prom_to_join"])}
        assert got["joining_instructions"] == run["joining_instructions"]
        n_same = n_all = 0
        for a in run["individualscores"]:
            for b, v in run["individualscores"][a].items():
                w = got["individualscores"][a][b]
                n_all += 1
                if abs(w - v) <= 1e-12:
                    n_same += 1
                    continue
                cands = [abs(10.0 ** -w - pv) * L for pv in ([1.0, 0.0] if v == 0 else [10.0 ** -v])]
                cands += [abs(pw - 10.0 ** -v) * L for pw in ([1.0, 0.0] if w == 0 else [])]
                assert any(abs(st - round(st)) < 1e-6 and 1 <= round(st) <= 64 for st in cands), (a, b, v, w)
        assert n_same > 0.4 * n_all
        want = no.make_network(names, X, list(run["snp_map"]), addresses, fx["glist"], mode="consistent", **kw)
        assert got["winners"] == want["winners"]
        for g, v in want["indmeasure"].items():
            assert abs(got["indmeasure"][g] - v) <= 1e-6


# ---- row N2: the global correlation list against the executed block 0-prepare-coex.py:623-716 -----------------------
def _global_list_runs():
    with gzip.open(os.path.join(HERE, "golden", "reference_global_list.json.gz"), "rt") as f:
        return json.load(f)["runs"]


class _OraclePairs:
    """Stands in for CoexContext in the CPU test: `pairs` through the unmodified reference C build (or its port)."""
    def __init__(self, X):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.R = no.rank_rows(self.X)

    def pairs(self, ia, ib, ranked):
        return no.get_pearson(self.R if ranked else self.X, ia, ib)


def _check_global_list(run, ctx, tmp_path):
    from coexpression_b200.coexfunctions import make_correlation_list
    corrfile = tmp_path / ("%s.list" % run["case"])
    if run["existing"]:
        with open(corrfile, "w") as o:
            for v in run["existing"]:
                o.write("%s\n" % v)
    got = make_correlation_list(ctx, run["names"], str(corrfile), run["measure"], run["gpl_len"], seed=run["seed"])
    want = np.asarray(run["values"], dtype=np.float64)
    assert got.shape == want.shape and len(want) > run["gpl_len"]
    assert np.array_equal(got, want)                    # same pairs drawn, same getPearson bits, same order
    with open(corrfile) as f:
        on_disk = np.asarray([float(x) for x in f if x.strip()], dtype=np.float64)
    assert np.array_equal(on_disk, want)                # the file 1-make-network.py:76-79 reads back


@pytest.mark.parametrize("k", [0, 1, 2])
def test_global_list_host_logic_reproduces_executed_reference(k, tmp_path):
    run = _global_list_runs()[k]
    _check_global_list(run, _OraclePairs(run["X"]), tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 1, 2])
def test_gpu_global_list_reproduces_executed_reference(k, tmp_path):
    from coexpression_b200.api import CoexContext
    run = _global_list_runs()[k]
    with CoexContext() as ctx:
        ctx.set_expression(np.asarray(run["X"], dtype=np.float64)); ctx.prepare()
        _check_global_list(run, ctx, tmp_path)


# ---- row N1: permuted locus positions against the executed permutation block 0-prepare-coex.py:471-613 --------------
def _permutation_runs():
    with open(os.path.join(HERE, "golden", "reference_permutations.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("mode", ["circular", "backcirc", "background"])
def test_permuted_positions_reproduce_executed_reference_block(mode):
    """The positions the reference hands to windowBed for every locus set (its temporary BED, captured by standing in
    for the windowBed binary) against oracle/permute_oracle.py, which the CUDA path (coex_permute_map) is compared with in
    tests/test_permute.py.  The reference keeps the loci of a set in a {chrom: {address: name}} dict, so two loci that land
    on one position collapse and the order is the dict's: positions are compared as sets."""
    import random
    fx = _permutation_runs()
    cr = {k: tuple(v) for k, v in fx["chromrange"].items()}
    cl = fx["chromlist"]
    run = next(r for r in fx["runs"] if r["mode"] == mode)
    snps = [(c, int(a)) for c, a in run["snps"]]
    backlines = [l.rstrip("\n").split("\t") for l in run["backlines"]]
    rnd = random.Random(run["seed"])
    if mode == "backcirc":
        backdict = {}
        for line in backlines:                                          # 0-prepare-coex.py:501-519: column 1 is the address
            chrom, address = "chr" + line[0].replace("chr", ""), int(float(line[1]))
            backdict["%s_%s" % (chrom, address)] = po.get_gwad(chrom, address, cr)
        for chrom, address in snps:                                     # :521-528: the real loci join the background
            backdict.setdefault("%s_%s" % (chrom, address), po.get_gwad(chrom, address, cr))
        sb = po.sorted_background(backdict)
    assert [s["shift"] for s in run["sets"]][-1] == 0 and len(run["sets"]) == run["numperms"] + 1
    if mode != "background":                                            # :476-497: the shifts come first off the RNG stream
        assert [s["shift"] for s in run["sets"]][:-1] == [int(rnd.random() * 10000000000) for _ in range(run["numperms"])]
    for k, s in enumerate(run["sets"]):
        got = {(p[0], p[1]) for p in s["positions"]}
        assert all(p[2] == p[1] + 1 for p in s["positions"])            # A = [address, address + 1)
        if s["shift"] == 0 or mode == "circular":
            want = {p for p in po.circular_positions(snps, s["shift"], cr, cl) if p is not None}
            assert all(p[3].startswith("rand_") == (s["shift"] != 0) for p in s["positions"])
        elif mode == "backcirc":
            want = set(po.backcirc_positions(snps, s["shift"], sb))
            if got != want:
                # Python 3 rounds exact halves to even, Python 2.7 (and the restatement) away from zero: the executed run can
                # differ from a real Python-2 run -- and from the restatement -- only on such indices
                n = len(sb)
                pos = {key: i for i, key in enumerate(sb)}
                halves = [c for c, a in snps if abs(((pos["%s_%s" % (c, a)] + s["shift"]) / n) % 1 - 0.5) < 1e-12 or
                          abs(((((pos["%s_%s" % (c, a)] + s["shift"]) / n) - round((pos["%s_%s" % (c, a)] + s["shift"]) / n)) * n) % 1 - 0.5) < 1e-9]
                assert halves and len(got ^ want) <= 2 * len(halves), (k, got ^ want)
                continue
        else:
            picks = rnd.sample(range(len(backlines)), len(snps))        # :568, one draw per permuted set, in loop order
            want = set(po.background_positions(backlines, picks))
        assert got == want, (mode, k, s["shift"], sorted(got ^ want)[:6])


# ---- row N3: collation against the executed 2-collate-results.py ------------------------------------------------------
def test_collate_reproduces_executed_reference_tables(tmp_path):
    """coexpression_b200/collate.py on the perm_results_store the reference's own 1-make-network.py wrote, against the
    `<label>.individual_scores` table and `allscores.txt` the reference's own 2-collate-results.py made of it: pooled null,
    empirical p, Benjamini-Hochberg column, score / number of groups, the (score, name) order -- text for text."""
    from coexpression_b200.collate import collate
    with gzip.open(os.path.join(HERE, "golden", "reference_collate.json.gz"), "rt") as f:
        fx = json.load(f)
    wd = tmp_path / "work"
    (wd / "perm_results_store").mkdir(parents=True)
    for name, text in fx["store"].items():
        (wd / "perm_results_store" / name).write_text(text)
    bed = tmp_path / "f5ep.bed"
    bed.write_text(fx["feature_bed"])
    settings = dict(fx["settings"], feature_coordinates=str(bed))
    (wd / "settings.txt").write_text(json.dumps(settings))
    res = collate(str(wd))
    assert res["num_completed_perms"] == 3
    assert open(res["individual_scores_file"]).read() == fx["individual_scores"]
    assert (wd / "allscores.txt").read_text() == fx["allscores"]
    assert len(fx["individual_scores"].splitlines()) > 40
    assert "reports_error" in res and "reports" not in res          # a store without the full settings: tables only


def test_collate_reports_reproduce_executed_reference_files(tmp_path):
    """The reporting files of 2-collate-results.py:144-430 (html table, circos network, BioLayout / d3 layout, stats, file list,
    the header-only qq_plot file) against what the executed reference wrote into its working directory -- text for text, except
    that the layout file's //NODECLASS lines follow a Python set's iteration order in the reference (compared as a set here)."""
    from coexpression_b200.collate import collate
    with gzip.open(os.path.join(HERE, "golden", "reference_collate.json.gz"), "rt") as f:
        fx = json.load(f)
    wd = tmp_path / "work"
    (wd / "perm_results_store").mkdir(parents=True)
    (wd / "permutation_store").mkdir()
    for name, text in fx["store"].items():
        (wd / "perm_results_store" / name).write_text(text)
    (wd / "permutation_store" / "f5ep.0.bed").write_text(fx["map0"])
    bed = tmp_path / "f5ep.bed"
    bed.write_text(fx["feature_bed"])
    cfg = tmp_path / "app.cfg"
    cfg.write_text("[directorypaths]\nsourcefilesdir=x\nresdir=x\npathtobedtools=/nonexistent\nf5resource=%s\n" % fx["f5resource"])
    settings = dict(fx["full_settings"], feature_coordinates=str(bed), config_file=str(cfg))
    (wd / "settings.txt").write_text(json.dumps(settings))
    res = collate(str(wd))
    assert sorted(res["reports"]) == sorted(fx["reports"])
    for name, want in fx["reports"].items():
        got = (wd / name).read_text()
        if name.endswith(".layout"):
            split = lambda text: ([l for l in text.split("\n") if not l.startswith("//NODECLASS\t")],
                                  sorted(l for l in text.split("\n") if l.startswith("//NODECLASS\t")))
            assert split(got) == split(want), name
            assert got.count("//NODECLASS\t") > 20
        else:
            assert got == want, name
    assert (wd / "golden.individual_scores").read_text() == fx["individual_scores"]


def test_py2_rewrite_is_line_local_and_mechanical():
    """The 2 -> 3 rewrite used to execute the reference's scripts (tests/golden/run_reference_py3.py): only the listed
    syntactic forms change, everything else passes through byte for byte."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from run_reference_py3 import convert_py2, py27_sum
    src = "\n".join([
        'if verbose: print "x", y  # note',
        'print "a %s" % (len(',
        '    z))',
        'print ("already", "a call")',
        'for k, v in d.iteritems():',
        '    for i in xrange(3): pass',
        'ptcf = exp_dict.keys()',
        'order = sorted(m.iteritems(), key=lambda (k, v): (v, k), reverse=True)',
        'import ConfigParser',
        'x = 1 / 2  # untouched: no arithmetic is rewritten',
    ])
    out = convert_py2(src).split("\n")
    assert out == [
        'if verbose: print("x", y)  # note',
        'print("a %s" % (len( z)))',
        'print ("already", "a call")',
        'for k, v in d.items():',
        '    for i in range(3): pass',
        'ptcf = list(exp_dict.keys())',
        'order = sorted(m.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)',
        'import configparser as ConfigParser',
        'x = 1 / 2  # untouched: no arithmetic is rewritten',
    ]
    vals = [0.1] * 10 + [1e16, -1e16]
    acc = 0
    for v in vals:
        acc = acc + v
    assert py27_sum(vals) == acc                         # plain left-to-right, unlike Python >= 3.12's compensated sum


@pytest.mark.parametrize("path", RUNS, ids=[os.path.basename(p)[14:-8] for p in RUNS])
def test_map_file_reader_reproduces_executed_reference_snp_map(path, tmp_path):
    """read_snp_map (1-make-network.py:131-164) on the windowBed-style map files the executed reference read: same
    promoters in the same (first-appearance) order, same `snps` strings, `p` left at -9999; the line of a promoter that is
    not in the expression table and the line with fewer than four columns are skipped."""
    from coexpression_b200.coexfunctions import read_snp_map
    with gzip.open(path, "rt") as f:
        fx = json.load(f)
    table = set(fx["names"])
    for k, text in enumerate(fx["mapfiles"]):
        assert "not_in_table" in text and "too\tshort\tline" in text
        mf = tmp_path / ("f5ep.%d.bed" % k)
        mf.write_text(text)
        got = read_snp_map(str(mf), table)
        want = fx["runs"][k]["snp_map"]
        assert got == want and list(got) == list(want)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["spearman_default", "pearson_default", "spearman_global_nsp0"])
def test_gpu_file_level_drop_in_against_executed_reference(case, tmp_path):
    """The whole drop-in at the file level: the files the reference's `1-make-network.py` read when the fixture was made
    (settings.txt, expression table, feature BED, correlation list, windowBed map files) go into
    coexpression_b200.make_network.run_batch, and the JSON objects it writes into perm_results_store/ are compared with
    the ones the reference wrote: snp_map, groups and joining instructions identical, every edge score identical or one
    documented bisect flip away."""
    from coexpression_b200.make_network import run_batch
    path = os.path.join(HERE, "golden", "reference_run_%s.json.gz" % case)
    with gzip.open(path, "rt") as f:
        fx = json.load(f)
    names, X = fx["names"], np.asarray(fx["X"], dtype=np.float64)
    wd = tmp_path / "work"
    (wd / "permutation_store").mkdir(parents=True)
    expr, bed, clist = tmp_path / "t.expression", tmp_path / "f5ep.bed", tmp_path / "t.expression.list"
    with open(expr, "w") as o:
        o.write("sample\t" + "\t".join("s%d" % j for j in range(X.shape[1])) + "\n")
        for nm, row in zip(names, X):
            o.write(nm + "\t" + "\t".join(repr(float(v)) for v in row) + "\n")
    with open(bed, "w") as o:
        for nm in names:
            c, s, e = fx["addresses"][nm]
            o.write("%s\t%s\t%s\t%s\n" % (c, s, e, nm))
    with open(clist, "w") as o:
        for v in fx["glist"]:
            o.write("%s\n" % (int(v) if v == -99 else repr(float(v))))
    for k, text in enumerate(fx["mapfiles"]):
        (wd / "permutation_store" / ("f5ep.%d.bed" % k)).write_text(text)
    settings = dict(fx["settings"], feature_coordinates=str(bed), correlationfile=str(clist), expression_file=str(expr),
                    config_file="app.cfg")
    (wd / "settings.txt").write_text(json.dumps(settings))
    written = run_batch(str(wd))
    assert len(written) == 8 * len(fx["mapfiles"])
    store = wd / "perm_results_store"
    L = (len(names) - 1) if float(fx["settings"]["nsp"]) > 0 else len(fx["glist"])
    for k, run in enumerate(fx["runs"]):
        got = {obj: json.load(open(store / ("f5ep.%s.%d" % (obj, k)))) for obj in
               ("snp_map", "prom_to_join", "joining_instructions", "individualscores", "winners", "indmeasure")}
        assert got["snp_map"] == run["snp_map"] and list(got["snp_map"]) == list(run["snp_map"])
        assert got["prom_to_join"] == run["prom_to_join"]
        assert got["joining_instructions"] == run["joining_instructions"]
        assert set(got["winners"]) == set(run["winners"]) and set(got["indmeasure"]) == set(run["indmeasure"])
        n_same = n_all = 0
        for a in run["individualscores"]:
            for b, v in run["individualscores"][a].items():
                w = got["individualscores"][a][b]
                n_all += 1
                if abs(w - v) <= 1e-12:
                    n_same += 1
                    continue
                cands = [abs(10.0 ** -w - pv) * L for pv in ([1.0, 0.0] if v == 0 else [10.0 ** -v])]
                cands += [abs(pw - 10.0 ** -v) * L for pw in ([1.0, 0.0] if w == 0 else [])]
                assert any(abs(st - round(st)) < 1e-6 and 1 <= round(st) <= 64 for st in cands), (a, b, v, w)
        assert n_all > 1000 and n_same > 0.4 * n_all
